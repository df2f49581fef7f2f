"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy/scipy restatement of the reference's sequencing-batch-reactor hot path, written from the
reference's behaviour (file:line citations on every function, all relative to /root/reference/gym_SBR/envs).
It is the checker the CUDA path is compared with; it is also timed as the CPU baseline (`bench.py`,
`cpu_baseline.kind == "port"`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product package `gym_sbr2_b200` never does.

Arithmetic: like the reference, every ODE solve is `scipy.integrate.odeint` (ODEPACK LSODA, default
rtol = atol = 1.49e-8) called once per PID interval with a Python RHS callback -- the integrator is a
third-party dependency the reference does not pin (setup.py:3 lists only `gym`); the golden vectors in
tests/golden/ were produced by the *unmodified reference* under scipy 1.18.1 / numpy 2.3.5 (recorded in the
fixtures).  Parity pin: tests/test_oracle_golden.py checks this file against those reference outputs
(whole-cycle state, reward, OCI, Qw, EQI, per-phase interval counts; per-step rewards/obs/done for SBROS-v1).
The reference holds no tests or golden vectors of its own, so these fixtures, generated
from the reference itself by oracle/make_golden.py, are the pin.

State vector (all files): 0=V 1=Si 2=Ss 3=Xi 4=Xs 5=Xbh 6=Xba 7=Xp 8=So 9=Sno 10=Snh 11=Snd 12=Xnd 13=Salk
(SBR_model_FB.py:199-203).  Time unit: day.
"""
import math

import numpy as np
from scipy.integrate import odeint

# ----------------------------------------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------------------------------------
WV = 1.32                                   # gym_SBR_env2.py:33
IV = 0.6161484733495801                     # gym_SBR_env2.py:85
QIN = WV - IV                               # gym_SBR_env2.py:93
QEFF = 0.66                                 # SBR_model_FB.py:226
BIOMASS_SETPOINT = 2700                     # SBR_model_FB.py:224
T_CYCLE = 12 / 24                           # gym_SBR_env2.py:39
T_RATIO = [4.2 / 100, 8.3 / 100, 37.5 / 100, 31.2 / 100, 2.1 / 100, 8.3 / 100, 2.1 / 100, 6.3 / 100]
DT = 0.002 / 24                             # gym_SBR_env2.py:36 (`t_delta`)
X0_INIT = [0.6161484733495801, 30, 0.571098000538576, 1440.01157895393,
           31.254221999137, 2599.2714348941, 168.915006750837, 551.901552960823, 2.16607843793004,
           13.3791460027604, 0.00562880208518134, 0.35996687629947, 1.86916737961228,
           3.790463057094611]              # gym_SBR_env2.py:78-80
SPAR = dict(Ya=0.24, Yh=0.67, fp=0.08, ixb=0.08, ixp=0.06)                       # SBR_model_FB.py:36
KPAR = dict(muh=4.0, Ks=10.0, Koh=0.2, Kno=0.5, bh=0.3, etag=0.8, etah=0.8, kh=3.0, Kx=0.1,
            mua=0.5, Knh=1.0, ba=0.05, Koa=0.4, ka=0.05)                          # SBR_model_FB.py:38


def do_saturation(temp_c=15.0):
    """Oxygen saturation vs temperature (module_temperature.py:3-20); 8.000000000006622 at 15 C."""
    tk = (temp_c + 273.15) / 100
    f = 56.12 * np.exp(-66.7354 + 87.4755 / tk + 24.4526 * np.log(tk))
    return 0.9997743214 * (8 / 10.5) * 6791.5 * f


SO_SAT = float(do_saturation(15))
# Path-A controller: Kc, tauI, PID dt, Kla_min, Kla_max, tauD  (gym_SBR_env2.py:48)
PID_A = dict(Kc=5.0, tauI=0.00035, dt=0.02 / 24, lo=0.0, hi=240.0, tauD=0.005)


# ----------------------------------------------------------------------------------------------------------
# kinetics: one RHS, three tails
# ----------------------------------------------------------------------------------------------------------
def asm1_rates(x, kla, so_sat=SO_SAT):
    """Component conversion rates r[1..13] of the ASM1-type model incl. aeration (sub_phases_FB.py:278-372).

    Quirks kept on purpose: `ixp` is used where ASM1 has `fp` in the decay stoichiometry
    (sub_phases_FB.py:325-333).
    """
    K, S = KPAR, SPAR
    Ss, Xs, Xbh, Xba, So, Sno, Snh, Snd, Xnd = x[2], x[4], x[5], x[6], x[8], x[9], x[10], x[11], x[12]
    # process rates (sub_phases_FB.py:280-303)
    rho1 = K['muh'] * (Ss / (K['Ks'] + Ss)) * (So / (K['Koh'] + So)) * Xbh
    rho2 = K['muh'] * (Ss / (K['Ks'] + Ss)) * (K['Koh'] / (So + K['Koh'])) * (Sno / (K['Kno'] + Sno)) * K['etag'] * Xbh
    rho3 = K['mua'] * (Snh / (K['Knh'] + Snh)) * (So / (K['Koa'] + So)) * Xba
    rho4 = K['bh'] * Xbh
    rho5 = K['ba'] * Xba
    rho6 = K['ka'] * Snd * Xbh
    rho7 = K['kh'] * ((Xs / Xbh) / (K['Kx'] + (Xs / Xbh))) * (
        (So / (K['Koh'] + So)) + K['etah'] * (K['Koh'] / (So + K['Koh'])) * (Sno / (K['Kno'] + Sno))) * Xbh
    rho8 = (Xnd / Xs) * rho7
    Ya, Yh, fp, ixb, ixp = S['Ya'], S['Yh'], S['fp'], S['ixb'], S['ixp']
    r = np.zeros(14)
    r[2] = (-1 / Yh) * rho1 + (-1 / Yh) * rho2 + rho7
    r[4] = (1 - ixp) * rho4 + (1 - ixp) * rho5 - rho7
    r[5] = rho1 + rho2 - rho4
    r[6] = rho3 - rho5
    r[7] = ixp * rho4 + ixp * rho5
    r[8] = (-(1 - Yh) / Yh) * rho1 + (-(4.57 - Ya) / Ya) * rho3 + kla * (so_sat - So)
    r[9] = (-((1 - Yh) / (2.86 * Yh))) * rho2 + (1 / Ya) * rho3
    r[10] = (-ixb) * rho1 + (-ixb) * rho2 + (-ixb - 1 / Ya) * rho3 + rho6
    r[11] = -rho6 + rho8
    r[12] = (ixb - fp * ixp) * rho4 + (ixb - fp * ixp) * rho5 - rho8
    r[13] = (-ixb / 14) * rho1 + ((1 - Yh) / (14 * 2.86 * Yh) - ixb / 14) * rho2 \
        + (-ixb / 14 - 1 / (7 * Ya)) * rho3 + (1 / 14) * rho6
    return r


def rhs_react(x, t, kla):
    """React tail: dx_i/dt = r_i, dV/dt = 0 (sub_phases_FB.py:374-404)."""
    return asm1_rates(x, kla)


def rhs_fill(x, t, kla, loading):
    """Fill tail: dV/dt = q, dx_i/dt = r_i + (q/V)(c_in,i - x_i) (sub_phases_FB.py:146-176)."""
    d = asm1_rates(x, kla)
    q = loading[0]
    for i in range(1, 14):
        d[i] = d[i] + (q / x[0]) * (loading[i] - x[i])
    d[0] = q
    return d


def rhs_react_ec(x, t, kla, ec, ec_conc):
    """React + external-carbon dosing tail (gym_SBR_oneshot.py:1757-1787)."""
    d = asm1_rates(x, kla)
    for i in range(1, 14):
        d[i] = d[i] + (ec / x[0]) * (-x[i])
    d[2] = d[2] + (ec / x[0]) * ec_conc
    d[0] = ec
    return d


# ----------------------------------------------------------------------------------------------------------
# Path A (SBR-v2): PID-controlled phases, settle, draw, reward
# ----------------------------------------------------------------------------------------------------------
def phase_grid(t_start, t_end, t_delta=DT):
    """Interval boundaries of one PID-controlled phase (sub_phases_FB.py:183-184, 226-231).

    Returns (t_save2, points_per_interval).  Both sizes come from int(float) truncation in the reference.
    """
    t_save2 = np.linspace(t_start, t_end, int((t_end - t_start) / (t_delta * 10)))
    pts = [int((t_save2[i + 1] - t_save2[i]) / t_delta) for i in range(len(t_save2) - 1)]
    return t_save2, pts


def cycle_schedule():
    """[(t_start, t_end)] for the 8 phases exactly as SBR_model_FB.run sequences them (SBR_model_FB.py:17-25,
    60-264): phase 1 starts at 0, every later phase starts at the previous end + t_delta."""
    out = []
    t_end = 0
    for k in range(8):
        t_start = t_end if k == 0 else t_end + DT
        t_end = t_start + T_CYCLE * T_RATIO[k]
        out.append((t_start, t_end))
    return out


def pid_phase(x, t_start, t_end, sp, kla_bias, rhs, rhs_args=(), pid=PID_A, ode_kw=None, tap=None, raw_kla=False):
    """One PID-controlled phase: filling.sim_rxn / rxn.sim_rxn (sub_phases_FB.py:178-271, 406-500).

    Positional PID on So sampled at interval starts; the bias is the *clamped* output of interval 0
    (Kla[0] is overwritten, sub_phases_FB.py:218,243); two independent clamp checks with anti-windup
    (:245-250).  Returns (x_end, Kla per interval).
    """
    ode_kw = ode_kw or {}
    t_save2, pts = phase_grid(t_start, t_end)
    n = len(t_save2) - 1
    Kc, tauI, tauD, dtc = pid['Kc'], pid['tauI'], pid['tauD'], pid['dt']
    So = np.zeros(n)
    e = np.zeros(n)
    ie = np.zeros(n)
    dcv = np.zeros(n)
    Kla = np.zeros(n)
    So[0] = x[8]
    Kla[0] = kla_bias
    x = np.asarray(x, dtype=float)
    for i in range(n):
        t_range = np.linspace(t_save2[i], t_save2[i + 1], pts[i])
        e[i] = sp - So[i]
        if i >= 1:
            dcv[i] = (So[i] - So[i - 1]) / dtc
            ie[i] = ie[i - 1] + e[i] * dtc
        Kla[i] = Kc * e[i] + Kc / tauI * ie[i] + Kc * tauD * dcv[i] + Kla[0]
        if raw_kla:
            # PID bypass (no reference env): `sp` IS the KLa, held over the phase -- the same odeint-per-interval
            # drive as the open-loop sim_rxn of sub_phases_PID_off.py:178-225 with the KLa fixed
            Kla[i] = sp
        if Kla[i] > pid['hi']:
            Kla[i] = pid['hi']
            ie[i] = ie[i] - e[i] * dtc
        if Kla[i] < pid['lo']:
            Kla[i] = pid['lo']
            ie[i] = ie[i] - e[i] * dtc
        soln = odeint(rhs, x, t_range, args=(Kla[i],) + tuple(rhs_args), **ode_kw)
        if tap is not None:
            tap.append(dict(x0=np.array(x), kla=float(Kla[i]), T=float(t_range[-1] - t_range[0]),
                            pts=len(t_range), x1=np.array(soln[-1]), args=rhs_args))
        if i < n - 1:
            So[i + 1] = soln[-1][8]
        x = soln[-1]
    return x, Kla


def settler_rhs(sX, t, z, Xf):
    """10-layer solids flux model (sub_phases_FB.py:622-714).  The reference takes max(vmax, ...), so the
    settling velocity is the constant vmax = 474 m/d in every layer (kept as written)."""
    vmax, rh, rp, fns = 474, 0.000576, 0.00286, 0.00228
    v = np.array([max(vmax, np.exp(-rh * (s - fns * Xf)) - np.exp(-rp * (s - fns * Xf))) for s in sX])
    J = v * sX
    d = np.zeros_like(sX)
    d[0] = J[1] / z
    for i in range(1, 9):
        d[i] = (J[i + 1] - J[i]) / z
    d[9] = (0 - J[9]) / z
    return d


def settle(x, t_start, t_end, ode_kw=None):
    """settling.sim_settling (sub_phases_FB.py:716-775) without the unused Xnd layers (their result is never
    read downstream: SBR_model_FB.py:196,234).  Returns (sX[10], Xf)."""
    ode_kw = ode_kw or {}
    t_save = np.linspace(t_start, t_end, int((t_end - t_start) / DT))
    Xf = 0.75 * (x[3] + x[4] + x[5] + x[6] + x[7])
    As = (1.25 / 2) ** 2
    z = x[0] / As
    sol = odeint(settler_rhs, [Xf] * 10, t_save, args=(z, Xf), **ode_kw)
    return sol[-1], Xf


def effluent_quality(xe):
    """drawing.cal_eq (sub_phases_FB.py:868-915): EQI and eff = [0.66, Ntot, COD, Snh, BOD5, Sno]."""
    Si, Ss, Xi, Xs, Xbh, Xba, Xp, So, Sno, Snh, Snd, Xnd = xe[1:13]
    Snkj = Snh + Snd + Xnd + 0.08 * (Xbh + Xba) + 0.06 * (Xp + Xi)
    Ntot = Sno + Snkj
    SS = 0.75 * (Xs + Xi + Xbh + Xba + Xp)
    BOD5 = 0.25 * (Ss + Xs + (1 - 0.08) * (Xbh + Xba))
    COD = Ss + Si + Xs + Xi + Xbh + Xba + Xp
    EQI = (2 * SS + 1 * COD + 30 * Snkj + 10 * Sno + 2 * BOD5) * (1 / 1000) * 0.66
    return EQI, [0.66, Ntot, COD, Snh, BOD5, Sno]


def draw(x, sX, Xf, Qeff=QEFF, biomass_setpoint=BIOMASS_SETPOINT):
    """drawing.sim_drawing (sub_phases_FB.py:780-864): decant `Qeff` from the top layers (the very top layer
    is left out of the effluent-solids sum by the `[-m:-1]` slice, :794), then waste sludge bottom-up until the
    remaining solids equal biomass_setpoint * remaining volume.  Returns (x_after, Qw, EQI, eff, status) where
    status = 1 flags the reference's unassigned-`Qw` path (loop ends without reaching the `else`, :817-836)."""
    x = np.array(x, dtype=float)
    sX = np.array(sX, dtype=float)
    V0 = x[0]
    lv = V0 / 10
    resV = V0 - Qeff
    m = int(math.ceil(round(Qeff / lv)))
    sX_eff = sum(sX[-m:-1] * lv)
    xe = x.copy()
    xe[0] = Qeff
    for i in (4, 7, 3, 5, 6):
        xe[i] = xe[i] * (1 / 0.75) * sX_eff / Xf
    res = sX[0:10 - m].copy()
    w_layer = lv * res
    waste = sum(w_layer) - biomass_setpoint * resV
    Qw = float('nan')
    status = 1
    for i in range(10 - m):
        left = waste - w_layer[i]
        if left > 0:
            waste = left
            res[i] = 0
            w_layer[i] = 0
            resV -= lv
        else:
            Qw = waste / (res[i] - biomass_setpoint)
            w_layer[i] = w_layer[i] - Qw * res[i]
            resV -= Qw
            res[i] = w_layer[i] / (lv - Qw)
            status = 0
            break
    sX2 = sum(w_layer) / resV
    x7 = x.copy()
    x7[0] = resV
    for i in (4, 7, 3, 5, 6):
        x7[i] = x[i] * (1 / 0.75) * sX2 / Xf
    EQI, eff = effluent_quality(xe)
    return x7, Qw, EQI, eff, status


def reward_v2(kla3, kla5, kla8, Qw, Qin, Qeff, Snh, so_sat=SO_SAT):
    """module_reward.sbr_reward (module_reward.py:4-51) -> (reward, OCI)."""
    t_delta = DT
    ME = 0.005 * 1.32 * 24 + 0.005 * 1.32 * 24
    AE_3 = 1.32 * sum(kla3) * t_delta / (len(kla3) * t_delta)
    AE_5 = 1.32 * sum(kla5) * t_delta / (len(kla5) * t_delta)
    AE_8 = (1.32 - Qw) * sum(kla8) * t_delta / (len(kla8) * t_delta)
    AE = so_sat / (1.8 * 1000) * (AE_3 + AE_5 + AE_8)
    PE = 0.004 * Qin + 0.05 * Qw + 0.004 * Qeff
    OCI = AE + PE + ME
    r_snh = 0 if Snh < 4 else -20
    return (5 - OCI) + r_snh, OCI


def fill_flow():
    """Fill flow rate written into influent_mixed[0] by SbrEnv2.step (gym_SBR_env2.py:144)."""
    return QIN / (T_CYCLE * T_RATIO[0])


def cycle_v2(setpoints3, influent, x0=X0_INIT, kla0=0.0, ode_kw=None, tap=None, raw_kla=False):
    """One whole 12-h cycle = SBR_model_FB.run (SBR_model_FB.py:8-295) for the SBR-v2 env.

    setpoints3: DO set-points (g/m3) of phases 3, 5 and 8 (already scaled, i.e. 8*action).
    influent:   14-vector, [0] = fill flow (m3/d), [1:] = influent concentrations.
    Set-points of the other reacting phases are 0 (gym_SBR_env2.py:54,184-186); each phase's bias is the last
    Kla of the previous reacting phase, the idle phase takes phase 5's (SBR_model_FB.py:94,120,146,172,266).
    """
    sched = cycle_schedule()
    sp = [0, 0, setpoints3[0], 0, setpoints3[1], 0, 0, setpoints3[2]]
    x = np.array(x0, dtype=float)
    klas = {}
    kla = kla0
    taps = {k: ([] if tap is not None else None) for k in range(8)}
    for k in range(5):
        if k == 0:
            x, K = pid_phase(x, sched[k][0], sched[k][1], sp[k], kla, rhs_fill, (list(influent),),
                             ode_kw=ode_kw, tap=taps[k], raw_kla=raw_kla)
        else:
            x, K = pid_phase(x, sched[k][0], sched[k][1], sp[k], kla, rhs_react, ode_kw=ode_kw, tap=taps[k],
                             raw_kla=raw_kla)
        klas[k] = K
        kla = K[-1]
    x5 = x
    sX, Xf = settle(x5, sched[5][0], sched[5][1], ode_kw=ode_kw)
    x7, Qw, EQI, eff, status = draw(x5, sX, Xf)
    x8, K8 = pid_phase(x7, sched[7][0], sched[7][1], sp[7], klas[4][-1], rhs_react, ode_kw=ode_kw, tap=taps[7],
                       raw_kla=raw_kla)
    klas[7] = K8
    if tap is not None:
        tap.update(taps)
    return dict(x_last=np.array(x8), x5=np.array(x5), x7=np.array(x7), sX=np.array(sX), Xf=Xf, Qw=Qw, EQI=EQI,
                eff=eff, kla=klas, status=status,
                n_intervals=[len(klas[k]) if k in klas else 0 for k in range(8)])


def sbr_v2_step(action, influent_mixed, x0=X0_INIT, ode_kw=None, tap=None, raw_kla=False):
    """SbrEnv2.step (gym_SBR_env2.py:131-171): returns dict(obs[3], reward, done, OCI, ...).
    raw_kla=True: the actions are KLa values (fractions of 240 1/d) of phases 3, 5, 8, the PID is bypassed."""
    a = np.clip(np.asarray(action, dtype=float), 0.0, 1.0)
    infl = list(influent_mixed)
    infl[0] = fill_flow()
    scale = PID_A['hi'] if raw_kla else 8
    out = cycle_v2([a[0] * scale, a[1] * scale, a[2] * scale], infl, x0=x0, ode_kw=ode_kw, tap=tap, raw_kla=raw_kla)
    Snh = out['eff'][3]
    reward, OCI = reward_v2(out['kla'][2], out['kla'][4], out['kla'][7], out['Qw'], QIN, QEFF, Snh)
    out.update(reward=reward, OCI=OCI, done=True,
               obs=np.array([QEFF, out['eff'][2], out['eff'][3] / 30]))
    return out


def sbr_v2_reset_obs(influent_mixed, x0=X0_INIT):
    """SbrEnv2.reset observation (gym_SBR_env2.py:108-118): element-wise SUM of x0 and influent_mixed."""
    s = np.asarray(x0, dtype=float) + np.asarray(influent_mixed, dtype=float)
    cod = s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7]
    return np.array([s[0], (cod - 5145) / 10, s[10] / 30])


# ----------------------------------------------------------------------------------------------------------
# Path B (SBROS-v1): interval-per-step env with DO-PID (KLa) and NO3-PID (external carbon)
# ----------------------------------------------------------------------------------------------------------
OS_DT = 0.002 / 24                      # gym_SBR_oneshot.py:30  (`dt`)
OS_T_DELTA = OS_DT * 10                 # gym_SBR_oneshot.py:31  (`t_delta`, one control interval)
OS_PID = dict(Kc_DO=100, tauI_DO=20, tauD_DO=0, Kc_EC=100, tauI_EC=20, tauD_EC=0,
              kla_lo=0, kla_hi=240, ec_lo=0, ec_hi=0.0005)              # gym_SBR_oneshot.py:80-94
EC_CONC = 1200000 * 4                   # gym_SBR_oneshot.py:96
X1_STATE = np.array([0.5, 1.32, 30, 30, 1500, 150, 3000, 2000, 600, 8, 20, 20, 10, 10, 10])   # :153
OBS_IDX_DO, X1_DO = [0, 5, 6, 8, 10], np.array([0.5, 2000, 500, 8., 10])                        # :150,155
OBS_IDX_EC, X1_EC = [0, 2, 5, 9, 10], np.array([0.5, 30, 2000, 10, 10])                         # :151,156


def batch_time_marks(t_cycle=T_CYCLE, t_ratio=T_RATIO, t_delta=OS_T_DELTA):
    """module_batch_time.batch_time (module_batch_time.py:3-116): first and last output stamp of each phase,
    [(t_memoryK[0], t_memoryK[-1])] for K = 1..8, with the reference's int(float) point counts."""
    marks = []
    t_end = 0
    for k in range(8):
        t_start = t_end if k == 0 else t_end + t_delta
        t_end = t_start + t_cycle * t_ratio[k]
        t_save = np.linspace(t_start, t_end, int((t_end - t_start) / (t_delta * 10)))
        mem = [t_save[0]]
        for i in range(len(t_save) - 1):
            t_range = np.linspace(t_save[i], t_save[i + 1], int((t_save[i + 1] - t_save[i]) / t_delta))
            mem.extend(t_range[1:])
        marks.append((float(mem[0]), float(mem[-1])))
    return marks


def reward_os(x, n_pts, span, kla_hist, ec_hist):
    """module_reward_EQIOCI.sbr_reward (module_reward_EQIOCI.py:4-115).  kla_hist / ec_hist are the reference's
    `Kla` / `EC` lists; note Kla holds ONE entry per interval, so Kla[-L:-1] sums the previous L-1 intervals."""
    EQI, _ = effluent_quality(x)
    t_delta = 0.002 / 24
    AE = 8 / (span * 1.8 * 1000) * (1.32 * sum(kla_hist[-n_pts:-1]) * t_delta)
    ECo = EC_CONC * sum(ec_hist[-n_pts:-1]) * t_delta / (span * 1000)
    return (1 - ((EQI / 10) ** 2 + (AE + ECo) ** 2)) / 473


def _clip1(v):
    return 1 if v > 1 else (-1 if v < -1 else v)


def _obs_os(t, x, x_first, x_last):
    """obs_DO / obs_EC / state of SbrOS.step (gym_SBR_oneshot.py:1015-1112)."""
    state = np.concatenate([[t], x]) / X1_STATE
    full = np.concatenate([[t], x])
    d = lambda i, s: _clip1((x_last[i] - x_first[i]) / s)
    # note: index 0 of obs_idx_* is time; other indices address x directly (x[i], not [t,x][i])
    o_do = np.array([t] + [x[i] for i in OBS_IDX_DO[1:]]) / X1_DO
    o_ec = np.array([t] + [x[i] for i in OBS_IDX_EC[1:]]) / X1_EC
    o_do = np.append(o_do, [d(5, 4000), d(6, 500), d(8, 8), d(10, 50)])
    o_ec = np.append(o_ec, [d(2, 50), d(5, 4000), d(9, 50), d(10, 50)])
    del full
    return o_do, o_ec, state


class SbrOsOracle(object):
    """Restatement of the reference's SbrOS env (gym_SBR_oneshot.py:98-2644) with explicit per-env state instead
    of module globals.  reset(influent_mixed) -> (obs_DO[9], obs_EC[9]); step(action[2]) ->
    ((obs_DO, obs_EC), state[15], reward, done)."""

    def __init__(self, ode_kw=None):
        self.ode_kw = ode_kw or {}
        self.marks = batch_time_marks()
        self.n_rhs = 0

    # -- controllers -------------------------------------------------------------------------------------
    def _pid_do(self, sp, t_start, aerobic):
        """DO-PID with incremental bias Kla[-1]; PID dt = 0.002/24, not the control interval
        (gym_SBR_oneshot.py:1889-1909).  In anoxic intervals the PID state advances but Kla := 0 (:1990)."""
        P = OS_PID
        e = sp - self.So[-1]
        if t_start > 0:
            dcv = (self.So[-1] - self.So[-2]) / OS_DT
            self.ie_DO = self.ie_DO + e * OS_DT
        else:
            dcv = 0
            self.ie_DO = 0
        if aerobic:
            kla = P['Kc_DO'] * e + P['Kc_DO'] / P['tauI_DO'] * self.ie_DO + P['Kc_DO'] * P['tauD_DO'] * dcv + self.Kla[-1]
        else:
            kla = 0
        if kla > P['kla_hi']:
            kla = P['kla_hi']
            self.ie_DO = self.ie_DO - e * OS_DT
        if kla < P['kla_lo']:
            kla = P['kla_lo']
            self.ie_DO = self.ie_DO - e * OS_DT
        self.Kla.append(kla)
        return kla

    def _pid_ec(self, sp, t_start, dosing, reversed_sign=True):
        """NO3-PID: error = Sno - sp (sign reversed), output EC in [0, 0.0005] (gym_SBR_oneshot.py:1917-1948);
        outside anoxic react intervals EC := 0 while the integral still advances (:1937)."""
        P = OS_PID
        e = (self.Sno[-1] - sp) if reversed_sign else (sp - self.Sno[-1])
        if t_start > 0:
            dcv = (self.Sno[-1] - self.Sno[-2]) / OS_DT
            self.ie_EC = self.ie_EC + e * OS_DT
        else:
            dcv = 0
            self.ie_EC = 0
        if dosing:
            ec = P['Kc_EC'] * e + P['Kc_EC'] / P['tauI_EC'] * self.ie_EC + P['Kc_EC'] * P['tauD_EC'] * dcv + self.EC[-1]
        else:
            ec = 0
        if ec < P['ec_lo']:
            ec = P['ec_lo']
            self.ie_EC = self.ie_EC - e * OS_DT
        elif ec > P['ec_hi']:
            ec = P['ec_hi']
            self.ie_EC = self.ie_EC - e * OS_DT
        self.EC.append(ec)
        return ec

    # -- reset: fill phase -------------------------------------------------------------------------------
    def reset(self, influent_mixed, x0=X0_INIT):
        """SbrOS.reset + Sim_filling (gym_SBR_oneshot.py:168-438, 1585-1654)."""
        self.infl = list(influent_mixed)
        x0 = np.array(x0, dtype=float)
        self.So, self.Ss, self.Sno = [x0[8]], [x0[2]], [x0[9]]
        self.Kla, self.EC = [0], [0]
        self.ie_DO = self.ie_EC = 0
        self.u_DO, self.u_EC = 0, 15
        self.infl[0] = QIN / self.marks[0][1]                                     # :287
        t_end = 0 + T_RATIO[0] * 0.5
        t_range = np.linspace(0, t_end, int((t_end - 0) / OS_DT))
        kla = self._pid_do(0, 0, aerobic=True)                                     # sp 0 -> clamps to 0
        # the fill-phase EC controller: error = 0 - Sno, EC := 0, clamp order upper-then-lower (:1620-1645)
        self.ie_EC = 0
        self.EC.append(0)
        x_out = odeint(rhs_fill, x0, t_range, args=(kla, self.infl), **self.ode_kw)
        self.So.append(x_out[-1][8])
        self.Ss.append(x_out[-1][2])
        self.Sno.append(x_out[-1][2])                                              # sic: Ss stored as Sno (:1652)
        rep = int(len(x_out) / len(self.Kla))
        self.Kla = self.Kla * rep                                                  # :322
        self.EC = self.EC * int(len(x_out) / len(self.EC))                         # :323
        self.t = float(t_range[-1])
        self.x = x_out[-1]
        self.n_pts_fill = len(t_range)
        x = self.x
        mix = lambda i: (QIN * self.infl[i] + x[i] * IV) / (QIN + IV)              # :347-364
        o_do = np.array([self.t] + [mix(i) for i in OBS_IDX_DO[1:]]) / X1_DO
        o_ec = np.array([self.t] + [mix(i) for i in OBS_IDX_EC[1:]]) / X1_EC
        d = lambda i, s: _clip1((x_out[-1][i] - x_out[0][i]) / s)
        o_do = np.append(o_do, [d(5, 4000), d(6, 500), d(8, 8), d(10, 50)])
        o_ec = np.append(o_ec, [d(2, 50), d(5, 4000), d(9, 50), d(10, 50)])
        self.schedule_log = []
        return o_do, o_ec

    # -- one control interval ----------------------------------------------------------------------------
    def _interval(self, aerobic):
        """run_aero_step / run_anaero_step + Sim_(an)aero_rxn (gym_SBR_oneshot.py:1331-1419, 1877-2051)."""
        t_start = self.t
        t_end = self.t + OS_T_DELTA
        t_range = np.linspace(t_start, t_end, int((t_end - t_start) / OS_DT))
        kla = self._pid_do(self.u_DO if aerobic else 0, t_range[0], aerobic)
        ec = self._pid_ec(self.u_EC, t_range[0], dosing=not aerobic)
        x_out = odeint(rhs_react_ec, self.x, t_range, args=(kla, ec, EC_CONC), **self.ode_kw)
        for _ in range(len(t_range) - 2):
            self.EC.append(self.EC[-1])
        self.So.append(x_out[-1][8])
        self.Ss.append(x_out[-1][2])
        self.Sno.append(x_out[-1][9])
        self.t = float(t_range[-1])
        self.x_first, self.x = x_out[0], x_out[-1]
        self.schedule_log.append((float(t_start), int(aerobic), len(t_range), float(t_range[-1] - t_range[0])))
        return t_range

    def step(self, action):
        """SbrOS.step (gym_SBR_oneshot.py:843-1273): four NON-exclusive `if`s on the running time, so a step that
        crosses a phase boundary runs two intervals."""
        m = self.marks
        tm3_0, tm3_1, tm4_1, tm5_1 = m[2][0], m[2][1], m[3][1], m[4][1]
        t_range = None
        if self.t < tm3_0:
            self.u_EC, self.u_DO = min(max(action[1], 0), 15), 0
            t_range = self._interval(aerobic=False)
        if (self.t >= tm3_0) and (self.t <= tm3_1):
            self.u_DO, self.u_EC = min(max(action[0], 0), 8), 0
            t_range = self._interval(aerobic=True)
        if (self.t > tm3_1) and (self.t <= tm4_1):
            self.u_EC, self.u_DO = min(max(action[1], 0), 15), 0
            t_range = self._interval(aerobic=False)
        if self.t > tm4_1:
            self.u_DO, self.u_EC = min(max(action[0], 0), 8), 0
            t_range = self._interval(aerobic=True)
        reward = reward_os(self.x, len(t_range), t_range[-1] - t_range[0], self.Kla, self.EC)
        o_do, o_ec, state = _obs_os(self.t, self.x, self.x_first, self.x)
        done = False
        if self.t >= tm5_1:
            done = True
            x_react_end = self.x
            # settle + draw (Sim_Settling_Drawing, :2264-2420): same algebra as Path A on a 49-point grid
            t_set = np.linspace(self.t, self.t + T_RATIO[5] * T_CYCLE, int((T_RATIO[5] * T_CYCLE) / OS_T_DELTA))
            Xf = 0.75 * (self.x[3] + self.x[4] + self.x[5] + self.x[6] + self.x[7])
            z = self.x[0] / ((1.25 / 2) ** 2)
            sX = odeint(settler_rhs, [Xf] * 10, t_set, args=(z, Xf), **self.ode_kw)[-1]
            x_n, self.Qw, _, _, self.draw_status = draw(self.x, sX, Xf)
            t_draw = np.linspace(t_set[-1], t_set[-1] + T_RATIO[6] * T_CYCLE, int((T_RATIO[6] * T_CYCLE) / OS_T_DELTA))
            for _ in range(len(t_set) + len(t_draw) - 2):
                self.EC.append(0)
            self.So += [self.x[8]] * len(t_set) + [x_n[8]] * (len(t_draw) - 1)     # :2415-2416
            # idle (Sim_idle, :2554-2597): one PID update with the last DO set-point, ONE solve to t_cycle
            t_idle = np.linspace(t_draw[-1], T_CYCLE, int((T_CYCLE - t_draw[-1]) / OS_DT))
            kla = self._pid_do(self.u_DO, t_idle[0], aerobic=True)
            x_idle = odeint(rhs_react, x_n, t_idle, args=(kla,), **self.ode_kw)
            self.So.append(x_idle[-1][8])
            self.Ss.append(x_idle[-1][2])
            self.Sno.append(x_idle[-1][9])
            self.t = float(t_idle[-1])
            self.x = x_idle[-1]
            self.idle_pts = len(t_idle)
            o_do, o_ec, state = _obs_os(self.t, self.x, x_react_end, self.x)
        return (o_do, o_ec), state, reward, done


# ----------------------------------------------------------------------------------------------------------
# SBR-v4 (SbrEnv4): interval-per-step env with the FILL phase stepped inside step(), 1-D delta-set-point action
# ----------------------------------------------------------------------------------------------------------
X1_V4 = np.array([1.32000000e+00, 3.00000000e+01, 3.81606587e+01, 6.94658685e+02, 1.07772100e+02, 1.22613841e+03,
                  7.88460027e+01, 2.57616136e+02, 1.01108024e+00, 6.24510635e+00, 1.78877937e+01, 3.95743344e+00,
                  5.70432163e+00, 5.50185509e+00])                                   # gym_SBR_env4.py:91
V4_PID = dict(Kc=5, tauI=0.00035, tauD=0.005, lo=0, hi=240)                           # gym_SBR_env4.py:61-69


def reward_v4(kla_list, batch_type, Qin, Qw, eff, so_sat=SO_SAT):
    """module_reward_continuous.sbr_reward (module_reward_continuous.py:4-65)."""
    t_delta = 0.002 / 24
    r_snh = 0
    if batch_type == 0:
        PE = 0.004 * Qin
        AE_dT = 1.32 * kla_list[-1] * t_delta
    elif batch_type == 1:
        AE_dT = 1.32 * kla_list[-1] * t_delta
        PE = 0
    else:
        PE = 0.05 * Qw + 0.004 * eff[0]
        AE_dT = 1.32 * sum(kla_list) * t_delta
        r_snh = 0 if eff[3] < 4 else -246
    AE = so_sat / (1.8 * 1000) * AE_dT
    return (0.5 - (AE + PE)) + r_snh


class SbrEnv4Oracle(object):
    """Restatement of the reference's SbrEnv4 (gym_SBR_env4.py:71-1294), run with numpy < 1.18 linspace semantics
    (float `num` truncated with int(); see oracle/make_golden_v4.py for the disclosure).
    reset(influent_mixed) -> state[14]; step(action) -> (state[14], reward, done)."""

    def __init__(self, ode_kw=None):
        self.ode_kw = ode_kw or {}
        self.marks = batch_time_marks()

    def reset(self, influent_mixed, x0=X0_INIT):
        self.infl = list(influent_mixed)
        self.x0 = np.array(x0, dtype=float)
        self.t = 0
        self.batch_type = 0
        self.dcv, self.ie, self.e, self.So, self.Kla = [], [], [], [], []
        x_2 = np.zeros(14)
        x_2[0] = QIN + IV
        for i in range(1, 14):
            x_2[i] = (QIN * self.infl[i] + self.x0[i] * IV) / (QIN + IV)               # :185-190
        self.infl[0] = QIN / self.marks[0][1]                                          # :193
        self.x = self.x0
        self.Qw = 0
        self.eff = []
        return x_2 / X1_V4

    def _pid(self, sp, t_start):
        """Sim_filling / Sim_rxn / Sim_idle controller block (gym_SBR_env4.py:497-524, 667-697, 1202-1234):
        incremental bias Kla[-1], PID dt = 0.002/24, two independent clamp checks with anti-windup."""
        P = V4_PID
        self.e.append(sp - self.So[-1])
        if t_start > 0:
            self.dcv.append((self.So[-1] - self.So[-2]) / OS_DT)
            self.ie.append(self.ie[-1] + self.e[-1] * OS_DT)
        else:
            self.dcv.append(0)
            self.ie.append(0)
        kla = P['Kc'] * self.e[-1] + P['Kc'] / P['tauI'] * self.ie[-1] + P['Kc'] * P['tauD'] * self.dcv[-1] + self.Kla[-1]
        self.Kla.append(kla)
        if self.Kla[-1] > P['hi']:
            self.Kla[-1] = P['hi']
            self.ie[-1] = self.ie[-1] - self.e[-1] * OS_DT
        if self.Kla[-1] < P['lo']:
            self.Kla[-1] = P['lo']
            self.ie[-1] = self.ie[-1] - self.e[-1] * OS_DT
        return self.Kla[-1]

    def step(self, action):
        m = self.marks
        if self.t == 0:
            self.u = 0
        self.u = self.u + action
        self.u = 0 if self.u < 0 else (8 if self.u > 8 else self.u)                    # :213-218
        x_in = self.x0 if self.t == 0 else self.x
        t = self.t
        if m[0][0] <= t < m[0][1]:
            bt = 0
        elif t < m[4][1]:
            bt = 1
        else:
            bt = 2
        if t == 0:
            self.So.append(x_in[8])
            self.Kla.append(0)
        t_start, t_end = t, t + OS_T_DELTA
        t_range = np.linspace(t_start, t_end, int((t_end - t_start) / OS_DT))
        self.n_pts = len(t_range)
        if bt == 0:
            kla = self._pid(self.u, t_range[0])
            x_out = odeint(rhs_fill, x_in, t_range, args=(kla, self.infl), **self.ode_kw)
            self.So.append(x_out[-1][8])
            self.Qw, self.eff = 0, []
        elif bt == 1:
            kla = self._pid(self.u, t_range[0])
            x_out = odeint(rhs_react, x_in, t_range, args=(kla,), **self.ode_kw)
            self.So.append(x_out[-1][8])
            self.Qw, self.eff = 0, []
        else:
            # Sim_Settling_Drawing (:919-1070; called with dt := t_delta) then Sim_idle (:1202-1242)
            t_set = np.linspace(t, t + T_RATIO[5] * T_CYCLE, int((T_RATIO[5] * T_CYCLE) / OS_T_DELTA))
            Xf = 0.75 * (x_in[3] + x_in[4] + x_in[5] + x_in[6] + x_in[7])
            z = x_in[0] / ((1.25 / 2) ** 2)
            sX = odeint(settler_rhs, [Xf] * 10, t_set, args=(z, Xf), **self.ode_kw)[-1]
            x_n, self.Qw, _, self.eff, self.draw_status = draw(x_in, sX, Xf)
            t_draw = np.linspace(t_set[-1], t_set[-1] + T_RATIO[6] * T_CYCLE, int((T_RATIO[6] * T_CYCLE) / OS_T_DELTA))
            self.So += [x_in[8]] * len(t_set) + [x_n[8]] * (len(t_draw) - 1)
            t_idle = np.linspace(t_draw[-1], T_CYCLE, int((T_CYCLE - t_draw[-1]) / OS_DT))
            kla = self._pid(self.u, t_draw[-1])
            x_out = odeint(rhs_react, x_n, t_idle, args=(kla,), **self.ode_kw)
            self.So.append(x_out[-1][8])
            self.idle_pts = len(t_idle)
            t_end = float(t_idle[-1])
        self.t = t_end
        self.batch_type = bt
        self.x = x_out[-1]
        reward = reward_v4(self.Kla, bt, QIN, self.Qw, self.eff)
        done = (bt == 2) and (self.t >= T_CYCLE)
        return self.x / X1_V4, reward, done


# ----------------------------------------------------------------------------------------------------------
# SBRCnt-v0/1/2, SBRCntMA-v1, SBROS-v2: gym_SBR_continuous0/1/2.py, gym_SBR_continuous_MA1.py,
# gym_SBR_oneshot1.py.  Pinned by tests/test_oracle_golden_cnt.py to episodes of the unmodified env modules run with the
# repaired reward (oracle/make_golden_cnt.py holds the disclosure).
# ----------------------------------------------------------------------------------------------------------
def batch_time_stamps(t_cycle=T_CYCLE, t_ratio=T_RATIO, t_delta=OS_T_DELTA):
    """module_batch_time.batch_time (module_batch_time.py:3-116): the full stamp list t_memoryK of each phase."""
    out = []
    t_end = 0
    for k in range(8):
        t_start = t_end if k == 0 else t_end + t_delta
        t_end = t_start + t_cycle * t_ratio[k]
        t_save = np.linspace(t_start, t_end, int((t_end - t_start) / (t_delta * 10)))
        mem = [t_save[0]]
        for i in range(len(t_save) - 1):
            t_range = np.linspace(t_save[i], t_save[i + 1], int((t_save[i + 1] - t_save[i]) / t_delta))
            mem.extend(t_range[1:])
        out.append([float(v) for v in mem])
    return out


CNT_KIND = {
    # DO PID (Kc, tauI, tauD); carbon PID (Kc, tauI, tauD); EC_conc; carbon set-point clip; controlled variable index
    "cnt0": dict(do=(10, 0.5, 0.00005), ec=None),                                           # gym_SBR_continuous0.py:80-82
    "cnt1": dict(do=(100, 20, 0), ec=None),                                                 # gym_SBR_continuous1.py:77-79
    "cnt2": dict(do=(100, 20, 0), ec=(1, 20, 0), conc=400000 / 20648.38 * 1.32, u_ec_max=5, cv=2),   # ..2.py:81-94
    "ma1": dict(do=(100, 20, 0), ec=(10, 0.5, 0), conc=4000 / 20648.38 * 1.32, u_ec_max=15, cv=9),   # .._MA1.py:80-93
    "os2": dict(do=(100, 20, 0), ec=(1, 20, 0), conc=400000 / 20648.38 * 1.32, u_ec_max=15, cv=9),   # gym_SBR_oneshot1.py:82-95
}


def reward_cnt(x_out, done, eff):
    """module_reward_continuous1.sbr_reward (module_reward_continuous1.py:5-65) with its three unbound names bound as
    oracle/make_golden_cnt.py documents."""
    so = x_out[8]
    if done:
        return 0 if np.isscalar(eff) else (0 if eff[3] < 4 else -246)
    if so < 1.5:
        return -100
    if 2.5 < so < 3.5:
        return 0
    if 3.5 <= so < 5:
        return -10
    if 5 <= so:
        return -50
    return 10


class SbrCntOracle(object):
    """Restatement of SbrCnt0 / SbrCnt1 / SbrCnt2 / SbrCntMA1 / SbrOS1 with explicit per-env state instead of module
    globals.  reset(influent_mixed) -> obs; step(action) -> (obs, reward, done), for "os2" ((obs_DO, obs_EC), state,
    reward, done)."""

    def __init__(self, kind, ode_kw=None):
        self.kind, self.K = kind, CNT_KIND[kind]
        self.ode_kw = ode_kw or {}
        self.tm = batch_time_stamps()

    # -- controllers (Sim_rxn of every file; e.g. gym_SBR_continuous2.py:900-965) ------------------------------------
    def _pid_do(self, sp, t_start, Kla):
        Kc, tauI, tauD = self.K["do"]
        self.e_DO.append(sp - self.So[-1])
        if t_start > 0:
            self.dcv_DO.append((self.So[-1] - self.So[-2]) / OS_DT)
            self.ie_DO.append(self.ie_DO[-1] + self.e_DO[-1] * OS_DT)
        else:
            self.dcv_DO.append(0)
            self.ie_DO.append(0)
        Kla.append(Kc * self.e_DO[-1] + Kc / tauI * self.ie_DO[-1] + Kc * tauD * self.dcv_DO[-1] + Kla[-1])
        if Kla[-1] > 240:
            Kla[-1] = 240
            self.ie_DO[-1] = self.ie_DO[-1] - self.e_DO[-1] * OS_DT
        if Kla[-1] < 0:
            Kla[-1] = 0
            self.ie_DO[-1] = self.ie_DO[-1] - self.e_DO[-1] * OS_DT
        return Kla[-1]

    def _pid_ec(self, sp, t_start, fill):
        Kc, tauI, tauD = self.K["ec"]
        self.e_EC.append(sp - self.CV[-1])
        if t_start > 0:
            self.dcv_EC.append((self.CV[-1] - self.CV[-2]) / OS_DT)
            self.ie_EC.append(self.ie_EC[-1] + self.e_EC[-1] * OS_DT)
        else:
            self.dcv_EC.append(0)
            self.ie_EC.append(0)
        ec = Kc * self.e_EC[-1] + Kc / tauI * self.ie_EC[-1] + Kc * tauD * self.dcv_EC[-1] + self.EC[-1]
        if fill:                                   # Sim_filling clamps to EC_control_par[4..5] = [0, 5] (:750-760)
            if ec > 5:
                ec = 5
                self.ie_EC[-1] = self.ie_EC[-1] - self.e_EC[-1] * OS_DT
            if ec < 0:
                ec = 0
                self.ie_EC[-1] = self.ie_EC[-1] - self.e_EC[-1] * OS_DT
            self.EC.append(ec)
        else:                                      # Sim_rxn: lower clamp only, ten copies appended (:958-965)
            if ec < 0:
                ec = 0
                self.ie_EC[-1] = self.ie_EC[-1] - self.e_EC[-1] * OS_DT
            self.EC.extend([ec] * 10)
        return ec

    def _sim_rxn(self, x, t_range, u_do, Kla, u_ec):
        kla = self._pid_do(u_do, t_range[0], Kla)
        if self.K["ec"] is None:
            x_out = odeint(rhs_react, x, t_range, args=(kla,), **self.ode_kw)
        else:
            ec = self._pid_ec(u_ec, t_range[0], False)
            x_out = odeint(rhs_react_ec, x, t_range, args=(kla, ec, self.K["conc"]), **self.ode_kw)
            self.CV.append(x_out[-1][self.K["cv"]])
        self.So.append(x_out[-1][8])
        return x_out

    # -- reset (gym_SBR_continuous0.py:120-235 and the same function of the other four files) -----------------------
    def reset(self, influent_mixed, x0=X0_INIT):
        k = self.kind
        self.infl = list(influent_mixed)
        x0 = np.array(x0, dtype=float)
        self.t, self.u_do, self.u_ec = 0, 0, (2 if self.K["ec"] is not None else 0)
        self.dcv_DO, self.ie_DO, self.e_DO, self.dcv_EC, self.ie_EC, self.e_EC = [], [], [], [], [], []
        self.So, self.Kla, self.EC = [x0[8]], [0], [0]
        self.CV = [x0[self.K["cv"]]] if self.K["ec"] is not None else []
        self.infl[0] = QIN / self.tm[0][-1]
        t_range = np.linspace(0, 0 + T_RATIO[0] * 0.5, int((0 + T_RATIO[0] * 0.5 - 0) / OS_DT))
        kla = self._pid_do(0, 0, self.Kla)
        if self.K["ec"] is not None:
            self._pid_ec(0, 0, True)               # EC = 0: the fill RHS's in-place dilution is then a no-op
        x_out = odeint(rhs_fill, x0, t_range, args=(kla, self.infl), **self.ode_kw)
        self.So.append(x_out[-1][8])
        if self.K["ec"] is not None:
            self.CV.append(x_out[-1][2])           # Ss list; SbrCntMA1 / SbrOS1 store Ss in their Sno list too (:788)
        self.t = float(t_range[-1])
        self.x = x_out[-1]
        self.x_out = x_out
        mix = lambda i: (QIN * self.infl[i] + self.x[i] * IV) / (QIN + IV)
        if k == "cnt0":
            return np.array([[self.t] + [mix(i) for i in (1, 5, 6, 8, 9, 10)]]) / np.array([0.5, 30, 2599., 168., 2., 13., 0.005])
        if k == "os2":
            o_do, o_ec, _ = _obs_os(self.t, np.array([0.0] + [mix(i) for i in range(1, 14)]), x_out[0], x_out[-1])
            return o_do, o_ec
        return self._obs5(np.array([0.0] + [mix(i) for i in range(1, 14)]), x_out)

    def _obs5(self, x, x_out):
        d_so, d_snh = (x_out[-1][8] - x_out[0][8]) / 8, (x_out[-1][10] - x_out[0][10]) / 20
        return np.array([self.t / 0.5, x[8] / 8., x[10] / 30, _clip1(d_so), _clip1(d_snh)])

    def _run_step(self, u_do, u_ec):
        """run_step (gym_SBR_continuous0.py:326-358): one control interval from the running time."""
        t_range = np.linspace(self.t, self.t + OS_T_DELTA, int(((self.t + OS_T_DELTA) - self.t) / OS_DT))
        self.x_out = self._sim_rxn(self.x, t_range, u_do, self.Kla, u_ec)
        self.x = self.x_out[-1]
        self.t = float(t_range[-1])

    def _whole_phase(self, stamps, u_ec):
        """SbrCnt1 / SbrCnt2: a whole anoxic phase in one Sim_rxn call at DO set-point 0 with a fresh `[0]` KLa list
        (gym_SBR_continuous1.py:281-295, 331-344); SbrCnt2's tuple unpacking rebinds the global Kla to that list (:351)."""
        Kla = [0]
        x_out1 = self._sim_rxn(self.x, stamps, 0, Kla, u_ec)
        if self.kind == "cnt2":
            self.Kla = Kla
        self.x = x_out1[-1]
        self.t = float(stamps[-1])

    def _terminal(self):
        """Sim_Settling_Drawing + Sim_idle (gym_SBR_continuous0.py:913-1236)."""
        t, x_in = self.t, self.x
        t_set = np.linspace(t, t + T_RATIO[5] * T_CYCLE, int((T_RATIO[5] * T_CYCLE) / OS_T_DELTA))
        Xf = 0.75 * (x_in[3] + x_in[4] + x_in[5] + x_in[6] + x_in[7])
        z = x_in[0] / ((1.25 / 2) ** 2)
        sX = odeint(settler_rhs, [Xf] * 10, t_set, args=(z, Xf), **self.ode_kw)[-1]
        x_n, self.Qw, _, eff, self.draw_status = draw(x_in, sX, Xf)
        t_draw = np.linspace(t_set[-1], t_set[-1] + T_RATIO[6] * T_CYCLE, int((T_RATIO[6] * T_CYCLE) / OS_T_DELTA))
        self.So += [x_in[8]] * len(t_set) + [x_n[8]] * (len(t_draw) - 1)
        t_idle = np.linspace(t_draw[-1], T_CYCLE, int((T_CYCLE - t_draw[-1]) / OS_DT))
        kla = self._pid_do(self.u_do, t_draw[-1], self.Kla)
        x_idle = odeint(rhs_react, x_n, t_idle, args=(kla,), **self.ode_kw)
        self.So.append(x_idle[-1][8])
        self.t = float(t_idle[-1])
        self.x = x_idle[-1]
        return x_in, x_n, x_idle[-1], eff

    def step(self, action):
        k, tm = self.kind, self.tm
        a = np.asarray(action, dtype=float).reshape(-1)
        clip = lambda v, hi: 0 if v < 0 else (hi if v > hi else v)
        if k in ("cnt0", "cnt1", "cnt2"):
            if k == "cnt1" and self.t < tm[1][0]:
                self._whole_phase(tm[1], 0)
            self.u_do = clip(self.u_do + a[0], 8)
            if k == "cnt2":
                self.u_ec = clip(self.u_ec, 5)
                if self.t < tm[1][0]:
                    self.u_ec = clip(self.u_ec + a[0], 5)
                    self._whole_phase(tm[1], self.u_ec)
            self._run_step(self.u_do, self.u_ec)
            x_obs, x_out = self.x, self.x_out
            if k != "cnt0" and tm[2][-1] <= self.t < tm[3][-1]:
                self._whole_phase(tm[3], self.u_ec)
        else:
            for p_ in range(4):
                t = self.t
                cond = (t < tm[2][0], tm[2][0] <= t <= tm[2][-1], tm[2][-1] < t <= tm[3][-1], t > tm[3][-1])[p_]
                if not cond:
                    continue
                if p_ % 2 == 0:
                    self.u_ec = clip(self.u_ec + a[0] if k == "ma1" else a[1], 15)
                    self.u_do = 0
                else:
                    self.u_do = clip(self.u_do + a[0] if k == "ma1" else a[0], 8)
                    self.u_ec = 0
                self._run_step(self.u_do, self.u_ec)
            x_obs, x_out = self.x, self.x_out
        reward = reward_cnt(x_out[-1], False, 0)
        if k == "cnt0":
            obs = np.array([[self.t] + [x_obs[i] for i in (1, 5, 6, 8, 9, 10)]]) / np.array([0.5, 30, 2599., 168., 2., 13., 0.005])
        elif k == "os2":
            obs = _obs_os(self.t, x_obs, x_out[0], x_out[-1])
        else:
            obs = self._obs5(x_obs, x_out)
        done = self.t >= tm[4][-1]
        if done:
            x_pre, x_draw, x_idle, eff = self._terminal()
            if k == "cnt0":
                reward = reward_cnt(x_idle, True, eff)
                obs = np.array([[self.t] + [x_idle[i] for i in (1, 5, 6, 8, 9, 10)]]) / np.array([0.5, 30, 2599., 168., 2., 13., 0.005])
            elif k == "os2":
                obs = _obs_os(self.t, x_draw, x_pre, x_idle)
        if k == "os2":
            return (obs[0], obs[1]), obs[2], reward, done
        return obs, reward, done
