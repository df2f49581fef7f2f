"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (gym_sbr2_b200/); tests/ and bench.py's CPU legs may.

CPU restatement of the reference's batch-to-batch (iterative-learning) feed-forward KLa path, the research feature of
`SBR-v0`:

  * `module_batch_PID.batch_PID`                       (module_batch_PID.py:7-275)   -> ilc_weights / ilc_e_batch / IlcMemory.update
  * `sub_phases_batchPID_fbPID.filling/rxn.sim_rxn`    (sub_phases_batchPID_fbPID.py:139-253, 388-500) -> ilc_phase(seed_bias=False)
  * `sub_phases_PID_on.filling/rxn.sim_rxn`            (sub_phases_PID_on.py:178-271, 406-500)         -> ilc_phase(seed_bias=True)
  * `SBR_model_batchPID_fbPID.run` / `SBR_model_PID_on.run` (phase sequencing, closed-form Qw, drawing)  -> ilc_cycle
  * `drawing.sim_drawing` of those model files          (sub_phases_batchPID_fbPID.py:781-812)          -> draw_fixed_qw

Pinned by tests/test_oracle_golden_ilc.py to outputs of the reference's own functions (oracle/make_golden_ilc.py; the
float-`num` np.linspace calls of sub_phases_batchPID_fbPID.py need numpy < 1.18 semantics, disclosed there).

What is NOT restated: `SbrEnv.step`'s reward (gym_SBR_env0.py:203 calls the ten-argument module_reward.sbr_reward with
seven arguments -> TypeError; no version of that function with this signature exists in the reference) and
`buffer_tank2` (the influent of this path is an input here).
"""
import numpy as np
from scipy.integrate import odeint

from . import sbr_oracle as O

DT = O.DT
# gym_SBR_env0.py:89: Kc, taui, delt, So_set, Kla_min, Kla_max, DKla_max, So_low, So_high, tauD, So_sat
PID_ILC = dict(Kc=0.5 / 1.18, tauI=0.0015, dt=0.05, lo=0.0, hi=240.0, tauD=0.005)
# gym_SBR_env0.py:93: (tau_w, theta_w) of phases 1..8
PAR_BATCH_PID = [0.002018, 0.003643, 0.004036, 0, 0.01875, 0.0004671, 0.01564, 0.003643, 0.001028, 0, 0, 0, 0, 0,
                 0.003027, 0.003643]
KC_B, TAUI_B, TAUC_B = 1 / 1.18, 0.25, 0.1                                  # module_batch_PID.py:15-17
X0_ILC = [0.66, 30.0, 0.5601630529230822, 1762.3890076468106, 30.97046860269441, 2628.6551849696393,
          188.71238190722482, 780.479571994941, 6.83620016588177, 14.575400491942467, 0.00872090237410032,
          0.36940333660700486, 1.896711744868243, 3.705237172170034]       # gym_SBR_env0.py:69-71
WV_ILC, IV_ILC = 1.32, 0.66                                                 # gym_SBR_env0.py:40-41
FILL_FLOW_ILC = 31.4285                                                     # gym_SBR_env0.py:76,193
BIOMASS_SETPOINT_ILC = 5400                                                 # SBR_model_batchPID_fbPID.py:283
PHASES = (0, 1, 2, 3, 4, 7)                                                 # the six PID-controlled phases (1,2,3,4,5,8)


# ----------------------------------------------------------------------------------------------------------
# batch-to-batch controller (module_batch_PID.py)
# ----------------------------------------------------------------------------------------------------------
def ilc_weights(t_memories, par=PAR_BATCH_PID, t_delta=DT):
    """Per phase: (w, tp).  w(t) = ((t - theta)/tau_a) exp(-(t - theta_b)/tau_b) for t > theta, else 0, with the
    reference's own mix-ups kept: phases 2, 3, 4 divide by tau_w1 in the linear factor (module_batch_PID.py:66,94,122)
    and phase 3's exponent uses theta_w1 and tau_w1 (:94).  tp = int(3 tau_w / t_delta) samples (:29)."""
    out = []
    for j, k in enumerate(PHASES):
        tau, theta = par[2 * k], par[2 * k + 1]
        tau1, theta1 = par[0], par[1]
        t = np.array(t_memories[k])
        tp = int(tau * 3 / t_delta)
        idx = np.where(t > theta)[0][0]
        ts = t[idx:]
        if j == 0 or j >= 4:                      # phases 1, 5, 8: as intended
            w2 = ((ts - theta) / tau) * np.exp(-((ts - theta) / tau))
        elif j == 2:                              # phase 3
            w2 = ((ts - theta) / tau1) * np.exp(-((ts - theta1) / tau1))
        else:                                     # phases 2, 4
            w2 = ((ts - theta) / tau1) * np.exp(-((ts - theta) / tau))
        out.append((np.concatenate([np.zeros(idx), w2]), tp))
    return out


def ilc_e_batch(sp_mem, so_mem, w, tp, t_delta=DT):
    """E_batch(t) = sum_{j in [t, t+tp)} (sp_j - So_j) w_j dt / sum w_j dt, window cut at the end of the phase
    (module_batch_PID.py:37-52).  Python's sequential sum, as the reference."""
    n = len(w)
    sp_mem, so_mem = np.array(sp_mem, dtype=float), np.array(so_mem, dtype=float)
    E = np.zeros(n)
    for t in range(n):
        hi = t + tp if t + tp <= n else n
        E[t] = np.divide(sum(np.multiply(sp_mem[t:hi] - so_mem[t:hi], w[t:hi]) * t_delta), sum(w[t:hi] * t_delta))
    return E


class IlcMemory(object):
    """memory_e_batch_k / u_batch_k of gym_SBR_env0.py:58-70 with the update of module_batch_PID.py:214-270:
    u = Kc (E_last + (1/tauI) sum_over_all_cycles E + tauD (E_last - E_previous)); both start as one row of zeros."""

    def __init__(self, lengths):
        self.e_rows = [np.zeros((1, n)) for n in lengths]
        self.u = [np.zeros(n) for n in lengths]

    def update(self, E_list):
        for j, E in enumerate(E_list):
            self.e_rows[j] = np.append(self.e_rows[j], E[None, :], axis=0)
            m = self.e_rows[j]
            ie = m.sum(axis=0)
            de = m[-1] - m[-2]
            self.u[j] = KC_B * m[-1] + KC_B * (1 / TAUI_B) * ie + KC_B * TAUC_B * de
        return self.u


# ----------------------------------------------------------------------------------------------------------
# one PID-controlled phase with feed-forward KLa
# ----------------------------------------------------------------------------------------------------------
def ilc_phase(x, t_start, t_end, sp, rhs, rhs_args=(), kla_memory=None, u_batch=None, kla_seed=None, pid=PID_ILC,
              ode_kw=None):
    """sim_rxn of sub_phases_batchPID_fbPID.py (kla_seed None: feed-forward from kla_memory + u_batch, feedback bias starts
    at 0, :173-232) or of sub_phases_PID_on.py (kla_seed given: no feed-forward, bias seeded with the incoming KLa, :218).
    Returns (x_end, So_memory, Kla_memory) with one entry per output sample (the phase start + len(t_range)-1 per interval).
    """
    ode_kw = ode_kw or {}
    t_save2 = np.linspace(t_start, t_end, int((t_end - t_start) / (DT * 10)))
    n = len(t_save2) - 1
    Kc, tauI, tauD, dtc = pid['Kc'], pid['tauI'], pid['tauD'], pid['dt']
    So = np.zeros(n); e = np.zeros(n); ie = np.zeros(n); dcv = np.zeros(n); Kla = np.zeros(n)
    Kla_memory = []
    ranges = [np.linspace(t_save2[i], t_save2[i + 1], int((t_save2[i + 1] - t_save2[i]) / DT)) for i in range(n)]
    if kla_seed is None:
        Kla_memory.append(kla_memory[0])
        for i in range(n):
            for ii in range(len(ranges[i]) - 1):
                v = u_batch[9 * i + ii + 1] + kla_memory[9 * i + ii + 1]
                if v > pid['hi']:
                    v = pid['hi']
                if v < pid['lo']:
                    v = pid['lo']
                Kla_memory.append(v)
    else:
        Kla[0] = kla_seed
        Kla_memory.append(kla_seed)
    x = np.asarray(x, dtype=float)
    So[0] = x[8]
    So_memory = [x[8]]
    for i in range(n):
        e[i] = sp - So[i]
        if i >= 1:
            dcv[i] = (So[i] - So[i - 1]) / dtc
            ie[i] = ie[i - 1] + e[i] * dtc
        Kla[i] = Kc * e[i] + Kc / tauI * ie[i] + Kc * tauD * dcv[i] + Kla[0]
        if Kla[i] > pid['hi']:
            Kla[i] = pid['hi']
            ie[i] = ie[i] - e[i] * dtc
        if Kla[i] < pid['lo']:
            Kla[i] = pid['lo']
            ie[i] = ie[i] - e[i] * dtc
        KLA = Kla[i] + Kla_memory[9 * i + 1] if kla_seed is None else Kla[i]
        soln = odeint(rhs, x, ranges[i], args=(KLA,) + tuple(rhs_args), **ode_kw)
        for ii in range(len(ranges[i]) - 1):
            So_memory.append(soln[ii + 1][8])
            if kla_seed is not None:
                Kla_memory.append(Kla[i])
        if i < n - 1:
            So[i + 1] = soln[-1][8]
        x = soln[-1]
    return x, np.array(So_memory), np.array(Kla_memory, dtype=float)


def draw_fixed_qw(x, sX, Xf, Qeff, Qw):
    """drawing.sim_drawing of the PID_on / batchPID model files (sub_phases_batchPID_fbPID.py:784-809): Qeff and Qw
    are given, the particulates are rescaled to the mixed residual solids."""
    x = np.array(x, dtype=float)
    init_V = x[0]
    V = init_V - Qeff - Qw
    sX2 = (sum(sX) * init_V / 10 - Qw * sX[0] - Qeff * sX[-1]) / V
    out = x.copy()
    out[0] = V
    for i in (4, 7, 3, 5, 6):
        out[i] = (0.75 * x[i] / Xf) * sX2
    return out


def ilc_cycle(x0, influent, setpoints, kla_memory=None, u_batch=None, ode_kw=None):
    """SBR_model_batchPID_fbPID.run (kla_memory / u_batch given, SBR_model_batchPID_fbPID.py:8-345) or
    SBR_model_PID_on.run (both None: the uncontrolled-feed-forward cycle 0 whose KLa profile is the feed-forward base,
    gym_SBR_env0.py:105-106; phase 1 starts from Kla_max, each phase from the previous one's last KLa, idle from phase 5's).
    setpoints: DO set-points of the 8 phases.  influent[0] = fill flow.
    Returns dict(x_last, So_memory[6], Kla_memory[6], Qeff, Qw)."""
    t_ph = [O.T_CYCLE * r for r in O.T_RATIO]
    Qin = WV_ILC - IV_ILC
    qin = Qin / t_ph[0]
    ff = kla_memory is not None
    x = np.array(x0, dtype=float)
    So_mem, Kla_mem = [], []
    t_end = 0
    kla = PID_ILC['hi']                                             # SBR_model_PID_on.py: kla0 = DO_control_par[5]
    for j, k in enumerate(PHASES[:5]):
        t_start = t_end if k == 0 else t_end + DT
        t_end = t_start + t_ph[k]
        rhs, args = (O.rhs_fill, (list(influent),)) if k == 0 else (O.rhs_react, ())
        if ff:
            x, so, km = ilc_phase(x, t_start, t_end, setpoints[k], rhs, args, kla_memory[j], u_batch[j], ode_kw=ode_kw)
        else:
            x, so, km = ilc_phase(x, t_start, t_end, setpoints[k], rhs, args, kla_seed=kla, ode_kw=ode_kw)
            kla = km[-1]
        So_mem.append(so); Kla_mem.append(km)
    kla5 = Kla_mem[4][-1]
    x5 = x
    t_start = t_end + DT
    t_end = t_start + t_ph[5]
    sX, Xf = O.settle(x5, t_start, t_end, ode_kw=ode_kw)
    biomass_eff, biomass_w = sX[-1], sX[0]
    Qw = (sum(sX) * WV_ILC / 10 - BIOMASS_SETPOINT_ILC * (WV_ILC - qin * t_ph[0]) - qin * t_ph[0] * biomass_eff) \
        / (biomass_w - biomass_eff)
    Qeff = qin * t_ph[0] - Qw
    t_start = t_end + DT
    t_end = t_start + t_ph[6]
    x7 = draw_fixed_qw(x5, sX, Xf, Qeff, Qw)
    t_start = t_end + DT
    t_end = t_start + t_ph[7]
    if ff:
        x8, so, km = ilc_phase(x7, t_start, t_end, setpoints[7], O.rhs_react, (), kla_memory[5], u_batch[5], ode_kw=ode_kw)
    else:
        x8, so, km = ilc_phase(x7, t_start, t_end, setpoints[7], O.rhs_react, (), kla_seed=kla5, ode_kw=ode_kw)
    So_mem.append(so); Kla_mem.append(km)
    return dict(x_last=np.array(x8), x5=np.array(x5), x7=np.array(x7), sX=np.array(sX), Xf=Xf, So_memory=So_mem,
                Kla_memory=Kla_mem, Qeff=Qeff, Qw=Qw)


def ilc_setpoint_memory(prev_sp_mem, action_value):
    """gym_SBR_env0.py:251-253: sp_memoryK_1 = sp_memoryK[:] / sp_memoryK[0] * action (0/0 -> NaN is the reference's)."""
    prev = np.array(prev_sp_mem, dtype=float)
    with np.errstate(invalid='ignore', divide='ignore'):
        return prev / prev[0] * action_value
