// sbr_cnt.cuh -- per-environment arithmetic of the remaining interval-per-step ids of the reference:
//   SBRCnt-v0  SbrCnt0    gym_SBR_continuous0.py      (kind SBR_CNT_V0)
//   SBRCnt-v1  SbrCnt1    gym_SBR_continuous1.py      (kind SBR_CNT_V1)
//   SBRCnt-v2  SbrCnt2    gym_SBR_continuous2.py      (kind SBR_CNT_V2)
//   SBRCntMA-v1 SbrCntMA1 gym_SBR_continuous_MA1.py   (kind SBR_CNT_MA1)
//   SBROS-v2   SbrOS1     gym_SBR_oneshot1.py         (kind SBR_CNT_OS2)
// Same physics as Path B (sbr_core.cuh: kinetics, tails, steppers, settler, draw); what differs between the five
// files is the controller wiring, the action semantics, which phases the agent acts in and the observation layout.
// One function pair (cnt_reset_env / cnt_step_env) serves all five: the differences are data (SbrCntConfig) and a few
// branches on `kind` that are uniform across a launch.  All integrations of a step go through ONE stepper call site
// per tail (the pass loop), as in os_step_env, so that the kernel stays inside the instruction cache.
//
// The reward of all five is module_reward_continuous1.sbr_reward, which cannot run as shipped (three unbound names);
// cnt_reward() is the repaired form documented in oracle/make_golden_cnt.py -- parity "by construction" (SURVEY 8f).
// Reference behaviour is cited as file:line relative to /root/reference/gym_SBR/envs.
#pragma once
#include "sbr_core.cuh"

namespace sbr {

struct CntCfg {
    int kind;
    double Kc_DO, KcI_DO, KcD_DO, Kc_EC, KcI_EC, KcD_EC, ec_conc, ec_fill_max, u_ec_init, u_ec_max;
    double tm2_0, tm2_1, tm4_0;
};

inline CntCfg make_cnt_cfg(const SbrCntConfig& g) {
    CntCfg q;
    q.kind = g.kind;
    q.Kc_DO = g.Kc_DO; q.KcI_DO = g.Kc_DO / g.tauI_DO; q.KcD_DO = g.Kc_DO * g.tauD_DO;
    q.Kc_EC = g.Kc_EC; q.KcI_EC = g.tauI_EC != 0.0 ? g.Kc_EC / g.tauI_EC : 0.0; q.KcD_EC = g.Kc_EC * g.tauD_EC;
    q.ec_conc = g.ec_conc; q.ec_fill_max = g.ec_fill_max; q.u_ec_init = g.u_ec_init; q.u_ec_max = g.u_ec_max;
    q.tm2_0 = g.tm2_0; q.tm2_1 = g.tm2_1; q.tm4_0 = g.tm4_0;
    return q;
}

SBR_HD constexpr bool cnt_has_ec(int kind) { return kind >= SBR_CNT_V2; }
SBR_HD constexpr int cnt_obs_rows(int kind) {
    return kind == SBR_CNT_V0 ? 7 : (kind == SBR_CNT_OS2 ? 2 * SBR_OS_NOBS + SBR_OS_NSTATE : 5);
}

// Controller scalars of one env: what the reference reads back from its module-level lists (So[-2:], the carbon
// controller's measured-value list [-2:], the two integrals, Kla[-1], EC[-1]) and the two set-points.
struct CntCtrl {
    double t, u_do, u_ec, so_prev, cv_last, cv_prev, ie_do, ie_ec, kla_last, ec_last;
};

// DO -> KLa PID, incremental bias (Sim_rxn, gym_SBR_continuous0.py:663-697 and the same block in every file;
// Sim_filling :493-527 takes the t_start == 0 branch = `first`).  PID dt = 0.002/24; two independent clamp checks.
SBR_HD double cnt_pid_do(CntCtrl& c, double so_last, double sp, bool first, double bias, const CntCfg& q,
                         const SbrParams& p, const Coef& coef) {
    const double dt = p.os_pid_dt;
    const double e = sp - so_last;
    double dcv = 0.0;
    if (!first) { dcv = (so_last - c.so_prev) * coef.os_inv_dt; c.ie_do = c.ie_do + e * dt; }
    else c.ie_do = 0.0;
    double kla = q.Kc_DO * e + q.KcI_DO * c.ie_do + q.KcD_DO * dcv + bias;
    if (kla > p.kla_max) { kla = p.kla_max; c.ie_do = c.ie_do - e * dt; }
    if (kla < p.kla_min) { kla = p.kla_min; c.ie_do = c.ie_do - e * dt; }
    return kla;
}

// Carbon controller (gym_SBR_continuous2.py:939-965 on Ss; gym_SBR_continuous_MA1.py:961-984 and
// gym_SBR_oneshot1.py:1468-1491 on Sno): error sp - cv, incremental bias EC[-1], and in the reacting phases a LOWER
// clamp only; the fill-phase copy (:733-760 / :755-781) clamps to [EC_control_par[4], EC_control_par[5]] = [0, 5].
SBR_HD double cnt_pid_ec(CntCtrl& c, double sp, bool first, bool fill, const CntCfg& q, const SbrParams& p,
                         const Coef& coef) {
    const double dt = p.os_pid_dt;
    const double e = sp - c.cv_last;
    double dcv = 0.0;
    if (!first) { dcv = (c.cv_last - c.cv_prev) * coef.os_inv_dt; c.ie_ec = c.ie_ec + e * dt; }
    else c.ie_ec = 0.0;
    double ec = q.Kc_EC * e + q.KcI_EC * c.ie_ec + q.KcD_EC * dcv + c.ec_last;
    if (fill) {
        if (ec > q.ec_fill_max) { ec = q.ec_fill_max; c.ie_ec = c.ie_ec - e * dt; }
        if (ec < 0.0) { ec = 0.0; c.ie_ec = c.ie_ec - e * dt; }
    } else if (ec < 0.0) {
        ec = 0.0; c.ie_ec = c.ie_ec - e * dt;
    }
    return ec;
}

// module_reward_continuous1.sbr_reward (module_reward_continuous1.py:5-65) with its unbound names bound as documented
// in oracle/make_golden_cnt.py: So := so; r_snh := 0 (at done: 0 / -246 on the effluent Snh when the env passes the
// effluent vector, as module_reward_continuous.py:40-51 does); r_e := 0 at done.
SBR_HD double cnt_reward(double so, bool done, bool have_eff, double eff_snh) {
    if (done) return have_eff ? (eff_snh < 4 ? 0.0 : -246.0) : 0.0;
    if (so < 1.5) return -100.0;
    if (so > 2.5 && so < 3.5) return 0.0;
    if (so >= 3.5 && so < 5) return -10.0;
    if (so >= 5) return -50.0;
    return 10.0;
}

SBR_HD double cnt_clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }   // if / elif / else

// Flow-weighted mix of influent and reactor content used by every reset observation
// (gym_SBR_continuous0.py:228-233; gym_SBR_oneshot1.py:325-344).
SBR_HD double cnt_mix(const double (&x)[SBR_NX], const Loading& load, int i, const SbrParams& p) {
    return (p.Qin * load(i) + x[i] * p.IV) / (p.Qin + p.IV);
}

// The 5-value observation of SbrCnt1 / SbrCnt2 / SbrCntMA1: [t, So, Snh] / [0.5, 8, 30] and the clipped change of So / 8
// and Snh / 20 over the step's last control interval (gym_SBR_continuous1.py:352-381).
SBR_HD void cnt_emit_obs5(const Column& obs, double t, double so, double snh, double so0, double snh0) {
    obs.set(0, t / 0.5); obs.set(1, so / 8.0); obs.set(2, snh / 30.0);
    obs.set(3, clip1((so - so0) / 8)); obs.set(4, clip1((snh - snh0) / 20));
}

// The 7-value observation of SbrCnt0: [t, Si, Xbh, Xba, So, Sno, Snh] / x_1 (gym_SBR_continuous0.py:115-117, 283-288).
SBR_HD void cnt_emit_obs7(const Column& obs, double t, double si, double xbh, double xba, double so, double sno,
                          double snh) {
    obs.set(0, t / 0.5); obs.set(1, si / 30); obs.set(2, xbh / 2599.); obs.set(3, xba / 168.); obs.set(4, so / 2.);
    obs.set(5, sno / 13.); obs.set(6, snh / 0.005);
}

// reset() of the five envs (gym_SBR_continuous0.py:120-235; ..1.py:127-275; ..2.py:141-309; .._MA1.py:141-314;
// gym_SBR_oneshot1.py:166-431): x0_init, one fill solve over [0, t_fill] with both controllers at set-point 0 on their
// t_start == 0 branch, history seeding (SbrCntMA1 / SbrOS1 store Ss in the Sno list, .._MA1.py:788), reset observation
// from the flow-weighted mix.  x: in = x0, out = state after the fill.
template <int MODE>
SBR_HD int cnt_reset_env(double (&x)[SBR_NX], const Loading& load, const CntCfg& q, const SbrParams& p,
                         const Coef& coef, const SbrOsSchedule& s, const SbrTol& tol, Dp45State& dp, CntCtrl& c,
                         const Column& obs) {
    const ObsRef x0r = obs_ref(x);
    const double so0 = x[iSo];
    const double cv0 = q.kind == SBR_CNT_V2 ? x[iSs] : x[iSno];
    c.u_do = 0.0; c.u_ec = q.u_ec_init;
    c.so_prev = so0; c.cv_last = cv0; c.cv_prev = cv0;
    c.ie_do = 0.0; c.ie_ec = 0.0; c.kla_last = 0.0; c.ec_last = 0.0;
    const double kla = cnt_pid_do(c, so0, 0.0, true, 0.0, q, p, coef);
    int status = 0;
    double ec = 0.0;
    if (cnt_has_ec(q.kind)) {
        ec = cnt_pid_ec(c, 0.0, true, true, q, p, coef);
        // a non-zero dosing flow inside the FILL solve makes the reference rewrite LSODA's state array in place on every
        // RHS call (gym_SBR_continuous2.py:634-662); it cannot happen from a non-negative start state (e <= 0 -> ec = 0)
        if (ec != 0.0) status |= SBR_ST_NONFINITE;
    } else {
        c.ie_ec = 0.0;
    }
    TailArgs a;
    a.kla = kla; a.q = load(0); a.ec_conc = 0.0; a.load = load;
    const int n_sub = s.rk4_sub_fill > 0 ? s.rk4_sub_fill : s.fill_pts - 1;
    SbrTol tl = tol;
    tl.max_steps = tol.max_steps * (int)ceil(s.t_fill / s.t_delta);
    status |= integrate_interval<TAIL_FILL, MODE>(x, s.t_fill, n_sub, coef, a, tl, dp);
    c.so_prev = so0;                   // So = [x0[8], x_fill[8]]
    c.cv_prev = cv0;                   // Ss = [x0[2], x_fill[2]] ; Sno = [x0[9], x_fill[2]] (sic)
    c.cv_last = x[iSs];
    c.kla_last = kla; c.ec_last = ec;
    c.t = s.t_fill;
    if (q.kind == SBR_CNT_V0) {
        cnt_emit_obs7(obs, c.t, cnt_mix(x, load, iSi, p), cnt_mix(x, load, iXbh, p), cnt_mix(x, load, iXba, p),
                      cnt_mix(x, load, iSo, p), cnt_mix(x, load, iSno, p), cnt_mix(x, load, iSnh, p));
    } else if (q.kind != SBR_CNT_OS2) {
        cnt_emit_obs5(obs, c.t, cnt_mix(x, load, iSo, p), cnt_mix(x, load, iSnh, p), 0.0, 0.0);
        obs.set(3, clip1((x[iSo] - x0r.So) / 8)); obs.set(4, clip1((x[iSnh] - x0r.Snh) / 20));
    } else {
        // obs_DO / obs_EC of the flow-weighted mix + clipped deltas over the fill (gym_SBR_oneshot1.py:325-399); the
        // reference's reset returns no `state`: rows 18..32 get [t, x] / x_1_state of the reactor content
        const Column od{obs.p, obs.stride}, oe{obs.p + SBR_OS_NOBS * obs.stride, obs.stride},
            os{obs.p + 2 * SBR_OS_NOBS * obs.stride, obs.stride};
        os_emit_obs(c.t, x, x0r, od, oe, os);
        const double mXbh = cnt_mix(x, load, iXbh, p), mSnh = cnt_mix(x, load, iSnh, p);
        od.set(1, mXbh / 2000.); od.set(2, cnt_mix(x, load, iXba, p) / 500.); od.set(3, cnt_mix(x, load, iSo, p) / 8.);
        od.set(4, mSnh / 10.);
        oe.set(1, cnt_mix(x, load, iSs, p) / 30.); oe.set(2, mXbh / 2000.); oe.set(3, cnt_mix(x, load, iSno, p) / 10.);
        oe.set(4, mSnh / 10.);
    }
    return status;
}

struct CntOut {
    double reward, Qw;
    int done, status;
};

// step() of the five envs.  Passes 0..3 are the reacting-phase solves of the step, pass 4 the reward / observation and,
// once t >= t_memory5[-1], settle + draw + the idle solve:
//   V0   pass 1: one control interval (gym_SBR_continuous0.py:237-275)
//   V1   pass 0: the WHOLE anoxic phase 2 in one solve at set-point 0 with a zero KLa bias, result not kept in the
//        `Kla` list (gym_SBR_continuous1.py:281-295: the caller hands Sim_rxn a fresh `[0]`); pass 1: one control
//        interval; pass 2: the whole anoxic phase 4, same way (:331-344) -- the agent only acts in the aerobic phases
//   V2   as V1, but both solves also run the carbon controller on Ss, whose set-point moves with the action in the
//        very first step only (gym_SBR_continuous2.py:319-356), and the tuple unpacking REBINDS the global `Kla` to the
//        fresh list (:351), so the whole-phase KLa becomes the next bias
//   MA1, OS2  passes 0..3: the four NON-exclusive ifs on the running time (gym_SBR_continuous_MA1.py:330-411,
//        gym_SBR_oneshot1.py:472-552): anoxic (carbon set-point acts, DO set-point 0) / aerobic (DO set-point acts,
//        carbon set-point 0) / anoxic / aerobic; a step that crosses a boundary runs two intervals and -- MA1 -- adds the
//        action to both set-points
template <int MODE>
SBR_HD void cnt_step_env(double (&x)[SBR_NX], CntCtrl& c, double a0, double a1, const CntCfg& q, const SbrParams& p,
                         const Coef& coef, const SbrOsSchedule& s, const SbrTol& tol, Dp45State& dp, const Column& obs,
                         CntOut& o) {
    const int kind = q.kind;
    const bool has_ec = cnt_has_ec(kind);
    const bool four_ifs = kind == SBR_CNT_MA1 || kind == SBR_CNT_OS2;
    int status = 0;
    ObsRef first = obs_ref(x);
    double so_obs = x[iSo], snh_obs = x[iSnh];
    TailArgs a;
    a.kla = 0.0; a.q = 0.0; a.ec_conc = q.ec_conc; a.load = Loading{nullptr, 0};
    SbrTol tl = tol;
    o.done = 0; o.Qw = NAN; o.reward = 0.0;
    if (kind == SBR_CNT_V0 || kind == SBR_CNT_V1 || kind == SBR_CNT_V2) {
        c.u_do = cnt_clip(c.u_do + a0, 0.0, p.do_sp_max);                // u = u + action, clipped to [0, 8]
        if (kind == SBR_CNT_V2) c.u_ec = cnt_clip(c.u_ec, 0.0, q.u_ec_max);
    }
    for (int pass = 0; pass < 5; ++pass) {
        const double t = c.t;
        double T, t_next, sp_do = 0.0, sp_ec = 0.0;
        int n_sub;
        bool run_ec = has_ec, zero_bias = false, keep_kla = true;
        if (pass < 4) {
            bool whole = false;
            if (four_ifs) {
                const bool cond = pass == 0 ? (t < s.tm3_0)
                                : pass == 1 ? (t >= s.tm3_0 && t <= s.tm3_1)
                                : pass == 2 ? (t > s.tm3_1 && t <= s.tm4_1)
                                            : (t > s.tm4_1);
                if (!cond) continue;
                if ((pass & 1) == 0) {                                   // anoxic: the carbon set-point acts
                    c.u_ec = cnt_clip(kind == SBR_CNT_MA1 ? c.u_ec + a0 : a1, 0.0, q.u_ec_max);
                    c.u_do = 0.0;
                } else {                                                 // aerobic: the DO set-point acts
                    c.u_do = cnt_clip(kind == SBR_CNT_MA1 ? c.u_do + a0 : a0, 0.0, p.do_sp_max);
                    c.u_ec = 0.0;
                }
                sp_do = c.u_do; sp_ec = c.u_ec;
            } else if (kind == SBR_CNT_V0) {
                if (pass != 1) continue;
                sp_do = c.u_do;
            } else {
                if (pass == 3) continue;
                if (pass == 0) {
                    if (!(t < q.tm2_0)) continue;
                    if (kind == SBR_CNT_V2) c.u_ec = cnt_clip(c.u_ec + a0, 0.0, q.u_ec_max);
                    whole = true;
                } else if (pass == 2) {
                    if (!(t >= s.tm3_1 && t < s.tm4_1)) continue;
                    whole = true;
                }
                sp_do = whole ? 0.0 : c.u_do;
                sp_ec = c.u_ec;
            }
            if (whole) {
                // one odeint call over the stamps of the whole phase: it integrates from t_memoryK[0] to t_memoryK[-1]
                const double t0 = pass == 0 ? q.tm2_0 : q.tm4_0;
                t_next = pass == 0 ? q.tm2_1 : s.tm4_1;
                T = sub_rn(t_next, t0);
                const int n_iv = (int)ceil(T / s.t_delta);
                n_sub = n_iv * (s.rk4_sub_interval > 0 ? s.rk4_sub_interval : 9);
                tl.max_steps = tol.max_steps * n_iv;
                zero_bias = true;
                keep_kla = kind == SBR_CNT_V2;
            } else {
                // run_step (gym_SBR_continuous0.py:326-358): t_range = linspace(t, t + t_delta, int(((t + t_delta) - t) / dt))
                t_next = add_rn(t, s.t_delta);
                T = sub_rn(t_next, t);
                const int L = (int)div_rn(T, s.dt);
                n_sub = s.rk4_sub_interval > 0 ? s.rk4_sub_interval : (L > 1 ? L - 1 : 1);
                tl.max_steps = tol.max_steps;
                first = obs_ref(x);
            }
        } else {
            // reward and observation of the step (module_reward_continuous1.py; gym_SBR_continuous1.py:346-381): for V1 / V2
            // they describe the END OF THE CONTROL INTERVAL even when the whole anoxic phase 4 was simulated after it,
            // while the time is the running time after everything
            o.reward = cnt_reward(so_obs, false, false, 0.0);
            if (kind == SBR_CNT_V0)
                cnt_emit_obs7(obs, c.t, x[iSi], x[iXbh], x[iXba], x[iSo], x[iSno], x[iSnh]);
            else if (kind != SBR_CNT_OS2)
                cnt_emit_obs5(obs, c.t, so_obs, snh_obs, first.So, first.Snh);
            const Column od{obs.p, obs.stride}, oe{obs.p ? obs.p + SBR_OS_NOBS * obs.stride : nullptr, obs.stride},
                os{obs.p ? obs.p + 2 * SBR_OS_NOBS * obs.stride : nullptr, obs.stride};
            if (kind == SBR_CNT_OS2) os_emit_obs(c.t, x, first, od, oe, os);
            if (!(c.t >= s.tm5_1)) break;
            // end of the reacting phases (gym_SBR_continuous0.py:291-322 and the same block in every file):
            // Sim_Settling_Drawing + Sim_idle inside this step
            o.done = 1;
            if (kind == SBR_CNT_OS2) first = obs_ref(x);                // x_out[0] of the stacked settle/draw/idle output
            const double t_set_end = add_rn(c.t, s.settle_len);
            double sX[10], Xf;
            settle_closed_form(x, sub_rn(t_set_end, c.t), p.settler_area, p.settler_vmax, sX, Xf);
            DrawOut d;
            draw_and_waste(x, sX, Xf, p.Qeff, p.biomass_setpoint, d);
            status |= d.status;
            o.Qw = d.Qw;
            if (kind == SBR_CNT_V0) o.reward = cnt_reward(0.0, true, true, d.eff[3]);      // :311
            // SbrOS1 recomputes obs / state from the POST-DRAW state at the end-of-cycle time (gym_SBR_oneshot1.py:703-775);
            // the deltas (post-idle minus pre-settle) are patched in after the idle solve
            if (kind == SBR_CNT_OS2) os_emit_obs(s.t_cycle, x, first, od, oe, os);
            const double t_draw_end = add_rn(t_set_end, s.draw_len);
            c.so_prev = x[iSo];                                          // So padded with the frozen value over settle + draw
            T = sub_rn(s.t_cycle, t_draw_end);
            const int n_iv = (int)ceil(T / s.t_delta);
            tl.max_steps = tol.max_steps * n_iv;
            const int pts = (int)div_rn(T, s.dt);
            n_sub = s.rk4_sub_idle > 0 ? s.rk4_sub_idle : (pts > 1 ? pts - 1 : 1);
            sp_do = c.u_do;
            run_ec = false;                                              // Sim_idle runs the DO controller only
            t_next = s.t_cycle;
        }
        // ---- Sim_rxn / Sim_idle: both controllers, then ONE odeint call ----
        const double so_start = x[iSo];
        a.kla = cnt_pid_do(c, so_start, sp_do, false, zero_bias ? 0.0 : c.kla_last, q, p, coef);
        a.q = run_ec ? cnt_pid_ec(c, sp_ec, false, false, q, p, coef) : 0.0;
        if (warp_any(a.q != 0.0)) status |= integrate_interval<TAIL_EC, MODE>(x, T, n_sub, coef, a, tl, dp);
        else status |= integrate_interval<TAIL_REACT, MODE>(x, T, n_sub, coef, a, tl, dp);
        c.so_prev = so_start;
        if (keep_kla) c.kla_last = a.kla;
        if (run_ec) {
            c.cv_prev = c.cv_last;
            c.cv_last = kind == SBR_CNT_V2 ? x[iSs] : x[iSno];
            c.ec_last = a.q;
        }
        c.t = t_next;
        if (pass == 1 || four_ifs) { so_obs = x[iSo]; snh_obs = x[iSnh]; }
        if (pass == 4) {
            if (kind == SBR_CNT_V0)                                      // state from the post-idle reactor (:313-320)
                cnt_emit_obs7(obs, c.t, x[iSi], x[iXbh], x[iXba], x[iSo], x[iSno], x[iSnh]);
            if (kind == SBR_CNT_OS2) {
                const Column od{obs.p, obs.stride}, oe{obs.p ? obs.p + SBR_OS_NOBS * obs.stride : nullptr, obs.stride};
                os_emit_deltas(x, first, od, oe);
            }
        }
    }
    bool finite = fabs(o.reward) < 1e300;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) finite = finite && (fabs(x[i]) < 1e300);
    if (!finite) status |= SBR_ST_NONFINITE;
    o.status = status;
}

}  // namespace sbr
