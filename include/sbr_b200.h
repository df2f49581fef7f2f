/*
 * sbr_b200.h -- C ABI of the B200-native batched sequencing-batch-reactor stepper.
 *
 * The reference (SungKu/gym-SBR2) has no FFI layer: its boundary is the Gym API and, underneath, one
 * `scipy.integrate.odeint(func, y0, t, args)` call per PID interval with a Python RHS callback
 * (sub_phases_FB.py:252,480,769-770; gym_SBR_oneshot.py:1647,1953,2041,2318-2319,2587).  Each entry point
 * below replaces a whole reference call tree for a BATCH of n independent environments; the reference
 * function each one stands in for is cited on the declaration.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every buffer is a DEVICE pointer owned by the caller (torch tensors on the Python
 *     side); the library allocates nothing and keeps no mutable global state; launches are asynchronous on the
 *     `stream` argument (a cudaStream_t passed as void*, NULL = legacy default stream); no implicit sync.
 *   - struct-of-arrays: component c of env i lives at  base[c * ld + i]  (ld >= n, in elements), so loads and
 *     stores of consecutive envs are coalesced.
 *   - state component order (SBR_model_FB.py:199-203):
 *       0=V 1=Si 2=Ss 3=Xi 4=Xs 5=Xbh 6=Xba 7=Xp 8=So 9=Sno 10=Snh 11=Snd 12=Xnd 13=Salk ; time unit = day.
 *   - return value: 0 = ok; < 0 = argument / launch error (text via sbr_last_error(), thread-local).
 *     Per-env numerical trouble never raises: it is reported in status[i] (SBR_ST_* bits).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns SBR_ERR_CUDA.
 */
#ifndef SBR_B200_H
#define SBR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBR_ABI_VERSION 8
#define SBR_NX 14            /* state components per env */
#define SBR_NPHASE 8         /* phases per cycle (Pons et al. B-SBR protocol) */

/* integrator modes */
#define SBR_MODE_RK4 0       /* fixed step on the reference's own output grid (len(t_range)-1 sub-steps) */
#define SBR_MODE_DP45 1      /* Dormand-Prince 5(4), per-env adaptive step, FSAL */

/* status bits */
#define SBR_ST_NONFINITE 1   /* state left the finite range                                             */
#define SBR_ST_WASTE 2       /* waste loop ended without fixing Qw (reference: NameError, sub_phases_FB.py:817-836) */
#define SBR_ST_STEPLIMIT 4   /* DP45 hit max_steps inside one interval                                 */
#define SBR_ST_LAYERS 8      /* decant layer count m outside 1..9 (reference: round(inf) OverflowError, gym_SBR_oneshot.py:2338) */
#define SBR_ST_DONE 16       /* sbr_os_step was called on an env whose episode had already ended (no-op) */

/* error codes */
#define SBR_OK 0
#define SBR_ERR_ARG -1
#define SBR_ERR_CUDA -2

/* aux[] rows written by sbr_cycle_v2 (aux[r * ld + i]) */
enum {
    SBR_AUX_OCI = 0, SBR_AUX_QW, SBR_AUX_EQI, SBR_AUX_EFF_Q, SBR_AUX_EFF_NTOT, SBR_AUX_EFF_COD,
    SBR_AUX_EFF_SNH, SBR_AUX_EFF_BOD5, SBR_AUX_EFF_SNO, SBR_AUX_KLA3_MEAN, SBR_AUX_KLA5_MEAN,
    SBR_AUX_KLA8_MEAN, SBR_AUX_ROWS
};

/* Model / plant / controller constants.  Defaults = SURVEY.md appendix A (sbr_params_default). */
typedef struct SbrParams {
    /* kinetic, Kpar (SBR_model_FB.py:38) */
    double muh, Ks, Koh, Kno, bh, etag, etah, kh, Kx, mua, Knh, ba, Koa, ka;
    /* stoichiometric, Spar (SBR_model_FB.py:36) */
    double Ya, Yh, fp, ixb, ixp;
    double so_sat;                 /* module_temperature.py:3-20, DO_set(15) */
    /* DO->KLa PID of the cycle-per-step path (gym_SBR_env2.py:48): Kc, tauI, tauD, PID dt, clamps */
    double pid_Kc, pid_tauI, pid_tauD, pid_dt, kla_min, kla_max;
    /* plant (gym_SBR_env2.py:33,85,93; SBR_model_FB.py:224-226; sub_phases_FB.py:737,662) */
    double WV, Qin, Qeff, biomass_setpoint, settler_area, settler_vmax;
    double kla0;                   /* initial KLa bias of phase 1 (gym_SBR_env2.py:56) */
    double action_scale;           /* DO set-point = action * action_scale (gym_SBR_env2.py:184-186) */
    /* interval-per-step path (gym_SBR_oneshot.py:83-96): DO and NO3 PIDs, EC dosing */
    double os_Kc_DO, os_tauI_DO, os_tauD_DO, os_Kc_EC, os_tauI_EC, os_tauD_EC;
    double os_pid_dt, ec_min, ec_max, ec_conc, do_sp_max, no_sp_max;
    double IV;                     /* initial (post-draw) volume used by the reset observation mix (gym_SBR_oneshot.py:347-364) */
} SbrParams;

/* Time schedule of one cycle (computed on the host with the reference's own float expressions,
 * sub_phases_FB.py:183-184,231; SBR_model_FB.py:17-25): per phase the number of PID intervals, the
 * number of RK4 sub-steps per interval and the interval length in days.  Phases 5 and 6 (settle, draw)
 * only use settle_time. */
typedef struct SbrSchedule {
    int32_t n_int[SBR_NPHASE];
    int32_t n_sub[SBR_NPHASE];
    double interval[SBR_NPHASE];
    double settle_time;
} SbrSchedule;

/* Rows of the persistent per-env state of the interval-per-step path (st[r * ld + i]); the reference keeps all
 * of this in module globals and ever-growing Python lists (gym_SBR_oneshot.py:207-257). */
enum {
    SBR_OS_X = 0,            /* rows 0..13: reactor state                                                     */
    SBR_OS_T = 14,           /* running time `t` (days)                                                       */
    SBR_OS_SO_PREV,          /* So[-2] (So[-1] is x[8])                                                       */
    SBR_OS_SNO_LAST,         /* Sno[-1]: x[9], except right after reset where the reference stores Ss (:1652)  */
    SBR_OS_SNO_PREV,         /* Sno[-2]                                                                       */
    SBR_OS_IE_DO,            /* DO-PID integral                                                               */
    SBR_OS_IE_EC,            /* NO3-PID integral                                                              */
    SBR_OS_EC_LAST,          /* EC[-1]: dosing flow of the last interval                                      */
    SBR_OS_H,                /* DP45 step-size proposal carried across intervals                              */
    SBR_OS_KLA_RING,         /* rows 22..31: last 10 entries of the `Kla` list, oldest first                  */
    SBR_OS_RETURN = SBR_OS_KLA_RING + 10,   /* sum of rewards since reset                                     */
    SBR_OS_STEPS,            /* env.step calls since reset                                                    */
    SBR_OS_QW,               /* waste-sludge volume of the terminal draw (the reference's global `Qw`); NaN before */
    SBR_OS_ROWS
};
#define SBR_OS_NOBS 9        /* len(obs_DO) = len(obs_EC): 5 normalised values + 4 clipped deltas             */
#define SBR_OS_NSTATE 15     /* [t, x] / x_1_state                                                            */

/* Time constants of the interval-per-step path, computed on the host with the reference's float expressions
 * (module_batch_time.py:3-116 called with t_delta = 10*dt, gym_SBR_oneshot.py:28-36).  The per-interval output
 * point count L = int(((t + t_delta) - t) / dt) (9 or 10, :1339,1384) depends on the running time and is
 * evaluated in the kernel with IEEE round-to-nearest add/sub/div (no contraction), hence bit-identical. */
typedef struct SbrOsSchedule {
    double tm3_0, tm3_1, tm4_1, tm5_1;   /* t_memory3[0], t_memory3[-1], t_memory4[-1], t_memory5[-1]          */
    double dt, t_delta;                  /* 0.002/24 and 10*dt                                                 */
    double t_fill;                       /* end of the fill phase = t_ratio[0] * 0.5 (:292,1585)               */
    double settle_len, draw_len;         /* t_ratio[5] * t_cycle, t_ratio[6] * t_cycle (:2264-2420)            */
    double t_cycle;
    int32_t fill_pts;                    /* int(t_fill / dt) = 252 output points of the fill solve (:1647)     */
    int32_t rk4_sub_interval;            /* RK4 sub-steps per interval; 0 = the reference grid (L - 1)         */
    int32_t rk4_sub_fill;                /* RK4 sub-steps of the fill solve; 0 = fill_pts - 1                  */
    int32_t rk4_sub_idle;                /* RK4 sub-steps of the idle solve; 0 = int((t_cycle - t0)/dt) - 1    */
} SbrOsSchedule;

/* Rows of the persistent per-env state of the SBR-v4 env (st[r * ld + i]); the reference keeps these in module
 * globals and growing lists (gym_SBR_env4.py:126-160, 205-222). */
enum {
    SBR_V4_X = 0,            /* rows 0..13: reactor state                                                      */
    SBR_V4_T = 14,           /* running time `t` (days)                                                        */
    SBR_V4_U,                /* DO set-point `u`, accumulated from the delta actions and clipped to [0, 8]      */
    SBR_V4_SO_PREV,          /* So[-2] (So[-1] is x[8])                                                        */
    SBR_V4_IE,               /* PID integral ie[-1]                                                            */
    SBR_V4_KLA_LAST,         /* Kla[-1]                                                                        */
    SBR_V4_KLA_SUM,          /* sum(Kla) over the episode (the terminal reward's aeration energy)              */
    SBR_V4_H,                /* DP45 step-size proposal carried across intervals                               */
    SBR_V4_RETURN,           /* sum of rewards since reset                                                     */
    SBR_V4_STEPS,            /* env.step calls since reset                                                     */
    SBR_V4_QW,               /* waste-sludge volume of the terminal draw; NaN before                           */
    SBR_V4_ROWS
};

/* Adaptive-step controls (SBR_MODE_DP45) and per-launch option flags (both modes). */
#define SBR_FLAG_RAW_KLA 1   /* sbr_cycle_v2: the actions are the KLa of phases 3, 5, 8 as fractions of kla_max (held
                                constant over the phase; the other reacting phases are unaerated), the DO -> KLa PID is
                                bypassed.  BASELINE configs[1] "random KLa actions".  No registered reference env takes a
                                raw KLa (the open-loop sim_rxn of sub_phases_PID_off.py:178-225 is dead code that ends up
                                at KLa = 0); the oracle is the same odeint-per-interval drive with the KLa held fixed. */
typedef struct SbrTol {
    double rtol, atol;     /* mixed tolerance: err_i <= atol * scale_i + rtol * |x_i|          */
    int32_t max_steps;     /* accepted + rejected steps per 72-s PID interval (default 200; typical need: 1-5);
                              the one-shot fill and idle solves of the interval-per-step path get max_steps x
                              (their length / control interval).  Bounds the work of an env that has left the
                              physical regime: one stuck env stalls its whole warp                       */
    int32_t flags;         /* SBR_FLAG_* (0 = the reference's behaviour)                                   */
} SbrTol;

int sbr_abi_version(void);
const char* sbr_last_error(void);
/* number of CUDA devices visible to the library (0 = none, compute entry points will fail) */
int sbr_device_count(void);

void sbr_params_default(SbrParams* p);

/*
 * One whole 12-h cycle for n envs = SbrEnv2.step (gym_SBR_env2.py:131-171) -> SBR_model_FB.run
 * (SBR_model_FB.py:8-295) -> filling/rxn.sim_rxn + odeint (sub_phases_FB.py:178-271, 406-500), settling
 * (:716-775), drawing + cal_eq (:780-915), module_reward.sbr_reward (module_reward.py:4-51).
 *   x0        [14][ld] in   start state
 *   influent  [14][ld] in   row 0 = fill flow m3/d (gym_SBR_env2.py:144), rows 1..13 = influent concentrations
 *   action    [3][ld]  in   raw action, clipped to [0,1] inside (gym_SBR_env2.py:133)
 *   x_last    [14][ld] out  state at the end of the idle phase
 *   obs       [3][ld]  out  [Qeff, COD_eff, Snh_eff/30]
 *   reward    [n]      out
 *   aux       [SBR_AUX_ROWS][ld] out (may be NULL)
 *   status    [n] out (may be NULL);  counters [2][ld] out: RHS evaluations, rejected steps (may be NULL)
 *   perm      [n] in (may be NULL): thread i works on env perm[i] (a permutation of 0..n-1); every buffer keeps the
 *             caller's env order.  With per-env adaptive steps a warp pays for its slowest env in every PID interval,
 *             so grouping similar envs (e.g. argsort of the first set-point) into warps removes most of the divergence
 *             (measured: 135.7 -> 86.4 ms per 2^20 cycles at rtol 1e-7).  The gathered loads/stores cost nothing
 *             here: the kernel moves 500 B per env for ~1-2 MFLOP.
 */
int sbr_cycle_v2(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                 const SbrParams* p, const SbrSchedule* s, double* x_last, double* obs, double* reward,
                 double* aux, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol,
                 const int64_t* perm, void* stream);

/*
 * sbr_cycle_v2 with a trajectory record -- what SBR_model_FB.run returns as `t`, `x` (stacked phase by phase,
 * SBR_model_FB.py:71-86 ...) and as its per-interval KLa arrays, sampled at the END of every PID interval:
 *   traj [R][SBR_TRAJ2_ROWS][ld] out, R = sbr_cycle_v2_traj_records(s) = the PID intervals of phases 1-5 and 8 (528 with the
 *   default schedule) + 1: records 0 .. 491 = phases 1-5 in order, record 492 = the post-draw state (t = start of the idle
 *   phase, KLa row 0), records 493 .. 528 = the idle phase.  Row SBR_TRAJ2_T = time in days, rows SBR_TRAJ2_X .. +13 = the
 *   14 state components, row SBR_TRAJ2_KLA = KLa of that interval.
 *   t_start [8] (host): start time of each phase (phase k starts t_delta after phase k-1 ends, SBR_model_FB.py:92,118,...;
 *   schedule.phase_bounds()).
 * A separate, untimed entry point for analysis and plotting (67 kB per env): it runs the cycle interval by interval in
 * both integrator modes, so in adaptive mode its step sequence differs from sbr_cycle_v2's three-segment kernel (same
 * results within the tolerance); every other argument as sbr_cycle_v2 (no perm).
 */
enum { SBR_TRAJ2_T = 0, SBR_TRAJ2_X = 1, SBR_TRAJ2_KLA = 15, SBR_TRAJ2_ROWS = 16 };
int sbr_cycle_v2_traj_records(const SbrSchedule* s);
int sbr_cycle_v2_traj(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                      const SbrParams* p, const SbrSchedule* s, const double* t_start, double* x_last, double* obs,
                      double* reward, double* aux, int32_t* status, uint32_t* counters, double* traj, int mode,
                      const SbrTol* tol, void* stream);

/*
 * Stage-level entry (unit tests, and the seam where the reference calls odeint): advance n envs over ONE
 * interval of length T with constant KLa (and EC dosing / fill loading) -- replaces a single
 * `odeint(dxdt, x, t_range, args=(..., Kla[i], ...))` call (sub_phases_FB.py:252,480; gym_SBR_oneshot.py:1953).
 *   tail: 0 = react, 1 = fill (loading [14][ld] required), 2 = react + EC dosing (ec [n] required)
 *   x [14][ld] in/out; kla [n]; n_sub = RK4 sub-steps (ignored by DP45)
 */
int sbr_integrate_interval(int64_t n, int64_t ld, double* x, const double* kla, const double* ec,
                           const double* loading, const SbrParams* p, int tail, double T, int n_sub,
                           int mode, const SbrTol* tol, uint32_t* counters, void* stream);

/*
 * Stage-level entry of the settle + draw phases (unit tests): settling.sim_settling (sub_phases_FB.py:716-775; the
 * reference's 10-layer odeint solve has constant settling velocity max(vmax, ...) and therefore a closed form) followed
 * by drawing.sim_drawing + cal_eq (:780-915).
 *   x [14][ld] in: state at the start of the settle phase; out: reactor state after decant and wasting
 *   sX [10][ld] out: layer solids after settling; out [9][ld]: Xf, Qw, EQI, eff[6] = [Qeff, Ntot, COD, Snh, BOD5, Sno]
 *   status [n] out (may be NULL): SBR_ST_WASTE / SBR_ST_LAYERS
 */
int sbr_settle_draw(int64_t n, int64_t ld, double* x, double settle_time, const SbrParams* p, double* sX, double* out,
                    int32_t* status, void* stream);

/* Kinetic right-hand side only: dx [14][ld] = f(x) -- replaces rxn.dxdt / filling.dxdt / reaction_dxdt
 * (sub_phases_FB.py:51-176, 278-404; gym_SBR_oneshot.py:1658-1787). */
int sbr_rhs(int64_t n, int64_t ld, const double* x, const double* kla, const double* ec,
            const double* loading, const SbrParams* p, int tail, double* dx, void* stream);

/*
 * Interval-per-step path, episode start = SbrOS.reset (gym_SBR_oneshot.py:168-438) -> Sim_filling (:1585-1654) ->
 * odeint(filling_dxdt) (:1647): one PID update at set-point 0, ONE fill solve over [0, t_fill], history seeding
 * (incl. the reference's `Sno := Ss` slip, :1652) and the reset observation from the flow-weighted mix of
 * influent and reactor (:347-364) plus 4 clipped deltas each (:388-429).
 *   x0        [14][ld] in   start state (NULL = the reference's x0_init for every env, :201-203)
 *   influent  [14][ld] in   row 0 = fill flow Qin / t_memory1[-1] (:287), rows 1..13 = influent concentrations
 *   mask      [n] in        reset only envs with mask[i] != 0 (NULL = all) -- lets a vector env restart finished
 *                           episodes while the others keep running
 *   st        [SBR_OS_ROWS][ld] out persistent per-env state
 *   obs_do, obs_ec [9][ld] out; done [n] out (cleared); status [n], counters [2][ld] may be NULL
 */
int sbr_os_reset(int64_t n, int64_t ld, const double* x0, const double* influent, const uint8_t* mask,
                 const SbrParams* p, const SbrOsSchedule* s, double* st, double* obs_do, double* obs_ec,
                 uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream);

/*
 * One env.step = SbrOS.step (gym_SBR_oneshot.py:843-1273): phase selection by the running time through four
 * NON-exclusive `if`s (:860,896,931,963 -- a step that crosses a phase boundary runs two intervals),
 * run_aero_step / run_anaero_step (:1331-1419) -> Sim_aero_rxn / Sim_anaero_rxn (:1877-2051: DO-PID -> KLa with
 * incremental bias, NO3-PID -> external-carbon flow EC) -> odeint(reaction_dxdt) (:1953,2041), reward
 * module_reward_EQIOCI.sbr_reward (module_reward_EQIOCI.py:4-115), observation epilogue (:1015-1114) and, when
 * t >= t_memory5[-1] (:1122), Sim_Settling_Drawing (:2264-2420) + Sim_idle (:2554-2597) inside the same call.
 *   st      [SBR_OS_ROWS][ld] in/out
 *   action  [2][ld] in   row 0 = DO set-point (clipped to [0, do_sp_max], aerobic phases), row 1 = NO3 set-point
 *                        (clipped to [0, no_sp_max], anoxic phases) (:862-906)
 *   obs_do, obs_ec [9][ld] out, state [15][ld] out: each may be NULL (an output the caller does not consume costs
 *                        no memory traffic); reward [n] out; done [n] in/out
 *   status [n] out (may be NULL); counters [2][ld] out (may be NULL)
 * An env whose done flag is already set is left untouched (reward 0, status SBR_ST_DONE).
 * Rows SBR_OS_KLA_RING.. of st are a CIRCULAR buffer: the KLa of the k-th interval since the reset sits in row
 * SBR_OS_KLA_RING + k % 10 (k follows from st[SBR_OS_T]); a step writes one row (two at a phase switch).
 * For full speed st, action and done should be 16-byte aligned and ld even (rows are then staged by bulk copies).
 *
 * sbr_os_step_k = K consecutive env.steps in ONE launch, the state staying in registers in between (frame-skip /
 * open-loop set-point sequences, e.g. a policy evaluated every K-th interval):
 *   action [K][2][ld] in: the set-points of step k in rows 2k, 2k+1; reward [K][ld] out: row k = reward of step k
 *   (0 for steps after the episode has ended); obs_do / obs_ec / state: the observation after the last step that
 *   ran (the terminal one if the episode ends inside the launch); status = OR, counters = sum over the K steps.
 * K = 1 is sbr_os_step; K steps of sbr_os_step give bit-identical st, rewards and final observation.
 */
int sbr_os_step(int64_t n, int64_t ld, double* st, const double* action, const SbrParams* p,
                const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream);
int sbr_os_step_k(int64_t n, int64_t ld, int K, double* st, const double* action, const SbrParams* p,
                  const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                  uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream);

/*
 * Optional trajectory dump of the interval-per-step path = what SbrOS.trajectory() hands a plotting / analysis script
 * (gym_SBR_oneshot.py:1275-1288: the module-level lists t_t, x_t, So_t, EC, reward_t, reward_EQI_t ...), sampled at
 * the ENDS of the PID intervals (the reference also keeps the 8-9 interior output points of every interval).
 * sbr_os_step_traj = sbr_os_step_k that also writes, for every interval it runs, the record
 *   traj[k][SBR_TRAJ_ROWS][ld], k = number of that interval since the reset (from 0):
 *   t, x[14], KLa, EC flow, the clipped DO / NO3 set-points in force, and -- in the record of a step's last interval --
 *   the step's reward and the diagnostics the reference appends to reward_EQI_t / reward_OCI_t / reward_AE_t /
 *   reward_EC_t: EQI/10, and the aeration / carbon cost terms normalised by their maxima and their sum
 *   (module_reward_EQIOCI.py:60-112; NaN in other records).
 * The terminal step appends two more records: the state after settle + draw and after the idle phase.
 * traj_cap = number of records the buffer holds (470 covers an episode); later records are dropped.  Not meant for
 * the timed path: 184 B per env and interval.
 */
enum { SBR_TRAJ_T = 0, SBR_TRAJ_X = 1, SBR_TRAJ_KLA = 15, SBR_TRAJ_EC = 16, SBR_TRAJ_U_DO = 17, SBR_TRAJ_U_EC = 18,
       SBR_TRAJ_REWARD = 19, SBR_TRAJ_EQI = 20, SBR_TRAJ_OCI = 21, SBR_TRAJ_AE = 22, SBR_TRAJ_ECO = 23,
       SBR_TRAJ_ROWS = 24 };
int sbr_os_step_traj(int64_t n, int64_t ld, int K, double* st, const double* action, const SbrParams* p,
                     const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                     uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol,
                     double* traj, int traj_cap, void* stream);

/*
 * SBR-v4 (SbrEnv4, gym_SBR_env4.py:71-1294): interval-per-step env whose FILL phase is stepped inside step() too,
 * with a 1-D action = change of the DO set-point.  NOTE: the reference's step() raises TypeError on numpy >= 1.18
 * (float `num` in np.linspace, :286,921,982,1207); parity is against the unmodified source run with numpy < 1.18
 * linspace semantics restored (oracle/make_golden_v4.py).
 * sbr_v4_reset = SbrEnv4.reset (:94-198): no integration; state := x0 (NULL = x0_init), t := 0, and the reset
 *   observation x_2 / x_1 with x_2 = flow-weighted mix of influent and reactor (:185-191).
 *     influent [14][ld] in: row 0 = fill flow Qin / t_memory1[-1] (:193), rows 1..13 = influent concentrations
 *     st [SBR_V4_ROWS][ld] out; obs [14][ld] out; done [n] out (cleared); mask [n] in (NULL = all)
 * sbr_v4_step = SbrEnv4.step (:200-247) -> run_step (:250-358): u := clip(u + action, 0, 8); phase from the running
 *   time (fill / react / settle+draw+idle); DO-PID with incremental bias (Sim_filling :497-534, Sim_rxn :667-707,
 *   Sim_idle :1202-1242) -> one odeint interval; terminal step = Sim_Settling_Drawing (:919-1070) + Sim_idle; reward
 *   module_reward_continuous.sbr_reward (module_reward_continuous.py:4-65); obs = x / x_1 (:236).
 *     action [n] in; influent [14][ld] in (read during the fill phase only); the schedule is the SbrOsSchedule of the
 *     SBROS-v1 path (same module_batch_time marks; t_fill doubles as t_memory1[-1])
 * order [n] (int32, may be NULL = identity): divergence-aware placement of the persistent state.  Row r of st holds at
 *   SLOT j the state of env order[j]; every other buffer (x0, influent, mask, action, obs, reward, done, status, counters)
 *   stays indexed by env.  The envs of this path need very different numbers of integrator steps per interval (1 to
 *   25), persistently so; a caller that re-sorts the slots by the previous steps' counters[0] every few steps (one
 *   sbr_permute_rows of st) gets warps whose envs finish together.  Keep runs of 4 consecutive envs together (one
 *   32-byte sector of every env-indexed row): scattering single envs makes every env-indexed access a partial sector
 *   (measured on 2^20 envs: 0.32 ms per step unordered, 0.64 ms with single envs scattered, 0.29 ms with groups of 4;
 *   0.17-0.18 ms is what the kernel takes when ALL its buffers are in sorted order, profiles/r02j_*).  Results do not
 *   depend on the placement (bit-identical, tested).
 */
int sbr_v4_reset(int64_t n, int64_t ld, const double* x0, const double* influent, const uint8_t* mask,
                 const SbrParams* p, double* st, double* obs, uint8_t* done, const int32_t* order, void* stream);
int sbr_v4_step(int64_t n, int64_t ld, double* st, const double* influent, const double* action,
                const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward, uint8_t* done,
                int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, const int32_t* order, void* stream);

/*
 * The other five interval-per-step ids of the reference, one entry-point pair for all of them:
 *   SBR_CNT_V0  SBRCnt-v0   SbrCnt0   gym_SBR_continuous0.py:85-1284   1-D action = change of the DO set-point, every
 *                                                                       reacting interval is a step, 7-value observation
 *   SBR_CNT_V1  SBRCnt-v1   SbrCnt1   gym_SBR_continuous1.py:82-1399   the agent acts in the aerobic phases only: the two
 *                                                                       anoxic phases are simulated whole inside a step
 *   SBR_CNT_V2  SBRCnt-v2   SbrCnt2   gym_SBR_continuous2.py:96-1560   as V1 + a carbon controller on Ss
 *   SBR_CNT_MA1 SBRCntMA-v1 SbrCntMA1 gym_SBR_continuous_MA1.py:96-1578  two agents sharing one 1-D action: carbon set-point
 *                                                                       (on Sno) in the anoxic phases, DO set-point in the aerobic
 *   SBR_CNT_OS2 SBROS-v2    SbrOS1    gym_SBR_oneshot1.py:98-2089      as MA1 with ABSOLUTE set-points [DO, NO3] and the
 *                                                                       SBROS-v1 observation triple (obs_DO, obs_EC, state)
 * Physics, steppers, settler and draw are those of sbr_os_step.  NOTE: in the reference every step() of these five ids
 * raises NameError inside module_reward_continuous1.sbr_reward (:32, :61); parity is against the unmodified env modules
 * run with that one function repaired as oracle/make_golden_cnt.py documents, and the reward here is that repaired form.
 *
 * SbrCntConfig: what differs between the five files as data.  gym_sbr2_b200.cnt.cnt_config(kind) fills it with the
 * reference's constants; the three time stamps come from module_batch_time.batch_time like SbrOsSchedule's.
 */
enum { SBR_CNT_V0 = 0, SBR_CNT_V1, SBR_CNT_V2, SBR_CNT_MA1, SBR_CNT_OS2, SBR_CNT_KINDS };
typedef struct SbrCntConfig {
    int32_t kind;                        /* SBR_CNT_*                                                                */
    int32_t reserved;
    double Kc_DO, tauI_DO, tauD_DO;      /* DO -> KLa PID (gym_SBR_continuous0.py:80-82: 10, 0.5, 5e-5; the others 100, 20, 0) */
    double Kc_EC, tauI_EC, tauD_EC;      /* carbon PID (V2 / OS2: 1, 20, 0; MA1: 10, 0.5, 0); unused by V0, V1          */
    double ec_conc;                      /* gCOD/m3 of the dosed carbon (V2 / OS2: 400000/20648.38*1.32; MA1: 4000/...) */
    double ec_fill_max;                  /* EC_control_par[5] = 5: upper clamp of the dosing flow, fill phase only     */
    double u_ec_init, u_ec_max;          /* carbon set-point after reset (2) and its upper clip (V2: 5; MA1 / OS2: 15)  */
    double tm2_0, tm2_1, tm4_0;          /* t_memory2[0], t_memory2[-1], t_memory4[0] (whole-phase solves of V1 / V2)   */
} SbrCntConfig;

/* Rows of the persistent per-env state st[r * ld + i] of these envs (module globals / growing lists in the reference). */
enum {
    SBR_CNT_X = 0,           /* rows 0..13: reactor state                                                      */
    SBR_CNT_T = 14,          /* running time `t`                                                               */
    SBR_CNT_U_DO,            /* DO set-point in force (`u` / `u_DO`)                                           */
    SBR_CNT_U_EC,            /* carbon-controller set-point in force (`u_EC`)                                  */
    SBR_CNT_SO_PREV,         /* So[-2]                                                                         */
    SBR_CNT_CV_LAST,         /* [-1] of the carbon controller's measured-value list (Ss / Sno)                 */
    SBR_CNT_CV_PREV,         /* [-2] of it                                                                     */
    SBR_CNT_IE_DO, SBR_CNT_IE_EC,   /* PID integrals                                                           */
    SBR_CNT_KLA_LAST,        /* Kla[-1]                                                                        */
    SBR_CNT_EC_LAST,         /* EC[-1]                                                                         */
    SBR_CNT_H,               /* DP45 step-size proposal carried across intervals                               */
    SBR_CNT_RETURN, SBR_CNT_STEPS, SBR_CNT_QW,
    SBR_CNT_ROWS
};
#define SBR_CNT_NOBS_MAX 33  /* observation rows: V0 7; V1, V2, MA1 5; OS2 9 + 9 + 15 (obs_DO, obs_EC, state)        */
/* observation rows of a kind (7 / 5 / 33), or -1 */
int sbr_cnt_obs_rows(int kind);
/*
 * sbr_cnt_reset = reset() (gym_SBR_continuous0.py:120-235 and the same function of the other four files): one fill
 *   solve over [0, t_fill] with the controllers at set-point 0, history seeding, reset observation from the
 *   flow-weighted mix of influent and reactor.
 *     x0 [14][ld] (NULL = x0_init), influent [14][ld] (row 0 = fill flow), mask [n] (NULL = all)
 *     st [SBR_CNT_ROWS][ld] out; obs [sbr_cnt_obs_rows(kind)][ld] out; done [n] out (cleared)
 * sbr_cnt_step = step(action):
 *     action [2][ld] in: row 0 = the 1-D action (V0..MA1) or the DO set-point (OS2); row 1 = the NO3 set-point (OS2 only)
 *     obs out (may be NULL), reward [n] out, done [n] in/out, status / counters may be NULL
 *   An env whose done flag is set is left untouched (reward 0, status SBR_ST_DONE).
 * The schedule is the SbrOsSchedule of the SBROS-v1 path (same module_batch_time marks).
 */
int sbr_cnt_reset(int64_t n, int64_t ld, const SbrCntConfig* cfg, const double* x0, const double* influent,
                  const uint8_t* mask, const SbrParams* p, const SbrOsSchedule* s, double* st, double* obs,
                  uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream);
int sbr_cnt_step(int64_t n, int64_t ld, const SbrCntConfig* cfg, double* st, const double* action, const SbrParams* p,
                 const SbrOsSchedule* s, double* obs, double* reward, uint8_t* done, int32_t* status,
                 uint32_t* counters, int mode, const SbrTol* tol, void* stream);

/*
 * Influent generator, the step before the path = buffer_tank3.influent.buffer_tank (buffer_tank3.py:18-108): per env
 * one shared rnd ~ N(0,1)^48 perturbs the 48-point mean profiles of the flow and the 13 concentrations
 * (profile = mean + std * rnd, std = 0.1 * mean for Ss, Xi, Xs, Xbh, Snh, Snd, Xnd and the flow, 0 otherwise, :50-66) and
 * the result is the flow-weighted mean influent_mixed = [0.66, sum(c q) / sum(q) ...] (:87-107).  Sums run
 * sequentially over the 48 points with round-to-nearest mul/add and one IEEE division, i.e. bit-identical to the
 * reference's numpy arithmetic for the same rnd.
 *   rnd      [48][ld] in   standard-normal draws (the caller's RNG: torch Philox on the device, or numpy's stream)
 *   mean,std [14][48] in   device copies of the scenario's tables (row 0 = flow), gym_sbr2_b200/data/influent_tables.npz
 *   influent [14][ld] out  influent_mixed
 */
#define SBR_INFLUENT_POINTS 48
int sbr_influent_mix(int64_t n, int64_t ld, const double* rnd, const double* mean, const double* std,
                     double* influent, void* stream);

/*
 * Influent generator with its own counter-based randomness = N independent `buffer_tank(scenario)` calls
 * (buffer_tank3.py:18-108; SbrEnv4.reset draws the scenario too: np.random.choice(8, 1), gym_SBR_env4.py:104).
 * The 48 standard normals of env i come from Philox4x32-10 keyed by `seed` with counter (env_offset + i, episode
 * number, block): they depend only on the run's seed, the env's GLOBAL index and how many episodes that env has
 * started -- not on n, the rank that owns the env, the world size or the order of resets, so results are invariant
 * to the sharding.  Mixing arithmetic = sbr_influent_mix (bit-identical for the same normals).
 *   seed        the run's seed (same on every rank)
 *   env_offset  global index of this shard's env 0
 *   epoch       [n] in/out (may be NULL): episode number per env; used for the draw and incremented for every env
 *               that draws.  NULL: every env uses epoch0.
 *   scenario    0..7, or -1 = drawn per env (uniform on 0..7); scenario_out [n] (may be NULL) receives it
 *   mean, std   [8][14][48] device copies of ALL scenario tables (row 0 = flow)
 *   mask        [n] (may be NULL): only envs with mask[i] != 0 draw (their influent column, epoch and scenario_out
 *               are updated; every other env is left untouched)
 *   influent    [14][ld] out
 * sbr_philox_normals writes the same normals z [48][ld] for epoch0 (tests; the RNG half of the parity story).
 */
int sbr_influent_sample(int64_t n, int64_t ld, uint64_t seed, int64_t env_offset, int64_t* epoch, int64_t epoch0,
                        int scenario, const double* mean, const double* std, const uint8_t* mask, double* influent,
                        int32_t* scenario_out, void* stream);
int sbr_philox_normals(int64_t n, int64_t ld, uint64_t seed, int64_t env_offset, int64_t epoch0, double* z,
                       void* stream);

/*
 * Row permutation of SoA buffers, up to SBR_PERMUTE_MAX buffers in one launch:
 *   gather  (scatter == 0): dst[k][r][i]       = src[k][r][perm[i]]
 *   scatter (scatter != 0): dst[k][r][perm[i]] = src[k][r][i]          r < rows[k], i < n
 * elem_bytes[k] is 8 (double) or 4 (int32 / uint32).  The vector env uses it to hand sbr_cycle_v2 (adaptive mode)
 * its envs in divergence-aware order -- sorted by the first DO set-point -- with unit-stride loads and stores, and to
 * put the results back into the caller's env order (no reference counterpart: the reference steps one env).
 */
#define SBR_PERMUTE_MAX 8
int sbr_permute_rows(int64_t n, const int64_t* perm, int nbuf, const void* const* src, void* const* dst,
                     const int64_t* ld_src, const int64_t* ld_dst, const int32_t* rows, const int32_t* elem_bytes,
                     int scatter, void* stream);

/*
 * Policy head of a rollout, fused into one launch on the SoA observation buffers (no reference counterpart: the
 * reference steps one env under a Python agent; BASELINE configs[4] feeds the observations of 2^20 envs to a small
 * policy every step):  action[o][i] = lo[o] + span[o] * sigmoid( sum_h w2[o][h] * tanh( sum_r w1[h][r] * obs[r][i] ) )
 * with obs = the rows of obs_a followed by the rows of obs_b (obs_b may be NULL with rows_b = 0), FP32 arithmetic.
 *   obs_a [rows_a][ld], obs_b [rows_b][ld] in (float64, e.g. obs_DO and obs_EC of sbr_os_step)
 *   w1 [hidden][rows_a + rows_b], w2 [n_out][hidden], lo / span [n_out]: device float32
 *   action [n_out][ld] out (float64, the layout sbr_os_step reads);  rows_a + rows_b <= 40, hidden <= 64, n_out <= 4
 */
int sbr_policy_mlp(int64_t n, int64_t ld, const double* obs_a, int rows_a, const double* obs_b, int rows_b,
                   const float* w1, const float* w2, const float* lo, const float* span, int hidden, int n_out,
                   double* action, void* stream);

/*
 * Fused rollout of the interval-per-step path: K consecutive env.steps in ONE launch with the policy head evaluated
 * in-kernel between them, on the registers that hold the state -- the observation never leaves the SM.  No reference
 * counterpart (the reference steps one env under a Python agent); it is sbr_os_step_k + sbr_policy_mlp called K times,
 * bit for bit, at the memory traffic of one step per K.
 *   action [2][ld] IN/OUT: in = the set-points of the launch's first step (after a reset: sbr_policy_mlp on the reset
 *          observation); out = the set-points of the step after the launch's last one, so consecutive launches chain
 *   policy: the two-layer perceptron of sbr_policy_mlp on obs = [obs_DO (9), obs_EC (9)] -> [DO set-point, NO3 set-point]
 *   reward [K][ld] out; obs_do / obs_ec / state: observation after the last step that ran (each may be NULL)
 *   act_log [K][2][ld] (may be NULL): the set-points step k ran with; obs_log [K][18][ld] (may be NULL): the
 *          observation after step k -- what a PPO update consumes.  Both are write-only streams.
 */
typedef struct SbrPolicyMlp {
    const float* w1;          /* device [hidden][n_in]  */
    const float* w2;          /* device [n_out][hidden] */
    const float* lo;          /* device [n_out]         */
    const float* span;        /* device [n_out]         */
    int32_t n_in, hidden, n_out, reserved;
} SbrPolicyMlp;
int sbr_os_rollout_k(int64_t n, int64_t ld, int K, double* st, double* action, const SbrPolicyMlp* policy,
                     const SbrParams* p, const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state,
                     double* reward, uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log,
                     int mode, const SbrTol* tol, void* stream);
/*
 * The same for SBR-v4: K consecutive SbrEnv4.step calls in one launch with a 14 -> hidden -> 1 policy head evaluated
 * in-kernel on obs = x / x_1.  Every per-env buffer is indexed like st (no `order` map): a caller that wants
 * divergence-aware placement permutes ALL of them between launches (sbr_permute_rows), which is what makes this path
 * fast -- no env-indexed access is left in the kernel (the slot placement of sbr_v4_step pays for its scattered
 * observation / action rows: 0.29 ms per step against 0.17-0.18 ms with everything sorted).
 *   action [n] IN/OUT (change of the DO set-point; chained like sbr_os_rollout_k's), reward [K][ld] out,
 *   obs [14][ld] out (may be NULL): observation after the last step that ran;
 *   act_log [K][ld], obs_log [K][14][ld] (may be NULL): per-step records for the learner.
 */
int sbr_v4_rollout_k(int64_t n, int64_t ld, int K, double* st, const double* influent, double* action,
                     const SbrPolicyMlp* policy, const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward,
                     uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log, int mode,
                     const SbrTol* tol, void* stream);
/*
 * ... and for the five ids of sbr_cnt_step: K consecutive step() calls in one launch with the policy head evaluated
 * in-kernel on the kind's observation (7 values for SBRCnt-v0, 5 for SBRCnt-v1/2 and SBRCntMA-v1, obs_DO ++ obs_EC = 18 for
 * SBROS-v2) -> 1 action (SBROS-v2: 2).
 *   action [2][ld] IN/OUT (row 1: SBROS-v2 only), reward [K][ld] out, obs out (may be NULL): observation after the last
 *   step that ran; act_log [K][2][ld], obs_log [K][n_in][ld] (may be NULL).
 */
int sbr_cnt_rollout_k(int64_t n, int64_t ld, int K, const SbrCntConfig* cfg, double* st, double* action,
                      const SbrPolicyMlp* policy, const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward,
                      uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log, int mode,
                      const SbrTol* tol, void* stream);

/*
 * Batch-to-batch (iterative-learning) feed-forward KLa of `SBR-v0` -- SURVEY.md 8(f) rank 4.  Two entry points replace
 * what SbrEnv.step does before its (unrunnable) reward call (gym_SBR_env0.py:184-203):
 *   sbr_ilc_update  = SbrEnv._take_action      -> module_batch_PID.batch_PID (module_batch_PID.py:7-275)
 *   sbr_cycle_ilc   = SbrEnv._next_observation -> SBR_model_batchPID_fbPID.run (SBR_model_batchPID_fbPID.py:8-345;
 *                     sim_rxn sub_phases_batchPID_fbPID.py:139-253, 388-500), and with kla_base == u == NULL the cycle 0
 *                     the module runs at import, SBR_model_PID_on.run (gym_SBR_env0.py:105-106; sub_phases_PID_on.py:178-271)
 * Per-env sample memories are [S][ld] (sample j of env i at base[j * ld + i]); S = layout.n_samples = the six
 * PID-controlled phases (1, 2, 3, 4, 5, 8) concatenated, one sample per output point of the reference's odeint grid
 * (phase start + n_sub per PID interval: 217 + 433 + 2008 + 1675 + 111 + 325 = 4769 with the default schedule).
 * The stepper stops at every output point (the schedule's n_sub must be the reference grid, schedule.cycle_schedule()):
 * mode SBR_MODE_RK4 = one RK4 step per point, SBR_MODE_DP45 = one adaptive Dormand-Prince solve per point (tol).
 */
typedef struct SbrIlcLayout {
    int32_t off[6];        /* sample offset of phases 1, 2, 3, 4, 5, 8 */
    int32_t n_samples;     /* S */
    int32_t tp[6];         /* window length of the batch-to-batch error, int(3 tau_w / t_delta) (module_batch_PID.py:29) */
} SbrIlcLayout;
enum {
    SBR_ILC_QEFF = 0, SBR_ILC_QW, SBR_ILC_REWARD, SBR_ILC_OCI, SBR_ILC_KLA3_MEAN, SBR_ILC_KLA5_MEAN, SBR_ILC_KLA8_MEAN,
    SBR_ILC_OUT_ROWS
};
/*
 *   x0 [14][ld], influent [14][ld] (row 0 = fill flow), sp [3][ld]: DO set-points of phases 3, 5, 8 (the others are 0;
 *     the env clips the action to [0, 5], gym_SBR_env0.py:187)
 *   p: SbrParams with THIS path's constants (pid_Kc 0.5/1.18, pid_tauI 0.0015, pid_tauD 0.005, pid_dt 0.05, kla_max 240 =
 *     the KLa cycle 0 starts from, biomass_setpoint 5400, WV 1.32, IV 0.66; gym_SBR_env0.py:40-41,89)
 *   t_fill: length of the fill phase in days (Qin = q_in t_fill)
 *   kla_base, u [S][ld] in: KLa memory of cycle 0 and u_batch; both NULL = cycle 0
 *   so_mem [S][ld] out (may be NULL: `SBR-v1`, which keeps no memories); kla_mem [S][ld] out (may be NULL): cycle 0: feedback KLa per sample (the later kla_base),
 *     feed-forward cycles: the clamped feed-forward profile the reference returns as Kla_memory
 *   x_last [14][ld] out; out [SBR_ILC_OUT_ROWS][ld] (may be NULL): Qeff, Qw, reward, OCI, mean applied KLa of phases
 *     3 / 5 / 8.  reward / OCI = module_reward.sbr_reward's formula on this cycle's applied KLa -- NOT pinned to the
 *     reference, whose call site raises TypeError (gym_SBR_env0.py:203)
 *   status [n] out (may be NULL): SBR_ST_NONFINITE / SBR_ST_STEPLIMIT; counters [2][ld] out (may be NULL): RHS
 *     evaluations, rejected steps
 */
int sbr_cycle_ilc(int64_t n, int64_t ld, const double* x0, const double* influent, const double* sp,
                  const SbrParams* p, const SbrSchedule* s, const SbrIlcLayout* lay, double t_fill,
                  const double* kla_base, const double* u, double* so_mem, double* kla_mem, double* x_last,
                  double* out, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream);
/*
 *   w, D [S] in (device, shared by all envs): window weights and their window sums, computed on the host with the
 *     reference's own expressions (gym_sbr2_b200/ilc.py)
 *   sp6 [6][ld] in: set-point memory value of the six phases (gym_SBR_env0.py:251-253)
 *   so_mem [S][ld] in: So memory the controller learns from
 *   e_sum, e_last [S][ld] in/out: sum over all cycles and last row of memory_e_batch (start at 0)
 *   u [S][ld] out: u_batch = Kc E + (Kc / tauI) e_sum + Kc tauD (E - previous E)   (module_batch_PID.py:214-262)
 *   dt = t_delta (0.002 / 24)
 */
int sbr_ilc_update(int64_t n, int64_t ld, const SbrIlcLayout* lay, const double* w, const double* D, const double* sp6,
                   const double* so_mem, double* e_sum, double* e_last, double* u, double dt, double Kc, double tauI,
                   double tauD, void* stream);

/* Per-GPU reduction of episode rewards (no reference counterpart; feeds the only collective of the design, an
 * NCCL all_gather of these 5 numbers per rank): stats[0..4] = sum, sum of squares, min, max, count over the
 * envs whose status is 0 (all envs if status == NULL).  stats must be zero-initialised by the caller with
 * stats[2] = +inf, stats[3] = -inf (sbr_reward_stats_init does it on the stream). */
int sbr_reward_stats_init(double* stats, void* stream);
int sbr_reward_stats(int64_t n, const double* reward, const int32_t* status, double* stats, void* stream);

/* FP64 pipe probe: runs `iters` rounds of 8 independent DFMA chains per thread on `blocks` x `threads`
 * and returns the number of floating-point operations issued (2 per DFMA) in *flops; time it with CUDA events
 * around the call to get the measured FP64 roofline denominator.  sink [blocks*threads] out. */
int sbr_fp64_probe(int blocks, int threads, int iters, double* sink, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SBR_B200_H */
