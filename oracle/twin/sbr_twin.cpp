// CPU TWIN -- TEST INFRASTRUCTURE ONLY (never linked into, imported by, or called from the product path).
//
// Compiles the product's per-env arithmetic (gym_sbr2_b200/csrc/sbr_core.cuh, the same source the CUDA
// kernels inline) with g++ so that tests can (a) debug the stepper logic against the scipy oracle
// (oracle/sbr_oracle.py) in a container without a GPU and (b) check every env of a large GPU batch at ~1e-11
// instead of a 64-env subset at the oracle's 1e-5.  It is NOT an independent restatement of the reference --
// that is oracle/sbr_oracle.py, which is pinned to the reference's own outputs in tests/golden/.
// Same SoA host buffers and argument meaning as include/sbr_b200.h, minus the stream.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../gym_sbr2_b200/csrc/sbr_core.cuh"
#include "../../gym_sbr2_b200/csrc/sbr_cnt.cuh"
#include "../../gym_sbr2_b200/csrc/sbr_ilc.cuh"

using namespace sbr;

extern "C" {

int twin_cycle_v2(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                  const SbrParams* p, const SbrSchedule* s, double* x_last, double* obs, double* reward,
                  double* aux, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 200; t.flags = 0;
    if (tol) t = *tol;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        double x[SBR_NX], a[3], load[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) { x[k] = x0[k * ld + i]; load[k] = influent[k * ld + i]; }
        for (int k = 0; k < 3; ++k) a[k] = action[k * ld + i];
        Dp45State st;
        st.h = s->interval[0] / (double)s->n_sub[0];
        st.n_rhs = 0; st.n_rej = 0;
        CycleOut o;
        Loading L{load, 1};
        double parkbuf[PARK_SLOTS];
        const Park park{parkbuf, 1};
        if (mode == SBR_MODE_RK4) cycle_v2<SBR_MODE_RK4>(x, a, L, load[0], *p, c, *s, t, st, o, park);
        else cycle_v2<SBR_MODE_DP45>(x, a, L, load[0], *p, c, *s, t, st, o, park);
        for (int k = 0; k < SBR_NX; ++k) x_last[k * ld + i] = x[k];
        for (int k = 0; k < 3; ++k) obs[k * ld + i] = o.obs[k];
        reward[i] = o.reward;
        if (aux) for (int k = 0; k < SBR_AUX_ROWS; ++k) aux[k * ld + i] = o.aux[k];
        if (status) status[i] = o.status;
        if (counters) { counters[i] = st.n_rhs; counters[ld + i] = st.n_rej; }
    }
    return 0;
}

int twin_integrate_interval(int64_t n, int64_t ld, double* x, const double* kla, const double* ec,
                            const double* loading, const SbrParams* p, int tail, double T, int n_sub, int mode,
                            const SbrTol* tol, uint32_t* counters) {
    const Coef c = make_coef(*p);
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 200; t.flags = 0;
    if (tol) t = *tol;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        double xx[SBR_NX], load[SBR_NX] = {0};
        for (int k = 0; k < SBR_NX; ++k) xx[k] = x[k * ld + i];
        TailArgs a;
        a.kla = kla[i]; a.q = 0.0; a.ec_conc = p->ec_conc; a.load = Loading{load, 1};
        if (tail == TAIL_FILL) { for (int k = 0; k < SBR_NX; ++k) load[k] = loading[k * ld + i]; a.q = load[0]; }
        if (tail == TAIL_EC) a.q = ec[i];
        Dp45State st;
        st.h = T / (double)n_sub; st.n_rhs = 0; st.n_rej = 0;
#define TW(TAIL, MODE) integrate_interval<TAIL, MODE>(xx, T, n_sub, c, a, t, st)
        if (mode == SBR_MODE_RK4) {
            if (tail == TAIL_REACT) TW(TAIL_REACT, SBR_MODE_RK4);
            else if (tail == TAIL_FILL) TW(TAIL_FILL, SBR_MODE_RK4);
            else TW(TAIL_EC, SBR_MODE_RK4);
        } else {
            if (tail == TAIL_REACT) TW(TAIL_REACT, SBR_MODE_DP45);
            else if (tail == TAIL_FILL) TW(TAIL_FILL, SBR_MODE_DP45);
            else TW(TAIL_EC, SBR_MODE_DP45);
        }
#undef TW
        for (int k = 0; k < SBR_NX; ++k) x[k * ld + i] = xx[k];
        if (counters) { counters[i] = st.n_rhs; counters[ld + i] = st.n_rej; }
    }
    return 0;
}

int twin_rhs(int64_t n, int64_t ld, const double* x, const double* kla, const double* ec, const double* loading,
             const SbrParams* p, int tail, double* dx) {
    const Coef c = make_coef(*p);
    for (int64_t i = 0; i < n; ++i) {
        double xx[SBR_NX], k[SBR_NX] = {0}, load[SBR_NX] = {0};
        for (int j = 0; j < SBR_NX; ++j) xx[j] = x[j * ld + i];
        TailArgs a;
        a.kla = kla[i]; a.q = 0.0; a.ec_conc = p->ec_conc; a.load = Loading{load, 1};
        if (tail == TAIL_FILL) { for (int j = 0; j < SBR_NX; ++j) load[j] = loading[j * ld + i]; a.q = load[0]; }
        if (tail == TAIL_EC) a.q = ec[i];
        a.kla_sat = a.kla * c.so_sat;
        if (tail == TAIL_REACT) rhs<TAIL_REACT>(xx, k, c, a);
        else if (tail == TAIL_FILL) rhs<TAIL_FILL>(xx, k, c, a);
        else rhs<TAIL_EC>(xx, k, c, a);
        for (int j = 0; j < SBR_NX; ++j) dx[j * ld + i] = k[j];
    }
    return 0;
}

static const double kX0Init[SBR_NX] = {   // gym_SBR_oneshot.py:201-203
    0.6161484733495801, 30, 0.571098000538576, 1440.01157895393, 31.254221999137, 2599.2714348941,
    168.915006750837, 551.901552960823, 2.16607843793004, 13.3791460027604, 0.00562880208518134,
    0.35996687629947, 1.86916737961228, 3.790463057094611};

static SbrTol tol_or_default(const SbrTol* tol) {
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 200; t.flags = 0;
    if (tol) t = *tol;
    return t;
}

static void store_ctrl(double* st, int64_t ld, int64_t i, const OsCtrl& c, double h) {
    st[SBR_OS_T * ld + i] = c.t; st[SBR_OS_SO_PREV * ld + i] = c.so_prev;
    st[SBR_OS_SNO_LAST * ld + i] = c.sno_last; st[SBR_OS_SNO_PREV * ld + i] = c.sno_prev;
    st[SBR_OS_IE_DO * ld + i] = c.ie_do; st[SBR_OS_IE_EC * ld + i] = c.ie_ec;
    st[SBR_OS_EC_LAST * ld + i] = c.ec_last; st[SBR_OS_H * ld + i] = h;
}

int twin_os_reset(int64_t n, int64_t ld, const double* x0, const double* influent, const uint8_t* mask,
                  const SbrParams* p, const SbrOsSchedule* s, double* st, double* obs_do, double* obs_ec,
                  uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    const SbrTol t = tol_or_default(tol);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        if (mask && mask[i] == 0) continue;
        double x[SBR_NX], load[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) { x[k] = x0 ? x0[k * ld + i] : kX0Init[k]; load[k] = influent[k * ld + i]; }
        Dp45State dp;
        dp.h = s->t_fill / (double)(s->fill_pts > 1 ? s->fill_pts - 1 : 1); dp.n_rhs = 0; dp.n_rej = 0;
        OsCtrl ctl;
        const Column ring{st + SBR_OS_KLA_RING * ld + i, ld}, od{obs_do + i, ld}, oe{obs_ec + i, ld};
        const Loading L{load, 1};
        int stt = mode == SBR_MODE_RK4 ? os_reset_env<SBR_MODE_RK4>(x, L, *p, c, *s, t, dp, ctl, ring, od, oe)
                                       : os_reset_env<SBR_MODE_DP45>(x, L, *p, c, *s, t, dp, ctl, ring, od, oe);
        for (int k = 0; k < SBR_NX; ++k) st[k * ld + i] = x[k];
        store_ctrl(st, ld, i, ctl, s->t_delta / 9.0);
        st[SBR_OS_RETURN * ld + i] = 0.0; st[SBR_OS_STEPS * ld + i] = 0.0; st[SBR_OS_QW * ld + i] = NAN;
        done[i] = 0;
        if (status) status[i] = stt;
        if (counters) { counters[i] = dp.n_rhs; counters[ld + i] = dp.n_rej; }
    }
    return 0;
}

int twin_os_step(int64_t n, int64_t ld, double* st, const double* action, const SbrParams* p,
                 const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                 uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    const SbrTol t = tol_or_default(tol);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        double x[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) x[k] = st[k * ld + i];
        OsCtrl ctl;
        ctl.t = st[SBR_OS_T * ld + i];
        const Column od{obs_do + i, ld}, oe{obs_ec + i, ld}, os{state + i, ld};
        if (done[i]) {
            os_emit_obs(ctl.t, x, obs_ref(x), od, oe, os);
            reward[i] = 0.0;
            if (status) status[i] = SBR_ST_DONE;
            if (counters) { counters[i] = 0; counters[ld + i] = 0; }
            continue;
        }
        ctl.so_prev = st[SBR_OS_SO_PREV * ld + i]; ctl.sno_last = st[SBR_OS_SNO_LAST * ld + i];
        ctl.sno_prev = st[SBR_OS_SNO_PREV * ld + i]; ctl.ie_do = st[SBR_OS_IE_DO * ld + i];
        ctl.ie_ec = st[SBR_OS_IE_EC * ld + i]; ctl.ec_last = st[SBR_OS_EC_LAST * ld + i];
        Dp45State dp;
        dp.h = st[SBR_OS_H * ld + i]; dp.n_rhs = 0; dp.n_rej = 0;
        const Column rcol{st + SBR_OS_KLA_RING * ld + i, ld};
        KlaRing ring{rcol, rcol, os_ring_head(ctl.t, *s)};
        ctl.kla_last = ring.back(1);
        OsStepOut o;
        if (mode == SBR_MODE_RK4)
            os_step_env<SBR_MODE_RK4>(x, ctl, ring, action[i], action[ld + i], *p, c, *s, t, dp, o, OsTraj{nullptr, 0, 0});
        else
            os_step_env<SBR_MODE_DP45>(x, ctl, ring, action[i], action[ld + i], *p, c, *s, t, dp, o, OsTraj{nullptr, 0, 0});
        os_emit_obs(ctl.t, x, o.first, od, oe, os);
        for (int k = 0; k < SBR_NX; ++k) st[k * ld + i] = x[k];
        store_ctrl(st, ld, i, ctl, dp.h);
        st[SBR_OS_RETURN * ld + i] += o.reward;
        st[SBR_OS_STEPS * ld + i] += 1.0;
        if (o.done) { st[SBR_OS_QW * ld + i] = o.Qw; done[i] = 1; }
        reward[i] = o.reward;
        if (status) status[i] = o.status;
        if (counters) { counters[i] = dp.n_rhs; counters[ld + i] = dp.n_rej; }
    }
    return 0;
}

int twin_v4_reset(int64_t n, int64_t ld, const double* x0, const double* influent, const uint8_t* mask,
                  const SbrParams* p, double* st, double* obs, uint8_t* done) {
    for (int64_t i = 0; i < n; ++i) {
        if (mask && mask[i] == 0) continue;
        double x[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) { x[k] = x0 ? x0[k * ld + i] : kX0Init[k]; st[k * ld + i] = x[k]; }
        const Loading load{influent + i, (int)ld};
        v4_reset_obs(x, load, *p, Column{obs + i, ld});
        for (int r = SBR_V4_T; r < SBR_V4_ROWS; ++r) st[r * ld + i] = 0.0;
        st[SBR_V4_QW * ld + i] = NAN;
        done[i] = 0;
    }
    return 0;
}

int twin_v4_step(int64_t n, int64_t ld, double* st, const double* influent, const double* action,
                 const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward, uint8_t* done,
                 int32_t* status, uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    const SbrTol t = tol_or_default(tol);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        double x[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) x[k] = st[k * ld + i];
        const Column ob{obs + i, ld};
        if (done[i]) {
            for (int k = 0; k < SBR_NX; ++k) ob.set(k, x[k] * inv_x1_v4(k));
            reward[i] = 0.0;
            if (status) status[i] = SBR_ST_DONE;
            if (counters) { counters[i] = 0; counters[ld + i] = 0; }
            continue;
        }
        V4Ctrl ctl;
        ctl.t = st[SBR_V4_T * ld + i]; ctl.u = st[SBR_V4_U * ld + i]; ctl.so_prev = st[SBR_V4_SO_PREV * ld + i];
        ctl.ie = st[SBR_V4_IE * ld + i]; ctl.kla_last = st[SBR_V4_KLA_LAST * ld + i];
        ctl.kla_sum = st[SBR_V4_KLA_SUM * ld + i];
        Dp45State dp;
        dp.h = st[SBR_V4_H * ld + i];
        if (!(dp.h > 0.0)) dp.h = s->t_delta / 9.0;
        dp.n_rhs = 0; dp.n_rej = 0;
        const Loading load{influent + i, (int)ld};
        V4Out o;
        if (mode == SBR_MODE_RK4) v4_step_env<SBR_MODE_RK4>(x, ctl, action[i], load, *p, c, *s, t, dp, ob, o);
        else v4_step_env<SBR_MODE_DP45>(x, ctl, action[i], load, *p, c, *s, t, dp, ob, o);
        for (int k = 0; k < SBR_NX; ++k) st[k * ld + i] = x[k];
        st[SBR_V4_T * ld + i] = ctl.t; st[SBR_V4_U * ld + i] = ctl.u; st[SBR_V4_SO_PREV * ld + i] = ctl.so_prev;
        st[SBR_V4_IE * ld + i] = ctl.ie; st[SBR_V4_KLA_LAST * ld + i] = ctl.kla_last;
        st[SBR_V4_KLA_SUM * ld + i] = ctl.kla_sum; st[SBR_V4_H * ld + i] = dp.h;
        st[SBR_V4_RETURN * ld + i] += o.reward; st[SBR_V4_STEPS * ld + i] += 1.0;
        if (o.done) { st[SBR_V4_QW * ld + i] = o.Qw; done[i] = 1; }
        reward[i] = o.reward;
        if (status) status[i] = o.status;
        if (counters) { counters[i] = dp.n_rhs; counters[ld + i] = dp.n_rej; }
    }
    return 0;
}

static void twin_cnt_load(const double* st, int64_t ld, int64_t i, CntCtrl& c, double& h) {
    c.t = st[SBR_CNT_T * ld + i]; c.u_do = st[SBR_CNT_U_DO * ld + i]; c.u_ec = st[SBR_CNT_U_EC * ld + i];
    c.so_prev = st[SBR_CNT_SO_PREV * ld + i]; c.cv_last = st[SBR_CNT_CV_LAST * ld + i];
    c.cv_prev = st[SBR_CNT_CV_PREV * ld + i]; c.ie_do = st[SBR_CNT_IE_DO * ld + i]; c.ie_ec = st[SBR_CNT_IE_EC * ld + i];
    c.kla_last = st[SBR_CNT_KLA_LAST * ld + i]; c.ec_last = st[SBR_CNT_EC_LAST * ld + i]; h = st[SBR_CNT_H * ld + i];
}
static void twin_cnt_store(double* st, int64_t ld, int64_t i, const CntCtrl& c, double h) {
    st[SBR_CNT_T * ld + i] = c.t; st[SBR_CNT_U_DO * ld + i] = c.u_do; st[SBR_CNT_U_EC * ld + i] = c.u_ec;
    st[SBR_CNT_SO_PREV * ld + i] = c.so_prev; st[SBR_CNT_CV_LAST * ld + i] = c.cv_last;
    st[SBR_CNT_CV_PREV * ld + i] = c.cv_prev; st[SBR_CNT_IE_DO * ld + i] = c.ie_do; st[SBR_CNT_IE_EC * ld + i] = c.ie_ec;
    st[SBR_CNT_KLA_LAST * ld + i] = c.kla_last; st[SBR_CNT_EC_LAST * ld + i] = c.ec_last; st[SBR_CNT_H * ld + i] = h;
}

int twin_cnt_reset(int64_t n, int64_t ld, const SbrCntConfig* cfg, const double* x0, const double* influent,
                   const uint8_t* mask, const SbrParams* p, const SbrOsSchedule* s, double* st, double* obs,
                   uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    const CntCfg q = make_cnt_cfg(*cfg);
    const SbrTol t = tol_or_default(tol);
    for (int64_t i = 0; i < n; ++i) {
        if (mask && mask[i] == 0) continue;
        double x[SBR_NX], load[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) { x[k] = x0 ? x0[k * ld + i] : kX0Init[k]; load[k] = influent[k * ld + i]; }
        Dp45State dp;
        dp.h = s->t_fill / (double)(s->fill_pts > 1 ? s->fill_pts - 1 : 1); dp.n_rhs = 0; dp.n_rej = 0;
        CntCtrl ctl;
        const Loading L{load, 1};
        const int stt = mode == SBR_MODE_RK4
            ? cnt_reset_env<SBR_MODE_RK4>(x, L, q, *p, c, *s, t, dp, ctl, Column{obs + i, ld})
            : cnt_reset_env<SBR_MODE_DP45>(x, L, q, *p, c, *s, t, dp, ctl, Column{obs + i, ld});
        for (int k = 0; k < SBR_NX; ++k) st[k * ld + i] = x[k];
        twin_cnt_store(st, ld, i, ctl, s->t_delta / 9.0);
        st[SBR_CNT_RETURN * ld + i] = 0.0; st[SBR_CNT_STEPS * ld + i] = 0.0; st[SBR_CNT_QW * ld + i] = NAN;
        done[i] = 0;
        if (status) status[i] = stt;
        if (counters) { counters[i] = dp.n_rhs; counters[ld + i] = dp.n_rej; }
    }
    return 0;
}

int twin_cnt_step(int64_t n, int64_t ld, const SbrCntConfig* cfg, double* st, const double* action, const SbrParams* p,
                  const SbrOsSchedule* s, double* obs, double* reward, uint8_t* done, int32_t* status,
                  uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    const CntCfg q = make_cnt_cfg(*cfg);
    const SbrTol t = tol_or_default(tol);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        if (done[i]) {
            reward[i] = 0.0;
            if (status) status[i] = SBR_ST_DONE;
            if (counters) { counters[i] = 0; counters[ld + i] = 0; }
            continue;
        }
        double x[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) x[k] = st[k * ld + i];
        CntCtrl ctl;
        Dp45State dp;
        twin_cnt_load(st, ld, i, ctl, dp.h);
        if (!(dp.h > 0.0)) dp.h = s->t_delta / 9.0;
        dp.n_rhs = 0; dp.n_rej = 0;
        const double a0 = action[i], a1 = q.kind == SBR_CNT_OS2 ? action[ld + i] : 0.0;
        CntOut o;
        const Column ob{obs ? obs + i : nullptr, ld};
        if (mode == SBR_MODE_RK4) cnt_step_env<SBR_MODE_RK4>(x, ctl, a0, a1, q, *p, c, *s, t, dp, ob, o);
        else cnt_step_env<SBR_MODE_DP45>(x, ctl, a0, a1, q, *p, c, *s, t, dp, ob, o);
        for (int k = 0; k < SBR_NX; ++k) st[k * ld + i] = x[k];
        twin_cnt_store(st, ld, i, ctl, dp.h);
        st[SBR_CNT_RETURN * ld + i] += o.reward; st[SBR_CNT_STEPS * ld + i] += 1.0;
        if (o.done) { st[SBR_CNT_QW * ld + i] = o.Qw; done[i] = 1; }
        reward[i] = o.reward;
        if (status) status[i] = o.status;
        if (counters) { counters[i] = dp.n_rhs; counters[ld + i] = dp.n_rej; }
    }
    return 0;
}

int twin_cycle_ilc(int64_t n, int64_t ld, const double* x0, const double* influent, const double* sp, const SbrParams* p,
                   const SbrSchedule* s, const SbrIlcLayout* lay, double t_fill, const double* kla_base, const double* u,
                   double* so_mem, double* kla_mem, double* x_last, double* out, int32_t* status, uint32_t* counters,
                   int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 200; t.flags = 0;
    if (tol) t = *tol;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n; ++i) {
        double x[SBR_NX], load[SBR_NX];
        for (int k = 0; k < SBR_NX; ++k) { x[k] = x0[k * ld + i]; load[k] = influent[k * ld + i]; }
        const double sp8[8] = {0.0, 0.0, sp[i], 0.0, sp[ld + i], 0.0, 0.0, sp[2 * ld + i]};
        const bool ff = kla_base != nullptr;
        IlcIo io;
        io.so = Column{so_mem ? so_mem + i : nullptr, ld};
        io.kla_mem = Column{kla_mem ? kla_mem + i : nullptr, ld};
        io.kla_base = Column{ff ? const_cast<double*>(kla_base) + i : nullptr, ld};
        io.u = Column{ff ? const_cast<double*>(u) + i : nullptr, ld};
        const int off[6] = {lay->off[0], lay->off[1], lay->off[2], lay->off[3], lay->off[4], lay->off[5]};
        IlcOut o;
        Dp45State st;
        st.h = s->interval[0] / (double)s->n_sub[0]; st.n_rhs = 0; st.n_rej = 0;
        if (mode == SBR_MODE_RK4) cycle_ilc<SBR_MODE_RK4>(x, sp8, Loading{load, 1}, load[0], t_fill, ff, *p, c, *s, t, st, io, off, o);
        else cycle_ilc<SBR_MODE_DP45>(x, sp8, Loading{load, 1}, load[0], t_fill, ff, *p, c, *s, t, st, io, off, o);
        for (int k = 0; k < SBR_NX; ++k) x_last[k * ld + i] = x[k];
        if (out) {
            out[SBR_ILC_QEFF * ld + i] = o.Qeff; out[SBR_ILC_QW * ld + i] = o.Qw;
            out[SBR_ILC_REWARD * ld + i] = o.reward; out[SBR_ILC_OCI * ld + i] = o.OCI;
            out[SBR_ILC_KLA3_MEAN * ld + i] = o.kla_mean[0]; out[SBR_ILC_KLA5_MEAN * ld + i] = o.kla_mean[1];
            out[SBR_ILC_KLA8_MEAN * ld + i] = o.kla_mean[2];
        }
        if (status) status[i] = o.status;
        if (counters) { counters[i] = st.n_rhs; counters[ld + i] = st.n_rej; }
    }
    return 0;
}

int twin_ilc_update(int64_t n, int64_t ld, const SbrIlcLayout* lay, const double* w, const double* D, const double* sp6,
                    const double* so_mem, double* e_sum, double* e_last, double* u, double dt, double Kc, double tauI,
                    double tauD) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < 6; ++j) {
            const int off = lay->off[j], m = (j < 5 ? lay->off[j + 1] : lay->n_samples) - off;
            const int64_t base = (int64_t)off * ld + i;
            ilc_update_phase(m, lay->tp[j], sp6[j * ld + i], dt, w + off, D + off,
                             Column{const_cast<double*>(so_mem) + base, ld}, Column{e_sum + base, ld},
                             Column{e_last + base, ld}, Column{u + base, ld}, Kc, Kc / tauI, Kc * tauD);
        }
    return 0;
}

int twin_cycle_v2_traj(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                       const SbrParams* p, const SbrSchedule* s, const double* t_start, double* x_last, double* obs,
                       double* reward, double* traj, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 200; t.flags = 0;
    if (tol) t = *tol;
    for (int64_t i = 0; i < n; ++i) {
        double x[SBR_NX], a[3], load[SBR_NX], ts[SBR_NPHASE];
        for (int k = 0; k < SBR_NX; ++k) { x[k] = x0[k * ld + i]; load[k] = influent[k * ld + i]; }
        for (int k = 0; k < 3; ++k) a[k] = action[k * ld + i];
        for (int k = 0; k < SBR_NPHASE; ++k) ts[k] = t_start[k];
        Dp45State st;
        st.h = s->interval[0] / (double)s->n_sub[0]; st.n_rhs = 0; st.n_rej = 0;
        CycleOut o;
        if (mode == SBR_MODE_RK4) cycle_v2_traj<SBR_MODE_RK4>(x, a, Loading{load, 1}, load[0], *p, c, *s, t, st, o, ts, Column{traj + i, ld});
        else cycle_v2_traj<SBR_MODE_DP45>(x, a, Loading{load, 1}, load[0], *p, c, *s, t, st, o, ts, Column{traj + i, ld});
        for (int k = 0; k < SBR_NX; ++k) x_last[k * ld + i] = x[k];
        for (int k = 0; k < 3; ++k) obs[k * ld + i] = o.obs[k];
        reward[i] = o.reward;
    }
    return 0;
}

}  // extern "C"
