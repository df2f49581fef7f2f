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

using namespace sbr;

extern "C" {

int twin_cycle_v2(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                  const SbrParams* p, const SbrSchedule* s, double* x_last, double* obs, double* reward,
                  double* aux, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol) {
    const Coef c = make_coef(*p);
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 4000; t.reserved = 0;
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
        if (mode == SBR_MODE_RK4) cycle_v2<SBR_MODE_RK4>(x, a, L, load[0], *p, c, *s, t, st, o);
        else cycle_v2<SBR_MODE_DP45>(x, a, L, load[0], *p, c, *s, t, st, o);
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
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 4000; t.reserved = 0;
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
        if (tail == TAIL_REACT) rhs<TAIL_REACT>(xx, k, c, a);
        else if (tail == TAIL_FILL) rhs<TAIL_FILL>(xx, k, c, a);
        else rhs<TAIL_EC>(xx, k, c, a);
        for (int j = 0; j < SBR_NX; ++j) dx[j * ld + i] = k[j];
    }
    return 0;
}

}  // extern "C"
