// sbr_kernels.cu -- sm_100a kernels and the C ABI (include/sbr_b200.h) of the batched SBR stepper.
//
// Mapping: one environment per thread, whole state + PID integrator + stepper stages register-resident for the
// whole launch; global memory is touched only at the launch boundaries with coalesced SoA loads/stores
// (component c of env i at base[c*ld + i]).  The path is FP64-FMA-pipe bound (no contraction => no tensor
// cores); see DESIGN.md for the roofline arithmetic.  Per-env influent concentrations sit in a shared-memory
// column (conflict-free: consecutive threads -> consecutive 8-byte words) so the fill tail costs no registers.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "sbr_core.cuh"
#include "sbr_cnt.cuh"
#include "sbr_ilc.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return SBR_ERR_CUDA;
    }
    return SBR_OK;
}

constexpr int kBlock = 64;   // 2 warps per CTA: fine-grained tail balancing across 148 SMs x 4 SMSPs
#ifndef SBR_OS_BLOCK
#define SBR_OS_BLOCK 64      // threads per CTA of the interval-step kernel
#endif
constexpr int kOsBlock = SBR_OS_BLOCK;
// Resident CTAs per SM the interval-step kernels are compiled for.  RK4 keeps ~41 doubles live and runs best at
// 6 CTAs (168-register cap, 12 warps/SM); Dormand-Prince keeps six stage vectors live and spills 830 B per thread
// under that cap -- at 4 CTAs (255 registers, 8 warps/SM) it spills nothing and one env.step of 2^20 envs takes
// 0.237 ms instead of 0.268 ms (measured, profiles/r01f_ab_os_step_hoist_minblocks.log).
#ifndef SBR_OS_STEP_MINBLOCKS_RK4
#define SBR_OS_STEP_MINBLOCKS_RK4 6
#endif
#ifndef SBR_OS_STEP_MINBLOCKS_DP45
#define SBR_OS_STEP_MINBLOCKS_DP45 4
#endif
__host__ __device__ constexpr int os_step_minblocks(int mode) {
    return mode == SBR_MODE_DP45 ? SBR_OS_STEP_MINBLOCKS_DP45 : SBR_OS_STEP_MINBLOCKS_RK4;
}
#ifndef SBR_V4_STEP_MINBLOCKS_DP45
#define SBR_V4_STEP_MINBLOCKS_DP45 6
#endif
#ifndef SBR_CNT_STEP_MINBLOCKS_DP45
#define SBR_CNT_STEP_MINBLOCKS_DP45 6
#endif
__host__ __device__ constexpr int cnt_step_minblocks(int mode) {
    return mode == SBR_MODE_DP45 ? SBR_CNT_STEP_MINBLOCKS_DP45 : SBR_OS_STEP_MINBLOCKS_RK4;
}
// the fused rollout runs on fully sorted slots (little divergence left): spill-free code beats the third warp, as in
// sbr_os_step
#ifndef SBR_V4_FUSED_MINBLOCKS_DP45
#define SBR_V4_FUSED_MINBLOCKS_DP45 4
#endif
__host__ __device__ constexpr int v4_step_minblocks(int mode) {
    return mode == SBR_MODE_DP45 ? SBR_V4_STEP_MINBLOCKS_DP45 : SBR_OS_STEP_MINBLOCKS_RK4;
}

struct CycleArgs {
    int64_t n, ld;
    const double* x0;
    const double* influent;
    const double* action;
    double* x_last;
    double* obs;
    double* reward;
    double* aux;
    int32_t* status;
    uint32_t* counters;
    const int64_t* perm;
};

#ifndef SBR_CYCLE_DP45_MINBLOCKS
#define SBR_CYCLE_DP45_MINBLOCKS 1   // resident CTAs per SM the DP45 cycle kernel is compiled for
#endif
// RK4 runs without a register cap (184 registers, 10 warps/SM).  A 128-register build (16 warps/SM, spills outside
// the step loop) is 0.7 % faster at 2^20 envs but 19 % slower at 4096 envs (one warp per sub-partition: latency is what
// counts) and differs from this build in the last bits, which would make results depend on the batch size
// (profiles/r01f_ab_cycle_rk4_occupancy.log) -- not taken.
#ifdef SBR_CYCLE_MAXNREG        // A/B builds: an explicit register cap instead of the resident-CTA hint
#define SBR_CYCLE_BOUNDS __maxnreg__(SBR_CYCLE_MAXNREG)
#else
#define SBR_CYCLE_BOUNDS __launch_bounds__(kBlock, MODE == SBR_MODE_DP45 ? SBR_CYCLE_DP45_MINBLOCKS : 1)
#endif
template <int MODE>
__global__ void SBR_CYCLE_BOUNDS sbr_cycle_v2_kernel(CycleArgs g, SbrParams p, sbr::Coef c, SbrSchedule s, SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    // adaptive mode: per-env scratch column for parked stage vectors and per-phase KLa sums (sbr::Park)
    __shared__ double s_park[(MODE == SBR_MODE_DP45 ? sbr::PARK_SLOTS : 1) * kBlock];
    const int64_t slot = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (slot >= g.n) return;
    const int64_t i = g.perm ? g.perm[slot] : slot;      // divergence-aware ordering: see include/sbr_b200.h
    double x[SBR_NX], action[3];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.x0[k * g.ld + i];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.influent[k * g.ld + i];
#pragma unroll
    for (int k = 0; k < 3; ++k) action[k] = g.action[k * g.ld + i];
    sbr::Loading load{&s_load[threadIdx.x], kBlock};
    sbr::Dp45State st;
    st.h = s.interval[0] / (double)s.n_sub[0];
    st.n_rhs = 0;
    st.n_rej = 0;
    sbr::CycleOut o;
    sbr::cycle_v2<MODE>(x, action, load, load(0), p, c, s, tol, st, o, sbr::Park{&s_park[threadIdx.x], kBlock});
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.x_last[k * g.ld + i] = x[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) g.obs[k * g.ld + i] = o.obs[k];
    g.reward[i] = o.reward;
    if (g.aux) {
#pragma unroll
        for (int k = 0; k < SBR_AUX_ROWS; ++k) g.aux[k * g.ld + i] = o.aux[k];
    }
    if (g.status) g.status[i] = o.status;
    if (g.counters) {
        g.counters[i] = st.n_rhs;
        g.counters[g.ld + i] = st.n_rej;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Batch-to-batch (ILC) feed-forward path of SBR-v0 (sbr_ilc.cuh).  Sample memories are [S][ld]: sample j of env i
// at base[j * ld + i], so that the per-sample stores of a warp are one 256-byte line.
// ---------------------------------------------------------------------------------------------------------
struct IlcCycleArgs {
    int64_t n, ld;
    const double* x0;
    const double* influent;
    const double* sp;
    const double* kla_base;
    const double* u;
    double* so_mem;
    double* kla_mem;
    double* x_last;
    double* out;
    int32_t* status;
    uint32_t* counters;
    double t_fill;
};

template <int MODE>
__global__ void __launch_bounds__(kBlock) sbr_cycle_ilc_kernel(IlcCycleArgs g, SbrParams p, sbr::Coef c, SbrSchedule s,
                                                              SbrIlcLayout lay, SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.x0[k * g.ld + i];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.influent[k * g.ld + i];
    const double sp8[8] = {0.0, 0.0, g.sp[i], 0.0, g.sp[g.ld + i], 0.0, 0.0, g.sp[2 * g.ld + i]};
    sbr::Loading load{&s_load[threadIdx.x], kBlock};
    const bool ff = g.kla_base != nullptr;
    sbr::IlcIo io;
    io.so = sbr::Column{g.so_mem ? g.so_mem + i : nullptr, g.ld};
    io.kla_mem = sbr::Column{g.kla_mem ? g.kla_mem + i : nullptr, g.ld};
    io.kla_base = sbr::Column{ff ? const_cast<double*>(g.kla_base) + i : nullptr, g.ld};
    io.u = sbr::Column{ff ? const_cast<double*>(g.u) + i : nullptr, g.ld};
    const int off[6] = {lay.off[0], lay.off[1], lay.off[2], lay.off[3], lay.off[4], lay.off[5]};
    sbr::IlcOut o;
    sbr::Dp45State st;
    st.h = s.interval[0] / (double)s.n_sub[0];
    st.n_rhs = 0;
    st.n_rej = 0;
    sbr::cycle_ilc<MODE>(x, sp8, load, load(0), g.t_fill, ff, p, c, s, tol, st, io, off, o);
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.x_last[k * g.ld + i] = x[k];
    if (g.out) {
        g.out[SBR_ILC_QEFF * g.ld + i] = o.Qeff; g.out[SBR_ILC_QW * g.ld + i] = o.Qw;
        g.out[SBR_ILC_REWARD * g.ld + i] = o.reward; g.out[SBR_ILC_OCI * g.ld + i] = o.OCI;
        g.out[SBR_ILC_KLA3_MEAN * g.ld + i] = o.kla_mean[0]; g.out[SBR_ILC_KLA5_MEAN * g.ld + i] = o.kla_mean[1];
        g.out[SBR_ILC_KLA8_MEAN * g.ld + i] = o.kla_mean[2];
    }
    if (g.status) g.status[i] = o.status;
    if (g.counters) {
        g.counters[i] = st.n_rhs;
        g.counters[g.ld + i] = st.n_rej;
    }
}

struct IlcUpdateArgs {
    int64_t n, ld;
    const double* w;
    const double* D;
    const double* sp6;
    const double* so_mem;
    double* e_sum;
    double* e_last;
    double* u;
    double dt, Kc, KcI, KcD;
};

// one thread per (env, phase): blockIdx.y = phase
__global__ void __launch_bounds__(128) sbr_ilc_update_kernel(IlcUpdateArgs g, SbrIlcLayout lay) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= g.n) return;
    const int j = blockIdx.y;
    const int off = lay.off[j], n = (j < 5 ? lay.off[j + 1] : lay.n_samples) - off;
    const int64_t base = (int64_t)off * g.ld + i;
    sbr::ilc_update_phase(n, lay.tp[j], g.sp6[j * g.ld + i], g.dt, g.w + off, g.D + off,
                          sbr::Column{const_cast<double*>(g.so_mem) + base, g.ld}, sbr::Column{g.e_sum + base, g.ld},
                          sbr::Column{g.e_last + base, g.ld}, sbr::Column{g.u + base, g.ld}, g.Kc, g.KcI, g.KcD);
}

struct CycleTrajArgs {
    CycleArgs c;
    double* traj;
    double t_start[SBR_NPHASE];
};

template <int MODE>
__global__ void __launch_bounds__(kBlock) sbr_cycle_v2_traj_kernel(CycleTrajArgs h, SbrParams p, sbr::Coef c, SbrSchedule s,
                                                                  SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    const CycleArgs& g = h.c;
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    double x[SBR_NX], action[3];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.x0[k * g.ld + i];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.influent[k * g.ld + i];
#pragma unroll
    for (int k = 0; k < 3; ++k) action[k] = g.action[k * g.ld + i];
    sbr::Loading load{&s_load[threadIdx.x], kBlock};
    sbr::Dp45State st;
    st.h = s.interval[0] / (double)s.n_sub[0];
    st.n_rhs = 0;
    st.n_rej = 0;
    sbr::CycleOut o;
    double ts[SBR_NPHASE];
#pragma unroll
    for (int k = 0; k < SBR_NPHASE; ++k) ts[k] = h.t_start[k];
    sbr::cycle_v2_traj<MODE>(x, action, load, load(0), p, c, s, tol, st, o, ts, sbr::Column{h.traj + i, g.ld});
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.x_last[k * g.ld + i] = x[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) g.obs[k * g.ld + i] = o.obs[k];
    g.reward[i] = o.reward;
    if (g.aux) {
#pragma unroll
        for (int k = 0; k < SBR_AUX_ROWS; ++k) g.aux[k * g.ld + i] = o.aux[k];
    }
    if (g.status) g.status[i] = o.status;
    if (g.counters) {
        g.counters[i] = st.n_rhs;
        g.counters[g.ld + i] = st.n_rej;
    }
}

struct IntervalArgs {
    int64_t n, ld;
    double* x;
    const double* kla;
    const double* ec;
    const double* loading;
    uint32_t* counters;
    double T;
    int n_sub;
};

template <int TAIL, int MODE>
__global__ void __launch_bounds__(kBlock) sbr_interval_kernel(IntervalArgs g, SbrParams p, sbr::Coef c, SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.x[k * g.ld + i];
    sbr::TailArgs a;
    a.kla = g.kla[i];
    a.q = 0.0;
    a.ec_conc = p.ec_conc;
    a.load = sbr::Loading{&s_load[threadIdx.x], kBlock};
    if (TAIL == sbr::TAIL_FILL) {
#pragma unroll
        for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.loading[k * g.ld + i];
        a.q = a.load(0);
    }
    if (TAIL == sbr::TAIL_EC) a.q = g.ec[i];
    sbr::Dp45State st;
    st.h = g.T / (double)g.n_sub;
    st.n_rhs = 0;
    st.n_rej = 0;
    sbr::integrate_interval<TAIL, MODE>(x, g.T, g.n_sub, c, a, tol, st);
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.x[k * g.ld + i] = x[k];
    if (g.counters) {
        g.counters[i] = st.n_rhs;
        g.counters[g.ld + i] = st.n_rej;
    }
}

struct RhsArgs {
    int64_t n, ld;
    const double* x;
    const double* kla;
    const double* ec;
    const double* loading;
    double* dx;
};

template <int TAIL>
__global__ void __launch_bounds__(kBlock) sbr_rhs_kernel(RhsArgs g, SbrParams p, sbr::Coef c) {
    __shared__ double s_load[SBR_NX * kBlock];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    double x[SBR_NX], k[SBR_NX];
#pragma unroll
    for (int j = 0; j < SBR_NX; ++j) { x[j] = g.x[j * g.ld + i]; k[j] = 0.0; }
    sbr::TailArgs a;
    a.kla = g.kla[i];
    a.q = 0.0;
    a.ec_conc = p.ec_conc;
    a.load = sbr::Loading{&s_load[threadIdx.x], kBlock};
    if (TAIL == sbr::TAIL_FILL) {
#pragma unroll
        for (int j = 0; j < SBR_NX; ++j) s_load[j * kBlock + threadIdx.x] = g.loading[j * g.ld + i];
        a.q = a.load(0);
    }
    if (TAIL == sbr::TAIL_EC) a.q = g.ec[i];
    a.kla_sat = a.kla * c.so_sat;
    sbr::rhs<TAIL>(x, k, c, a);
#pragma unroll
    for (int j = 0; j < SBR_NX; ++j) g.dx[j * g.ld + i] = k[j];
}

// ---------------------------------------------------------------------------------------------------------
// Path B kernels: persistent per-env state st[SBR_OS_ROWS][ld] (SoA), one env per thread.
// ---------------------------------------------------------------------------------------------------------
struct OsArgs {
    int64_t n, ld;
    double* st;
    const double* x0;         // reset only (may be NULL)
    const double* influent;   // reset only
    const uint8_t* mask;      // reset only (may be NULL)
    double* obs_do;
    double* obs_ec;
    uint8_t* done;
    int32_t* status;
    uint32_t* counters;
};

__constant__ double c_x0_init[SBR_NX] = {   // gym_SBR_oneshot.py:201-203
    0.6161484733495801, 30, 0.571098000538576, 1440.01157895393, 31.254221999137, 2599.2714348941,
    168.915006750837, 551.901552960823, 2.16607843793004, 13.3791460027604, 0.00562880208518134,
    0.35996687629947, 1.86916737961228, 3.790463057094611};

__device__ __forceinline__ void os_store_ctrl(double* st, int64_t ld, int64_t i, const sbr::OsCtrl& c, double h) {
    st[SBR_OS_T * ld + i] = c.t;
    st[SBR_OS_SO_PREV * ld + i] = c.so_prev;
    st[SBR_OS_SNO_LAST * ld + i] = c.sno_last;
    st[SBR_OS_SNO_PREV * ld + i] = c.sno_prev;
    st[SBR_OS_IE_DO * ld + i] = c.ie_do;
    st[SBR_OS_IE_EC * ld + i] = c.ie_ec;
    st[SBR_OS_EC_LAST * ld + i] = c.ec_last;
    st[SBR_OS_H * ld + i] = h;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) sbr_os_reset_kernel(OsArgs g, SbrParams p, sbr::Coef c, SbrOsSchedule s,
                                                              SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    if (g.mask && g.mask[i] == 0) return;
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.x0 ? g.x0[k * g.ld + i] : c_x0_init[k];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.influent[k * g.ld + i];
    sbr::Loading load{&s_load[threadIdx.x], kBlock};
    sbr::Dp45State dp;
    dp.h = s.t_fill / (double)(s.fill_pts > 1 ? s.fill_pts - 1 : 1);
    dp.n_rhs = 0; dp.n_rej = 0;
    sbr::OsCtrl ctl;
    const sbr::Column ring{g.st + SBR_OS_KLA_RING * g.ld + i, g.ld};
    const sbr::Column od{g.obs_do + i, g.ld}, oe{g.obs_ec + i, g.ld};
    int status = sbr::os_reset_env<MODE>(x, load, p, c, s, tol, dp, ctl, ring, od, oe);
    bool finite = true;
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) { g.st[k * g.ld + i] = x[k]; finite = finite && (fabs(x[k]) < 1e300); }
    if (!finite) status |= SBR_ST_NONFINITE;
    os_store_ctrl(g.st, g.ld, i, ctl, s.t_delta / 9.0);
    g.st[SBR_OS_RETURN * g.ld + i] = 0.0;
    g.st[SBR_OS_STEPS * g.ld + i] = 0.0;
    g.st[SBR_OS_QW * g.ld + i] = NAN;
    g.done[i] = 0;
    if (g.status) g.status[i] = status;
    if (g.counters) { g.counters[i] = dp.n_rhs; g.counters[g.ld + i] = dp.n_rej; }
}

// Touch the 128-byte lines of a row that this warp will read later, so that the later read is an L2 hit instead of
// an exposed DRAM round trip.  No destination register, no scoreboard entry.
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

// ---- TMA tile loads global -> shared completing on an mbarrier ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const CUtensorMap* tm, int c0, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Policy head of the rollout path (BASELINE configs[4]: observations -> policy -> actions, every step): a two-layer
// perceptron per env, action = lo + span * sigmoid(W2 tanh(W1 obs)), in ONE launch on the kernels' own SoA layout.
// The torch expression of the same policy is nine launches that move ~1.3 GB per step at 2^20 envs -- as much time as
// the env.step it feeds; this reads the 144 B of observation and writes the 16 B of action per env.  FP32 arithmetic
// (the policy's dtype), weights broadcast from shared memory.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPolMaxIn = 40, kPolMaxHidden = 64, kPolMaxOut = 4;
struct PolicyArgs {
    int64_t n, ld;
    const double* obs_a;
    const double* obs_b;      // may be NULL
    const float* w1;          // [hidden][rows_a + rows_b]
    const float* w2;          // [n_out][hidden]
    const float* lo;          // [n_out]
    const float* span;        // [n_out]
    double* action;           // [n_out][ld]
    int rows_a, rows_b, hidden, n_out;
};

// The perceptron itself, shared by the stand-alone kernel and the fused rollout (sbr_os_rollout_k) so that both produce
// the same bits: explicit fmaf chains in input order, tanh through ex2 / rcp, 1 / (1 + expf(-y)).
__device__ __forceinline__ void policy_eval(const float* s_w1, const float* s_w2, const float* lo, const float* span,
                                            int n_in, int hidden, int n_out, const float (&x)[kPolMaxIn],
                                            float (&out)[kPolMaxOut]) {
    float y[kPolMaxOut] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int h = 0; h < hidden; ++h) {
        const float* w = &s_w1[h * n_in];
        float acc = 0.0f;
#pragma unroll
        for (int r = 0; r < kPolMaxIn; ++r)
            if (r < n_in) acc = fmaf(w[r], x[r], acc);
        // tanh(a) = 1 - 2 / (exp(2a) + 1) on the MUFU ex2 / rcp units (6 instructions, absolute error ~2e-7; tanhf is ~16
        // with a branch): the head runs once per env and step inside the fused rollout, where it was 11 % of the instructions
        const float t = 1.0f - __fdividef(2.0f, __expf(2.0f * acc) + 1.0f);
#pragma unroll
        for (int o = 0; o < kPolMaxOut; ++o)
            if (o < n_out) y[o] = fmaf(s_w2[o * hidden + h], t, y[o]);
    }
#pragma unroll
    for (int o = 0; o < kPolMaxOut; ++o) {
        out[o] = 0.0f;
        if (o < n_out) {
            const float sg = 1.0f / (1.0f + expf(-y[o]));
            out[o] = fmaf(span[o], sg, lo[o]);
        }
    }
}

__global__ void __launch_bounds__(128) sbr_policy_mlp_kernel(PolicyArgs g) {
    __shared__ float s_w1[kPolMaxHidden * kPolMaxIn];
    __shared__ float s_w2[kPolMaxOut * kPolMaxHidden];
    const int n_in = g.rows_a + g.rows_b;
    for (int k = threadIdx.x; k < g.hidden * n_in; k += blockDim.x) s_w1[k] = g.w1[k];
    for (int k = threadIdx.x; k < g.n_out * g.hidden; k += blockDim.x) s_w2[k] = g.w2[k];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    float x[kPolMaxIn];
#pragma unroll
    for (int r = 0; r < kPolMaxIn; ++r) {
        if (r < g.rows_a) x[r] = (float)g.obs_a[(int64_t)r * g.ld + i];
        else if (r < n_in) x[r] = (float)g.obs_b[(int64_t)(r - g.rows_a) * g.ld + i];
        else x[r] = 0.0f;
    }
    float out[kPolMaxOut];
    policy_eval(s_w1, s_w2, g.lo, g.span, n_in, g.hidden, g.n_out, x, out);
#pragma unroll
    for (int o = 0; o < kPolMaxOut; ++o)
        if (o < g.n_out) g.action[(int64_t)o * g.ld + i] = (double)out[o];
}

// ---------------------------------------------------------------------------------------------------------
// sbr_os_step / sbr_os_step_k: persistent warps, one tile of 32 consecutive envs at a time.
//
// One env.step is ~1.5 k FP64 instructions on ~0.65 kB of state: arithmetic intensity sits at the ridge, and with
// 254 registers per thread only two warps per scheduler are resident -- too few to hide a DRAM round trip behind
// another warp's arithmetic (round 1: 58 % of HBM peak, 41 % of the FP64 pipe, stalled on long_scoreboard).  So the
// loads are taken off the warps: while a warp computes tile k, the TMA unit copies the rows of its NEXT tile -- a
// [34 rows x 32 envs] box of st, the step's [2 x 32] action box and the 32 done flags: three tensor-map copies issued
// by one lane, completing on the warp's own mbarrier -- into the other half of the warp's double buffer.  Every warp
// runs its own pipeline (no CTA-wide barrier couples the two warps of a CTA); the lanes read their state from shared
// memory (conflict-free: consecutive lanes, consecutive words) and write results straight to global memory
// (coalesced stores do not stall a warp).  TMA zero-fills the ragged last tile.  Buffers that are not 16-byte aligned
// (or an odd ld) take a plain-load fill of the stage instead.
// ---------------------------------------------------------------------------------------------------------
struct OsStepArgs {
    int64_t n, ld, num_tiles;
    double* st;
    double* action;           // [K][2][ld] in; fused rollout: [2][ld] in/out
    double* obs_do;           // may be NULL
    double* obs_ec;           // may be NULL
    double* state;            // may be NULL
    double* reward;           // [K][ld]
    uint8_t* done;
    int32_t* status;          // may be NULL
    uint32_t* counters;       // may be NULL
    double* traj;             // may be NULL: [traj_cap][SBR_TRAJ_ROWS][ld]
    int K, tma, traj_cap;
    // fused rollout (sbr_os_rollout_k): the policy head evaluated in-kernel between the K steps (NULL = off)
    const float* pol_w1;      // [hidden][18]
    const float* pol_w2;      // [2][hidden]
    const float* pol_lo;      // [2]
    const float* pol_span;    // [2]
    int pol_hidden;
    double* act_log;          // may be NULL: [K][2][ld] the set-points each step ran with
    double* obs_log;          // may be NULL: [K][18][ld] the observation after each step
};

constexpr int kOsTile = 32;                        // envs per tile = one warp
constexpr int kOsWarps = kOsBlock / 32;
constexpr int kOsStageRows = SBR_OS_QW + 2;        // st rows 0..SBR_OS_QW-1, then the first step's two action rows
constexpr int kOsRowAct = SBR_OS_QW;
constexpr uint32_t kOsStageBytes = kOsStageRows * kOsTile * sizeof(double) + kOsTile;

template <int MODE>
__global__ void __launch_bounds__(kOsBlock, os_step_minblocks(MODE)) sbr_os_step_kernel(
        OsStepArgs g, SbrParams p, sbr::Coef c, SbrOsSchedule s, SbrTol tol, const __grid_constant__ CUtensorMap tm_st,
        const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_done) {
    __shared__ __align__(128) double s_stage[kOsWarps][2][kOsStageRows * kOsTile];
    __shared__ __align__(128) uint8_t s_done[kOsWarps][2][128];
    __shared__ __align__(8) uint64_t s_bar[kOsWarps][2];
    __shared__ float s_pw1[kPolMaxHidden * 2 * SBR_OS_NOBS];
    __shared__ float s_pw2[2 * kPolMaxHidden];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (g.pol_w1) {
        for (int k = threadIdx.x; k < g.pol_hidden * 2 * SBR_OS_NOBS; k += kOsBlock) s_pw1[k] = g.pol_w1[k];
        for (int k = threadIdx.x; k < 2 * g.pol_hidden; k += kOsBlock) s_pw2[k] = g.pol_w2[k];
        __syncthreads();
    }
    if (lane == 0) {
        mbar_init(&s_bar[warp][0], 1);
        mbar_init(&s_bar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t bar0 = smem_u32(&s_bar[warp][0]);

    // fill the warp's stage `stg` with tile `tile`: three TMA box copies, or plain loads of the lane's own column
    auto fill = [&](int64_t tile, int stg) {
        const int64_t i0 = tile * kOsTile;
        double* sg = s_stage[warp][stg];
        if (g.tma) {
            if (lane == 0) {
                const uint32_t bar = bar0 + 8u * stg;
                mbar_expect_tx(bar, kOsStageBytes);
                tma_load_2d(smem_u32(sg), &tm_st, (int)i0, 0, bar);
                tma_load_2d(smem_u32(sg + kOsRowAct * kOsTile), &tm_act, (int)i0, 0, bar);
                tma_load_1d(smem_u32(s_done[warp][stg]), &tm_done, (int)i0, bar);
            }
        } else if (i0 + lane < g.n) {
            const int64_t i = i0 + lane;
#pragma unroll
            for (int r = 0; r < SBR_OS_QW; ++r) sg[r * kOsTile + lane] = g.st[r * g.ld + i];
            sg[kOsRowAct * kOsTile + lane] = g.action[i];
            sg[(kOsRowAct + 1) * kOsTile + lane] = g.action[g.ld + i];
            s_done[warp][stg][lane] = g.done[i];
        }
    };

    const int64_t stride = (int64_t)gridDim.x * kOsWarps;
    int64_t tile = (int64_t)blockIdx.x * kOsWarps + warp;
    uint32_t parity = 0u;                            // bit stg = phase parity the warp waits for next on that stage
    int stg = 0;
    if (tile < g.num_tiles) fill(tile, 0);
    for (; tile < g.num_tiles; tile += stride, stg ^= 1) {
        // the other stage was released at the end of the previous iteration: prefetch the next tile into it
        if (tile + stride < g.num_tiles) fill(tile + stride, stg ^ 1);
        if (g.tma) {
            mbar_wait(bar0 + 8u * stg, (parity >> stg) & 1u);
            parity ^= 1u << stg;
        }
        const int64_t i = tile * kOsTile + lane;
        if (i < g.n) {
            double* sg = s_stage[warp][stg] + lane;
            double x[SBR_NX];
#pragma unroll
            for (int k = 0; k < SBR_NX; ++k) x[k] = sg[k * kOsTile];
            sbr::OsCtrl ctl;
            ctl.t = sg[SBR_OS_T * kOsTile];
            ctl.so_prev = sg[SBR_OS_SO_PREV * kOsTile];
            ctl.sno_last = sg[SBR_OS_SNO_LAST * kOsTile];
            ctl.sno_prev = sg[SBR_OS_SNO_PREV * kOsTile];
            ctl.ie_do = sg[SBR_OS_IE_DO * kOsTile];
            ctl.ie_ec = sg[SBR_OS_IE_EC * kOsTile];
            ctl.ec_last = sg[SBR_OS_EC_LAST * kOsTile];
            sbr::Dp45State dp;
            dp.h = sg[SBR_OS_H * kOsTile];
            dp.n_rhs = 0; dp.n_rej = 0;
            sbr::KlaRing ring{sbr::Column{sg + SBR_OS_KLA_RING * kOsTile, kOsTile},
                              sbr::Column{g.st + SBR_OS_KLA_RING * g.ld + i, g.ld}, sbr::os_ring_head(ctl.t, s)};
            ctl.kla_last = ring.back(1);
            bool is_done = s_done[warp][stg][lane] != 0;
            const bool was_done = is_done;
            int status = was_done ? SBR_ST_DONE : 0;
            double qw = NAN;
            const sbr::Column od{g.obs_do ? g.obs_do + i : nullptr, g.ld}, oe{g.obs_ec ? g.obs_ec + i : nullptr, g.ld},
                os{g.state ? g.state + i : nullptr, g.ld};
            double a_do = 0.0, a_ec = 0.0;
            if (g.K > 1 && !g.pol_w1) {
                for (int k = 1; k < g.K; ++k) {
                    prefetch_l2(g.action + (int64_t)(2 * k) * g.ld + i);
                    prefetch_l2(g.action + (int64_t)(2 * k + 1) * g.ld + i);
                }
            }
            for (int k = 0; k < g.K; ++k) {
                if (is_done) {
                    // stepping a finished episode is a no-op: reward 0 (and, if the launch starts on a finished env,
                    // the same observation with zero deltas and status SBR_ST_DONE)
                    g.reward[(int64_t)k * g.ld + i] = 0.0;
                    continue;
                }
                if (k == 0) { a_do = sg[kOsRowAct * kOsTile]; a_ec = sg[(kOsRowAct + 1) * kOsTile]; }
                else if (!g.pol_w1) { a_do = g.action[(int64_t)(2 * k) * g.ld + i]; a_ec = g.action[(int64_t)(2 * k + 1) * g.ld + i]; }
                if (g.act_log) { g.act_log[(int64_t)(2 * k) * g.ld + i] = a_do; g.act_log[(int64_t)(2 * k + 1) * g.ld + i] = a_ec; }
                sbr::OsStepOut o;
                sbr::os_step_env<MODE>(x, ctl, ring, a_do, a_ec, p, c, s, tol, dp, o,
                                       sbr::OsTraj{g.traj ? g.traj + i : nullptr, g.ld, g.traj_cap});
                g.reward[(int64_t)k * g.ld + i] = o.reward;
                // episode return and step count accumulate in their staged slots (no registers across the stepper),
                // step by step, so that K steps in one launch round exactly like K launches
                sg[SBR_OS_RETURN * kOsTile] += o.reward;
                sg[SBR_OS_STEPS * kOsTile] += 1.0;
                status |= o.status;
                if (o.done) { is_done = true; qw = o.Qw; }
                // the observation of the last step that ran (the terminal one if the episode ends inside the launch)
                if (k == g.K - 1 || o.done) sbr::os_emit_obs(ctl.t, x, o.first, od, oe, os);
                if (g.pol_w1) {
                    // fused rollout: the set-points of the NEXT step from this step's observation, evaluated on the
                    // registers that hold the state (same arithmetic as os_emit_obs -> sbr_policy_mlp, hence same bits)
                    double ob[2 * SBR_OS_NOBS];
                    sbr::os_emit_obs(ctl.t, x, o.first, sbr::Column{ob, 1}, sbr::Column{ob + SBR_OS_NOBS, 1},
                                     sbr::Column{nullptr, 1});
                    if (g.obs_log) {
#pragma unroll
                        for (int r = 0; r < 2 * SBR_OS_NOBS; ++r)
                            g.obs_log[((int64_t)k * 2 * SBR_OS_NOBS + r) * g.ld + i] = ob[r];
                    }
                    float xin[kPolMaxIn], act[kPolMaxOut];
#pragma unroll
                    for (int r = 0; r < kPolMaxIn; ++r) xin[r] = r < 2 * SBR_OS_NOBS ? (float)ob[r] : 0.0f;
                    policy_eval(s_pw1, s_pw2, g.pol_lo, g.pol_span, 2 * SBR_OS_NOBS, g.pol_hidden, 2, xin, act);
                    a_do = (double)act[0]; a_ec = (double)act[1];
                }
            }
            // fused rollout: the first step of the next launch finds its set-points where this launch found its own
            if (g.pol_w1 && !was_done) { g.action[i] = a_do; g.action[g.ld + i] = a_ec; }
            if (was_done) {
                sbr::os_emit_obs(ctl.t, x, sbr::obs_ref(x), od, oe, os);
            } else {
#pragma unroll
                for (int k = 0; k < SBR_NX; ++k) g.st[k * g.ld + i] = x[k];
                os_store_ctrl(g.st, g.ld, i, ctl, dp.h);
                g.st[SBR_OS_RETURN * g.ld + i] = sg[SBR_OS_RETURN * kOsTile];
                g.st[SBR_OS_STEPS * g.ld + i] = sg[SBR_OS_STEPS * kOsTile];
                if (is_done) { g.st[SBR_OS_QW * g.ld + i] = qw; g.done[i] = 1; }
            }
            if (g.status) g.status[i] = status;
            if (g.counters) { g.counters[i] = dp.n_rhs; g.counters[g.ld + i] = dp.n_rej; }
        }
        __syncwarp();             // every lane is done with stage `stg`: the next iteration refills it
    }
}

// Persistent grid of the interval-step kernel: resident CTAs per SM (occupancy query, cached per device) x SMs.
template <int MODE>
static int os_step_grid(int64_t num_tiles) {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int per_sm = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sbr_os_step_kernel<MODE>, kOsBlock, 0);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    }
    const int64_t ctas = (num_tiles + kOsWarps - 1) / kOsWarps;
    return (int)(ctas < (int64_t)cached[dev] ? ctas : (int64_t)cached[dev]);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
        cudaGetLastError();
    }
    return fn;
}

// Tensor map of a row-major [rows][ld] array seen as {inner = n envs, outer = rows}, box = {box_cols, box_rows}.
static bool make_map_2d(CUtensorMap* tm, CUtensorMapDataType dt, size_t elem, const void* base, int64_t n, int64_t rows,
                        int64_t ld, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * elem};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    return fn(tm, dt, rows > 1 ? 2 : 1, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------
// SBR-v4 kernels: persistent per-env state st[SBR_V4_ROWS][ld] + the env's influent column [14][ld].
// ---------------------------------------------------------------------------------------------------------
struct V4Args {
    int64_t n, ld;
    double* st;
    const double* x0;         // reset only (may be NULL)
    const double* influent;
    const uint8_t* mask;      // reset only (may be NULL)
    const double* action;     // step only
    double* obs;
    double* reward;           // step only
    uint8_t* done;
    int32_t* status;
    uint32_t* counters;
    const int32_t* order;     // may be NULL: slot j of st holds env order[j]; every other buffer is indexed by env
    // fused rollout (sbr_v4_rollout_k): K steps per launch, the policy head evaluated in-kernel between them (NULL = off)
    int K;
    const float* pol_w1;      // [hidden][14]
    const float* pol_w2;      // [1][hidden]
    const float* pol_lo;      // [1]
    const float* pol_span;    // [1]
    int pol_hidden;
    double* action_io;        // fused rollout: [n] in/out
    double* act_log;          // may be NULL: [K][ld]
    double* obs_log;          // may be NULL: [K][14][ld]
};

// `order` (divergence-aware placement): the persistent state st is kept in SLOT order -- slots sorted by how many
// steps the env needed recently, so that the 32 envs of a warp finish together -- while everything the caller sees
// (action, observation, reward, done, status, counters, influent, x0, mask) stays indexed by env.  j = thread's slot,
// i = the env it works on.
__global__ void __launch_bounds__(128) sbr_v4_reset_kernel(V4Args g, SbrParams p) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    const int64_t i = g.order ? (int64_t)g.order[j] : j;
    if (g.mask && g.mask[i] == 0) return;
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) { x[k] = g.x0 ? g.x0[k * g.ld + i] : c_x0_init[k]; g.st[k * g.ld + j] = x[k]; }
    const sbr::Loading load{g.influent + i, (int)g.ld};
    sbr::v4_reset_obs(x, load, p, sbr::Column{g.obs + i, g.ld});
#pragma unroll
    for (int r = SBR_V4_T; r < SBR_V4_ROWS; ++r) g.st[r * g.ld + j] = 0.0;
    g.st[SBR_V4_QW * g.ld + j] = NAN;
    g.done[i] = 0;
}

// FUSED = false: sbr_v4_step (K = 1, no policy code in the kernel); FUSED = true: sbr_v4_rollout_k.
template <int MODE, bool FUSED>
__global__ void __launch_bounds__(kBlock, (FUSED && MODE == SBR_MODE_DP45) ? SBR_V4_FUSED_MINBLOCKS_DP45 : v4_step_minblocks(MODE))
sbr_v4_step_kernel(V4Args g, SbrParams p, sbr::Coef c, SbrOsSchedule s,
                                                                SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    __shared__ float s_pw1[FUSED ? kPolMaxHidden * SBR_NX : 1];
    __shared__ float s_pw2[FUSED ? kPolMaxHidden : 1];
    if (FUSED) {
        for (int k = threadIdx.x; k < g.pol_hidden * SBR_NX; k += kBlock) s_pw1[k] = g.pol_w1[k];
        for (int k = threadIdx.x; k < g.pol_hidden; k += kBlock) s_pw2[k] = g.pol_w2[k];
        __syncthreads();
    }
    const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;          // slot of the state
    if (j >= g.n) return;
    const int64_t i = g.order ? (int64_t)g.order[j] : j;                   // env (see sbr_v4_reset_kernel)
    // every load of the launch is issued before the first use (one DRAM round trip, as in sbr_os_step_kernel)
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.st[k * g.ld + j];
    const uint8_t was_done = g.done[i];
    sbr::V4Ctrl ctl;
    ctl.t = g.st[SBR_V4_T * g.ld + j];
    ctl.u = g.st[SBR_V4_U * g.ld + j];
    ctl.so_prev = g.st[SBR_V4_SO_PREV * g.ld + j];
    ctl.ie = g.st[SBR_V4_IE * g.ld + j];
    ctl.kla_last = g.st[SBR_V4_KLA_LAST * g.ld + j];
    ctl.kla_sum = g.st[SBR_V4_KLA_SUM * g.ld + j];
    sbr::Dp45State dp;
    dp.h = g.st[SBR_V4_H * g.ld + j];
    double action = FUSED ? g.action_io[i] : g.action[i];
    double ret = g.st[SBR_V4_RETURN * g.ld + j], steps = g.st[SBR_V4_STEPS * g.ld + j];
    const sbr::Column ob{g.obs ? g.obs + i : nullptr, g.ld};
    if (was_done) {
        // stepping a finished episode is a no-op: same observation, reward 0
#pragma unroll
        for (int k = 0; k < SBR_NX; ++k) ob.set(k, x[k] * sbr::inv_x1_v4(k));
        for (int k = 0; k < (FUSED ? g.K : 1); ++k) g.reward[(int64_t)k * g.ld + i] = 0.0;
        if (g.status) g.status[i] = SBR_ST_DONE;
        if (g.counters) { g.counters[i] = 0; g.counters[g.ld + i] = 0; }
        return;
    }
    if (!(dp.h > 0.0)) dp.h = s.t_delta / 9.0;
    dp.n_rhs = 0; dp.n_rej = 0;
    if (ctl.t < s.t_fill) {          // the influent column is only read while the reactor fills (26 of 493 steps)
#pragma unroll
        for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.influent[k * g.ld + i];
    }
    const sbr::Loading load{&s_load[threadIdx.x], kBlock};
    int status = 0;
    bool is_done = false;
    double qw = NAN;
    for (int k = 0; k < (FUSED ? g.K : 1); ++k) {
        if (is_done) { g.reward[(int64_t)k * g.ld + i] = 0.0; continue; }
        if (FUSED && g.act_log) g.act_log[(int64_t)k * g.ld + i] = action;
        sbr::V4Out o;
        // the observation of the last step that ran goes to `obs`; the steps before it only feed the policy head
        double obl[FUSED ? SBR_NX : 1];
        const bool last = !FUSED || k == g.K - 1;
        const sbr::Column oc = (FUSED && !last) ? sbr::Column{obl, 1} : ob;
        sbr::v4_step_env<MODE>(x, ctl, action, load, p, c, s, tol, dp, oc, o);
        g.reward[(int64_t)k * g.ld + i] = o.reward;
        ret += o.reward;
        steps += 1.0;
        status |= o.status;
        if (o.done) {
            is_done = true; qw = o.Qw;
            if (FUSED && !last) {
#pragma unroll
                for (int r = 0; r < SBR_NX; ++r) ob.set(r, obl[FUSED ? r : 0]);
            }
        }
        if (FUSED) {
            // fused rollout: the set-point change of the NEXT step from this step's observation x / x_1 (same arithmetic
            // as v4_step_env -> sbr_policy_mlp, hence the same bits)
            float xin[kPolMaxIn], act[kPolMaxOut];
#pragma unroll
            for (int r = 0; r < kPolMaxIn; ++r) xin[r] = 0.0f;
#pragma unroll
            for (int r = 0; r < SBR_NX; ++r) xin[r] = (float)(x[r] * sbr::inv_x1_v4(r));
            if (g.obs_log) {
#pragma unroll
                for (int r = 0; r < SBR_NX; ++r) g.obs_log[((int64_t)k * SBR_NX + r) * g.ld + i] = x[r] * sbr::inv_x1_v4(r);
            }
            policy_eval(s_pw1, s_pw2, g.pol_lo, g.pol_span, SBR_NX, g.pol_hidden, 1, xin, act);
            action = (double)act[0];
        }
    }
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.st[k * g.ld + j] = x[k];
    g.st[SBR_V4_T * g.ld + j] = ctl.t;
    g.st[SBR_V4_U * g.ld + j] = ctl.u;
    g.st[SBR_V4_SO_PREV * g.ld + j] = ctl.so_prev;
    g.st[SBR_V4_IE * g.ld + j] = ctl.ie;
    g.st[SBR_V4_KLA_LAST * g.ld + j] = ctl.kla_last;
    g.st[SBR_V4_KLA_SUM * g.ld + j] = ctl.kla_sum;
    g.st[SBR_V4_H * g.ld + j] = dp.h;
    g.st[SBR_V4_RETURN * g.ld + j] = ret;
    g.st[SBR_V4_STEPS * g.ld + j] = steps;
    if (is_done) { g.st[SBR_V4_QW * g.ld + j] = qw; g.done[i] = 1; }
    if (FUSED) g.action_io[i] = action;             // the first step of the next launch finds its action here
    if (g.status) g.status[i] = status;
    if (g.counters) { g.counters[i] = dp.n_rhs; g.counters[g.ld + i] = dp.n_rej; }
}

// ---------------------------------------------------------------------------------------------------------
// SBRCnt-v0/1/2, SBRCntMA-v1, SBROS-v2 (sbr_cnt.cuh): persistent per-env state st[SBR_CNT_ROWS][ld].  One env per
// thread, every load issued before the first use; the kind is uniform over the launch.
// ---------------------------------------------------------------------------------------------------------
struct CntArgs {
    int64_t n, ld;
    double* st;
    const double* x0;         // reset only (may be NULL)
    const double* influent;   // reset only
    const uint8_t* mask;      // reset only (may be NULL)
    const double* action;     // step only: [2][ld]
    double* obs;              // may be NULL in step
    double* reward;           // step only ([K][ld] in the fused rollout)
    uint8_t* done;
    int32_t* status;
    uint32_t* counters;
    // fused rollout (sbr_cnt_rollout_k): K steps per launch, the policy head evaluated in-kernel between them
    int K;
    const float* pol_w1;      // [hidden][n_in]
    const float* pol_w2;      // [n_out][hidden]
    const float* pol_lo;      // [n_out]
    const float* pol_span;    // [n_out]
    int pol_hidden;
    double* action_io;        // fused rollout: [2][ld] in/out
    double* act_log;          // may be NULL: [K][2][ld]
    double* obs_log;          // may be NULL: [K][n_in][ld]
};

__device__ __forceinline__ void cnt_load_ctrl(const CntArgs& g, int64_t i, sbr::CntCtrl& c, double& h) {
    c.t = g.st[SBR_CNT_T * g.ld + i];
    c.u_do = g.st[SBR_CNT_U_DO * g.ld + i];
    c.u_ec = g.st[SBR_CNT_U_EC * g.ld + i];
    c.so_prev = g.st[SBR_CNT_SO_PREV * g.ld + i];
    c.cv_last = g.st[SBR_CNT_CV_LAST * g.ld + i];
    c.cv_prev = g.st[SBR_CNT_CV_PREV * g.ld + i];
    c.ie_do = g.st[SBR_CNT_IE_DO * g.ld + i];
    c.ie_ec = g.st[SBR_CNT_IE_EC * g.ld + i];
    c.kla_last = g.st[SBR_CNT_KLA_LAST * g.ld + i];
    c.ec_last = g.st[SBR_CNT_EC_LAST * g.ld + i];
    h = g.st[SBR_CNT_H * g.ld + i];
}
__device__ __forceinline__ void cnt_store_ctrl(const CntArgs& g, int64_t i, const sbr::CntCtrl& c, double h) {
    g.st[SBR_CNT_T * g.ld + i] = c.t;
    g.st[SBR_CNT_U_DO * g.ld + i] = c.u_do;
    g.st[SBR_CNT_U_EC * g.ld + i] = c.u_ec;
    g.st[SBR_CNT_SO_PREV * g.ld + i] = c.so_prev;
    g.st[SBR_CNT_CV_LAST * g.ld + i] = c.cv_last;
    g.st[SBR_CNT_CV_PREV * g.ld + i] = c.cv_prev;
    g.st[SBR_CNT_IE_DO * g.ld + i] = c.ie_do;
    g.st[SBR_CNT_IE_EC * g.ld + i] = c.ie_ec;
    g.st[SBR_CNT_KLA_LAST * g.ld + i] = c.kla_last;
    g.st[SBR_CNT_EC_LAST * g.ld + i] = c.ec_last;
    g.st[SBR_CNT_H * g.ld + i] = h;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) sbr_cnt_reset_kernel(CntArgs g, sbr::CntCfg q, SbrParams p, sbr::Coef c,
                                                               SbrOsSchedule s, SbrTol tol) {
    __shared__ double s_load[SBR_NX * kBlock];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    if (g.mask && g.mask[i] == 0) return;
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.x0 ? g.x0[k * g.ld + i] : c_x0_init[k];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) s_load[k * kBlock + threadIdx.x] = g.influent[k * g.ld + i];
    const sbr::Loading load{&s_load[threadIdx.x], kBlock};
    sbr::Dp45State dp;
    dp.h = s.t_fill / (double)(s.fill_pts > 1 ? s.fill_pts - 1 : 1); dp.n_rhs = 0; dp.n_rej = 0;
    sbr::CntCtrl ctl;
    const int status = sbr::cnt_reset_env<MODE>(x, load, q, p, c, s, tol, dp, ctl, sbr::Column{g.obs + i, g.ld});
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.st[k * g.ld + i] = x[k];
    cnt_store_ctrl(g, i, ctl, s.t_delta / 9.0);
    g.st[SBR_CNT_RETURN * g.ld + i] = 0.0;
    g.st[SBR_CNT_STEPS * g.ld + i] = 0.0;
    g.st[SBR_CNT_QW * g.ld + i] = NAN;
    g.done[i] = 0;
    if (g.status) g.status[i] = status;
    if (g.counters) { g.counters[i] = dp.n_rhs; g.counters[g.ld + i] = dp.n_rej; }
}

// observation rows the policy head of the fused rollout reads: all of them, for SBROS-v2 obs_DO and obs_EC (not `state`)
__host__ __device__ constexpr int cnt_policy_inputs(int kind) {
    return kind == SBR_CNT_OS2 ? 2 * SBR_OS_NOBS : sbr::cnt_obs_rows(kind);
}

// FUSED = false: sbr_cnt_step (one step, no policy code in the kernel); FUSED = true: sbr_cnt_rollout_k.
template <int MODE, bool FUSED>
__global__ void __launch_bounds__(kBlock, (FUSED && MODE == SBR_MODE_DP45) ? SBR_V4_FUSED_MINBLOCKS_DP45 : cnt_step_minblocks(MODE))
sbr_cnt_step_kernel(CntArgs g, sbr::CntCfg q, SbrParams p,
                                                                                     sbr::Coef c, SbrOsSchedule s, SbrTol tol) {
    __shared__ float s_pw1[FUSED ? kPolMaxHidden * 2 * SBR_OS_NOBS : 1];
    __shared__ float s_pw2[FUSED ? 2 * kPolMaxHidden : 1];
    const int n_in = cnt_policy_inputs(q.kind), n_out = q.kind == SBR_CNT_OS2 ? 2 : 1;
    if (FUSED) {
        for (int k = threadIdx.x; k < g.pol_hidden * n_in; k += kBlock) s_pw1[k] = g.pol_w1[k];
        for (int k = threadIdx.x; k < n_out * g.pol_hidden; k += kBlock) s_pw2[k] = g.pol_w2[k];
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= g.n) return;
    double x[SBR_NX];
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k] = g.st[k * g.ld + i];
    const uint8_t was_done = g.done[i];
    sbr::CntCtrl ctl;
    sbr::Dp45State dp;
    cnt_load_ctrl(g, i, ctl, dp.h);
    const double* act_in = FUSED ? g.action_io : g.action;
    double a0 = act_in[i];
    double a1 = q.kind == SBR_CNT_OS2 ? act_in[g.ld + i] : 0.0;
    double ret = g.st[SBR_CNT_RETURN * g.ld + i], steps = g.st[SBR_CNT_STEPS * g.ld + i];
    const int K = FUSED ? g.K : 1;
    if (was_done) {
        // stepping a finished episode is a no-op: the observation buffer keeps the terminal observation, reward 0
        for (int k = 0; k < K; ++k) g.reward[(int64_t)k * g.ld + i] = 0.0;
        if (g.status) g.status[i] = SBR_ST_DONE;
        if (g.counters) { g.counters[i] = 0; g.counters[g.ld + i] = 0; }
        return;
    }
    if (!(dp.h > 0.0)) dp.h = s.t_delta / 9.0;
    dp.n_rhs = 0; dp.n_rej = 0;
    const sbr::Column ob{g.obs ? g.obs + i : nullptr, g.ld};
    int status = 0;
    bool is_done = false;
    double qw = NAN;
    for (int k = 0; k < K; ++k) {
        if (is_done) { g.reward[(int64_t)k * g.ld + i] = 0.0; continue; }
        if (FUSED && g.act_log) { g.act_log[(int64_t)(2 * k) * g.ld + i] = a0; g.act_log[(int64_t)(2 * k + 1) * g.ld + i] = a1; }
        sbr::CntOut o;
        double obl[FUSED ? SBR_CNT_NOBS_MAX : 1];
        sbr::cnt_step_env<MODE>(x, ctl, a0, a1, q, p, c, s, tol, dp, FUSED ? sbr::Column{obl, 1} : ob, o);
        g.reward[(int64_t)k * g.ld + i] = o.reward;
        ret += o.reward;
        steps += 1.0;
        status |= o.status;
        if (o.done) { is_done = true; qw = o.Qw; }
        if (FUSED) {
            const int rows = sbr::cnt_obs_rows(q.kind);
            // the observation of the last step that ran goes to `obs`; every step's feeds the policy head
            if (k == K - 1 || o.done) {
#pragma unroll
                for (int r = 0; r < SBR_CNT_NOBS_MAX; ++r)
                    if (r < rows) ob.set(r, obl[FUSED ? r : 0]);
            }
            if (g.obs_log) {
#pragma unroll
                for (int r = 0; r < 2 * SBR_OS_NOBS; ++r)
                    if (r < n_in) g.obs_log[((int64_t)k * n_in + r) * g.ld + i] = obl[FUSED ? r : 0];
            }
            float xin[kPolMaxIn], act[kPolMaxOut];
#pragma unroll
            for (int r = 0; r < kPolMaxIn; ++r) xin[r] = 0.0f;
#pragma unroll
            for (int r = 0; r < 2 * SBR_OS_NOBS; ++r)
                if (r < n_in) xin[r] = (float)obl[FUSED ? r : 0];
            policy_eval(s_pw1, s_pw2, g.pol_lo, g.pol_span, n_in, g.pol_hidden, n_out, xin, act);
            a0 = (double)act[0];
            if (q.kind == SBR_CNT_OS2) a1 = (double)act[1];
        }
    }
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) g.st[k * g.ld + i] = x[k];
    cnt_store_ctrl(g, i, ctl, dp.h);
    g.st[SBR_CNT_RETURN * g.ld + i] = ret;
    g.st[SBR_CNT_STEPS * g.ld + i] = steps;
    if (is_done) { g.st[SBR_CNT_QW * g.ld + i] = qw; g.done[i] = 1; }
    if (FUSED) { g.action_io[i] = a0; if (q.kind == SBR_CNT_OS2) g.action_io[g.ld + i] = a1; }
    if (g.status) g.status[i] = status;
    if (g.counters) { g.counters[i] = dp.n_rhs; g.counters[g.ld + i] = dp.n_rej; }
}

// Influent mixing (buffer_tank3.py:50-107): one env per thread, tables staged in shared memory, 13 running sums in
// registers, rnd read coalesced ([48][N]).  No FMA contraction and sequential sums => bit-identical to numpy.
__global__ void __launch_bounds__(128) sbr_influent_mix_kernel(int64_t n, int64_t ld, const double* __restrict__ rnd,
                                                               const double* __restrict__ mean,
                                                               const double* __restrict__ stdv, double* influent) {
    __shared__ double s_mean[SBR_NX * SBR_INFLUENT_POINTS], s_std[SBR_NX * SBR_INFLUENT_POINTS];
    for (int k = threadIdx.x; k < SBR_NX * SBR_INFLUENT_POINTS; k += blockDim.x) { s_mean[k] = mean[k]; s_std[k] = stdv[k]; }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc[SBR_NX];
#pragma unroll
    for (int j = 0; j < SBR_NX; ++j) acc[j] = 0.0;
    for (int t = 0; t < SBR_INFLUENT_POINTS; ++t) {
        const double z = rnd[t * ld + i];
        const double q = __dadd_rn(s_mean[t], __dmul_rn(s_std[t], z));
        acc[0] = __dadd_rn(acc[0], q);
#pragma unroll
        for (int j = 1; j < SBR_NX; ++j) {
            const double cj = __dadd_rn(s_mean[j * SBR_INFLUENT_POINTS + t], __dmul_rn(s_std[j * SBR_INFLUENT_POINTS + t], z));
            acc[j] = __dadd_rn(acc[j], __dmul_rn(cj, q));
        }
    }
    influent[i] = 0.66;                                                    // buffer_tank3.py:92
#pragma unroll
    for (int j = 1; j < SBR_NX; ++j) influent[j * ld + i] = __ddiv_rn(acc[j], acc[0]);
}

// ---------------------------------------------------------------------------------------------------------
// Counter-based influent draws: Philox4x32-10 keyed by the run's seed, counter = (GLOBAL env index, episode number,
// block).  An env's draws depend on nothing but (seed, its global index, its episode number): not on the batch
// size, the rank that owns it, the world size or the order of resets -- which is what makes results invariant
// to the sharding.  Normals by Box-Muller from two 53-bit uniforms; 24 Philox blocks give the 48
// draws of one buffer_tank call, block 24 the scenario of SbrEnv4's np.random.choice(8, 1).
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct SampleArgs {
    int64_t n, ld, env_offset, epoch0;
    uint64_t seed;
    int64_t* epoch;            // [n] in/out (may be NULL: every env uses epoch0)
    const uint8_t* mask;       // [n] (may be NULL)
    const double* mean;        // [8][14][48]
    const double* stdv;        // [8][14][48]
    double* influent;          // [14][ld]
    int32_t* scenario_out;     // [n] (may be NULL)
    int scenario;              // 0..7, or -1: drawn per env
};

__global__ void __launch_bounds__(128) sbr_influent_sample_kernel(SampleArgs g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    if (g.mask && g.mask[i] == 0) return;
    const uint64_t env = (uint64_t)(g.env_offset + i);
    const int64_t ep = g.epoch ? g.epoch[i] : g.epoch0;
    const uint32_t k0 = (uint32_t)g.seed, k1 = (uint32_t)(g.seed >> 32);
    int scn = g.scenario;
    if (scn < 0) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)ep, 24u};
        philox4x32_10(c, k0, k1);
        scn = (int)(c[0] >> 29);                                            // uniform on 0..7
    }
    const double* __restrict__ mean = g.mean + (size_t)scn * SBR_NX * SBR_INFLUENT_POINTS;
    const double* __restrict__ stdv = g.stdv + (size_t)scn * SBR_NX * SBR_INFLUENT_POINTS;
    double acc[SBR_NX];
#pragma unroll
    for (int j = 0; j < SBR_NX; ++j) acc[j] = 0.0;
    for (int b = 0; b < SBR_INFLUENT_POINTS / 2; ++b) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)ep, (uint32_t)b};
        philox4x32_10(c, k0, k1);
        // two uniforms in (0, 1): 53 random bits + half an ulp
        const double u1 = ((double)((((uint64_t)c[1] << 32) | c[0]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)((((uint64_t)c[3] << 32) | c[2]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double r = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        const double zz[2] = {r * cs, r * sn};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = 2 * b + h;
            const double z = zz[h];
            // same arithmetic as sbr_influent_mix_kernel (bit-identical to numpy for the same z)
            const double q = __dadd_rn(__ldg(mean + t), __dmul_rn(__ldg(stdv + t), z));
            acc[0] = __dadd_rn(acc[0], q);
#pragma unroll
            for (int j = 1; j < SBR_NX; ++j) {
                const double cj = __dadd_rn(__ldg(mean + j * SBR_INFLUENT_POINTS + t),
                                            __dmul_rn(__ldg(stdv + j * SBR_INFLUENT_POINTS + t), z));
                acc[j] = __dadd_rn(acc[j], __dmul_rn(cj, q));
            }
        }
    }
    g.influent[i] = 0.66;                                                   // buffer_tank3.py:92
#pragma unroll
    for (int j = 1; j < SBR_NX; ++j) g.influent[j * g.ld + i] = __ddiv_rn(acc[j], acc[0]);
    if (g.scenario_out) g.scenario_out[i] = scn;
    if (g.epoch) g.epoch[i] = ep + 1;
}

// Standard normals of one env's episode, as the sampler draws them: z [48][ld] (tests, and the parity story of the
// generator: mixing these with sbr_influent_mix reproduces sbr_influent_sample bit for bit).
__global__ void __launch_bounds__(128) sbr_philox_normals_kernel(int64_t n, int64_t ld, uint64_t seed, int64_t env_offset,
                                                                 int64_t epoch0, double* z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t env = (uint64_t)(env_offset + i);
    for (int b = 0; b < SBR_INFLUENT_POINTS / 2; ++b) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)epoch0, (uint32_t)b};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = ((double)((((uint64_t)c[1] << 32) | c[0]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)((((uint64_t)c[3] << 32) | c[2]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        const double r = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        z[(2 * b) * ld + i] = r * cs;
        z[(2 * b + 1) * ld + i] = r * sn;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Row permutation of SoA buffers: dst[r][i] = src[r][perm[i]] (gather) or dst[r][perm[i]] = src[r][i] (scatter), for
// up to SBR_PERMUTE_MAX buffers per launch.  Used to hand the adaptive cycle kernel its envs in divergence-aware
// order with unit-stride loads and stores.  blockIdx.y = (buffer, row): the blocks of one row run together, so the
// row's 8 n bytes and perm stay in L2 while the random side touches each 32-B sector four times -- DRAM traffic
// stays at the algorithmic bytes.
// ---------------------------------------------------------------------------------------------------------
struct PermuteArgs {
    int64_t n;
    const int64_t* perm;
    const void* src[SBR_PERMUTE_MAX];
    void* dst[SBR_PERMUTE_MAX];
    int64_t ld_src[SBR_PERMUTE_MAX], ld_dst[SBR_PERMUTE_MAX];
    int32_t row0[SBR_PERMUTE_MAX + 1];      // first blockIdx.y of each buffer
    int32_t elem[SBR_PERMUTE_MAX];          // element size: 4 or 8 bytes
    int32_t nbuf, scatter;
};

__global__ void __launch_bounds__(256) sbr_permute_rows_kernel(PermuteArgs g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    int b = 0;
#pragma unroll
    for (int k = 1; k < SBR_PERMUTE_MAX; ++k)
        if (k < g.nbuf && (int)blockIdx.y >= g.row0[k]) b = k;
    const int64_t r = (int64_t)blockIdx.y - g.row0[b];
    const int64_t j = g.perm[i];
    const int64_t is = g.scatter ? i : j, id = g.scatter ? j : i;
    if (g.elem[b] == 8)
        ((double*)g.dst[b])[r * g.ld_dst[b] + id] = ((const double*)g.src[b])[r * g.ld_src[b] + is];
    else
        ((int32_t*)g.dst[b])[r * g.ld_dst[b] + id] = ((const int32_t*)g.src[b])[r * g.ld_src[b] + is];
}

// Stage-level seam of the settle + draw phases (unit tests): x in/out, sX [10][ld], out [9][ld] = Xf, Qw, EQI, eff[6].
__global__ void __launch_bounds__(128) sbr_settle_draw_kernel(int64_t n, int64_t ld, double* x, double T, SbrParams p,
                                                              double* sX_out, double* out, int32_t* status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xs[SBR_NX], sX[10], Xf;
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) xs[k] = x[k * ld + i];
    sbr::settle_closed_form(xs, T, p.settler_area, p.settler_vmax, sX, Xf);
    sbr::DrawOut d;
    sbr::draw_and_waste(xs, sX, Xf, p.Qeff, p.biomass_setpoint, d);
#pragma unroll
    for (int k = 0; k < SBR_NX; ++k) x[k * ld + i] = xs[k];
#pragma unroll
    for (int k = 0; k < 10; ++k) sX_out[k * ld + i] = sX[k];
    out[i] = Xf; out[ld + i] = d.Qw; out[2 * ld + i] = d.EQI;
#pragma unroll
    for (int k = 0; k < 6; ++k) out[(3 + k) * ld + i] = d.eff[k];
    if (status) status[i] = d.status;
}

// FP64 pipe probe: 8 independent DFMA chains per thread, `iters` rounds of 8 DFMAs each.
__global__ void sbr_fp64_probe_kernel(int iters, double* sink) {
    const double a = 1.0000001, b = 1e-9 * (double)(threadIdx.x + 1);
    double v0 = 1.0, v1 = 1.1, v2 = 1.2, v3 = 1.3, v4 = 1.4, v5 = 1.5, v6 = 1.6, v7 = 1.7;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

// Reward statistics: grid-stride partial sums, warp shuffle + one atomic per warp.
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double(assumed) <= v) break;
        old = atomicCAS(a, assumed, __double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double(assumed) >= v) break;
        old = atomicCAS(a, assumed, __double_as_longlong(v));
    } while (assumed != old);
}

__global__ void sbr_reward_stats_init_kernel(double* stats) {
    stats[0] = 0.0; stats[1] = 0.0; stats[2] = INFINITY; stats[3] = -INFINITY; stats[4] = 0.0;
}

__global__ void __launch_bounds__(256) sbr_reward_stats_kernel(int64_t n, const double* __restrict__ reward,
                                                               const int32_t* __restrict__ status, double* stats) {
    double s = 0.0, ss = 0.0, mn = INFINITY, mx = -INFINITY, cnt = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (status && status[i] != 0) continue;
        const double r = reward[i];
        s += r; ss = fma(r, r, ss); mn = fmin(mn, r); mx = fmax(mx, r); cnt += 1.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        ss += __shfl_down_sync(0xffffffffu, ss, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
        atomicAdd(&stats[0], s);
        atomicAdd(&stats[1], ss);
        atomic_min_f64(&stats[2], mn);
        atomic_max_f64(&stats[3], mx);
        atomicAdd(&stats[4], cnt);
    }
}

int check_common(int64_t n, int64_t ld, const SbrParams* p) {
    if (n <= 0) return fail(SBR_ERR_ARG, "n must be positive%s");
    if (ld < n) return fail(SBR_ERR_ARG, "ld must be >= n%s");
    if (!p) return fail(SBR_ERR_ARG, "params pointer is NULL%s");
    if ((n + kBlock - 1) / kBlock > 2147483647LL) return fail(SBR_ERR_ARG, "n too large for one launch%s");
    return SBR_OK;
}

SbrTol tol_or_default(const SbrTol* tol) {
    SbrTol t;
    t.rtol = 1e-8; t.atol = 1e-10; t.max_steps = 200; t.flags = 0;
    if (tol) t = *tol;
    return t;
}

int check_os_schedule(const SbrOsSchedule* s) {
    if (!s) return fail(SBR_ERR_ARG, "schedule pointer is NULL%s");
    if (!(s->dt > 0) || !(s->t_delta > 0) || !(s->t_fill > 0) || !(s->t_cycle > 0) || s->fill_pts < 2 ||
        !(s->tm3_0 < s->tm3_1 && s->tm3_1 < s->tm4_1 && s->tm4_1 <= s->tm5_1) || !(s->settle_len > 0) ||
        !(s->draw_len >= 0) || s->rk4_sub_interval < 0 || s->rk4_sub_fill < 0 || s->rk4_sub_idle < 0)
        return fail(SBR_ERR_ARG, "SbrOsSchedule: inconsistent time constants%s");
    return SBR_OK;
}

}  // namespace

extern "C" {

int sbr_abi_version(void) { return SBR_ABI_VERSION; }

const char* sbr_last_error(void) { return g_err; }

int sbr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void sbr_params_default(SbrParams* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->muh = 4.0; p->Ks = 10.0; p->Koh = 0.2; p->Kno = 0.5; p->bh = 0.3; p->etag = 0.8; p->etah = 0.8;
    p->kh = 3.0; p->Kx = 0.1; p->mua = 0.5; p->Knh = 1.0; p->ba = 0.05; p->Koa = 0.4; p->ka = 0.05;
    p->Ya = 0.24; p->Yh = 0.67; p->fp = 0.08; p->ixb = 0.08; p->ixp = 0.06;
    {   // module_temperature.py:3-20 at 15 C
        const double tk = (15 + 273.15) / 100;
        const double f = 56.12 * exp(-66.7354 + 87.4755 / tk + 24.4526 * log(tk));
        p->so_sat = 0.9997743214 * (8 / 10.5) * 6791.5 * f;
    }
    p->pid_Kc = 5.0; p->pid_tauI = 0.00035; p->pid_tauD = 0.005; p->pid_dt = 0.02 / 24;
    p->kla_min = 0.0; p->kla_max = 240.0;
    p->WV = 1.32; p->Qin = 1.32 - 0.6161484733495801; p->Qeff = 0.66; p->biomass_setpoint = 2700.0;
    p->settler_area = (1.25 / 2) * (1.25 / 2); p->settler_vmax = 474.0;
    p->kla0 = 0.0; p->action_scale = 8.0;
    p->os_Kc_DO = 100.0; p->os_tauI_DO = 20.0; p->os_tauD_DO = 0.0;
    p->os_Kc_EC = 100.0; p->os_tauI_EC = 20.0; p->os_tauD_EC = 0.0;
    p->os_pid_dt = 0.002 / 24; p->ec_min = 0.0; p->ec_max = 0.0005; p->ec_conc = 1200000.0 * 4;
    p->do_sp_max = 8.0; p->no_sp_max = 15.0;
    p->IV = 0.6161484733495801;
}

int sbr_cycle_v2(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                 const SbrParams* p, const SbrSchedule* s, double* x_last, double* obs, double* reward,
                 double* aux, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol,
                 const int64_t* perm, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if (!x0 || !influent || !action || !s || !x_last || !obs || !reward)
        return fail(SBR_ERR_ARG, "sbr_cycle_v2: NULL buffer%s");
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "sbr_cycle_v2: bad mode%s");
    for (int k = 0; k < SBR_NPHASE; ++k) {
        if (k == 5 || k == 6) continue;
        if (s->n_int[k] < 1 || s->n_sub[k] < 1 || !(s->interval[k] > 0))
            return fail(SBR_ERR_ARG, "sbr_cycle_v2: schedule needs n_int, n_sub >= 1 and interval > 0%s");
    }
    CycleArgs g{n, ld, x0, influent, action, x_last, obs, reward, aux, status, counters, perm};
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SBR_MODE_RK4)
        sbr_cycle_v2_kernel<SBR_MODE_RK4><<<grid, kBlock, 0, st>>>(g, *p, c, *s, t);
    else
        sbr_cycle_v2_kernel<SBR_MODE_DP45><<<grid, kBlock, 0, st>>>(g, *p, c, *s, t);
    return check_launch("sbr_cycle_v2");
}

int sbr_cycle_v2_traj_records(const SbrSchedule* s) {
    if (!s) return 0;
    int r = 1;
    for (int k = 0; k < SBR_NPHASE; ++k)
        if (k != 5 && k != 6) r += s->n_int[k];
    return r;
}

int sbr_cycle_v2_traj(int64_t n, int64_t ld, const double* x0, const double* influent, const double* action,
                      const SbrParams* p, const SbrSchedule* s, const double* t_start, double* x_last, double* obs,
                      double* reward, double* aux, int32_t* status, uint32_t* counters, double* traj, int mode,
                      const SbrTol* tol, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if (!x0 || !influent || !action || !s || !t_start || !x_last || !obs || !reward || !traj)
        return fail(SBR_ERR_ARG, "sbr_cycle_v2_traj: NULL buffer%s");
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "sbr_cycle_v2_traj: bad mode%s");
    for (int k = 0; k < SBR_NPHASE; ++k) {
        if (k == 5 || k == 6) continue;
        if (s->n_int[k] < 1 || s->n_sub[k] < 1 || !(s->interval[k] > 0))
            return fail(SBR_ERR_ARG, "sbr_cycle_v2_traj: schedule needs n_int, n_sub >= 1 and interval > 0%s");
    }
    CycleTrajArgs h;
    h.c = CycleArgs{n, ld, x0, influent, action, x_last, obs, reward, aux, status, counters, nullptr};
    h.traj = traj;
    for (int k = 0; k < SBR_NPHASE; ++k) h.t_start[k] = t_start[k];
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SBR_MODE_RK4)
        sbr_cycle_v2_traj_kernel<SBR_MODE_RK4><<<grid, kBlock, 0, st>>>(h, *p, c, *s, t);
    else
        sbr_cycle_v2_traj_kernel<SBR_MODE_DP45><<<grid, kBlock, 0, st>>>(h, *p, c, *s, t);
    return check_launch("sbr_cycle_v2_traj");
}

static int check_ilc_layout(const SbrIlcLayout* lay, const SbrSchedule* s, const char* who) {
    if (!lay || lay->n_samples < 1) return fail(SBR_ERR_ARG, "%s: NULL or empty layout", who);
    for (int j = 0; j < 6; ++j) {
        const int hi = j < 5 ? lay->off[j + 1] : lay->n_samples;
        if (lay->off[j] < 0 || hi <= lay->off[j]) return fail(SBR_ERR_ARG, "%s: layout offsets must increase", who);
        if (s) {
            const int ph = j < 5 ? j : 7;
            // the indices 9 i + ii + 1 the reference reads (sub_phases_batchPID_fbPID.py:185,230) must stay inside the phase
            const int need = s->n_int[ph] * s->n_sub[ph] + 1, last = 9 * (s->n_int[ph] - 1) + s->n_sub[ph];
            if (hi - lay->off[j] < need || last >= hi - lay->off[j])
                return fail(SBR_ERR_ARG, "%s: layout holds fewer samples than the schedule writes", who);
        }
    }
    return SBR_OK;
}

int sbr_cycle_ilc(int64_t n, int64_t ld, const double* x0, const double* influent, const double* sp,
                  const SbrParams* p, const SbrSchedule* s, const SbrIlcLayout* lay, double t_fill,
                  const double* kla_base, const double* u, double* so_mem, double* kla_mem, double* x_last,
                  double* out, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream) {
    int rc = check_common(n, ld, p);
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "sbr_cycle_ilc: bad mode%s");
    if (rc) return rc;
    if (!x0 || !influent || !sp || !s || !x_last) return fail(SBR_ERR_ARG, "sbr_cycle_ilc: NULL buffer%s");
    if ((kla_base == nullptr) != (u == nullptr))
        return fail(SBR_ERR_ARG, "sbr_cycle_ilc: kla_base and u go together (both NULL = cycle 0)%s");
    if (!(t_fill > 0.0)) return fail(SBR_ERR_ARG, "sbr_cycle_ilc: t_fill must be positive%s");
    for (int k = 0; k < SBR_NPHASE; ++k) {
        if (k == 5 || k == 6) continue;
        if (s->n_int[k] < 1 || s->n_sub[k] < 1 || !(s->interval[k] > 0))
            return fail(SBR_ERR_ARG, "sbr_cycle_ilc: schedule needs n_int, n_sub >= 1 and interval > 0%s");
    }
    rc = check_ilc_layout(lay, s, "sbr_cycle_ilc");
    if (rc) return rc;
    IlcCycleArgs g{n, ld, x0, influent, sp, kla_base, u, so_mem, kla_mem, x_last, out, status, counters, t_fill};
    const sbr::Coef c = sbr::make_coef(*p);
    const SbrTol t = tol_or_default(tol);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    if (mode == SBR_MODE_RK4)
        sbr_cycle_ilc_kernel<SBR_MODE_RK4><<<grid, kBlock, 0, (cudaStream_t)stream>>>(g, *p, c, *s, *lay, t);
    else
        sbr_cycle_ilc_kernel<SBR_MODE_DP45><<<grid, kBlock, 0, (cudaStream_t)stream>>>(g, *p, c, *s, *lay, t);
    return check_launch("sbr_cycle_ilc");
}

int sbr_ilc_update(int64_t n, int64_t ld, const SbrIlcLayout* lay, const double* w, const double* D, const double* sp6,
                   const double* so_mem, double* e_sum, double* e_last, double* u, double dt, double Kc, double tauI,
                   double tauD, void* stream) {
    if (n <= 0 || ld < n) return fail(SBR_ERR_ARG, "sbr_ilc_update: need 0 < n <= ld%s");
    if (!w || !D || !sp6 || !so_mem || !e_sum || !e_last || !u) return fail(SBR_ERR_ARG, "sbr_ilc_update: NULL buffer%s");
    if (!(tauI != 0.0)) return fail(SBR_ERR_ARG, "sbr_ilc_update: tauI must be non-zero%s");
    int rc = check_ilc_layout(lay, nullptr, "sbr_ilc_update");
    if (rc) return rc;
    IlcUpdateArgs g{n, ld, w, D, sp6, so_mem, e_sum, e_last, u, dt, Kc, Kc / tauI, Kc * tauD};
    const dim3 grid((unsigned)((n + 127) / 128), 6);
    sbr_ilc_update_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(g, *lay);
    return check_launch("sbr_ilc_update");
}

int sbr_integrate_interval(int64_t n, int64_t ld, double* x, const double* kla, const double* ec,
                           const double* loading, const SbrParams* p, int tail, double T, int n_sub,
                           int mode, const SbrTol* tol, uint32_t* counters, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if (!x || !kla) return fail(SBR_ERR_ARG, "sbr_integrate_interval: NULL buffer%s");
    if (tail == sbr::TAIL_FILL && !loading) return fail(SBR_ERR_ARG, "sbr_integrate_interval: fill tail needs loading%s");
    if (tail == sbr::TAIL_EC && !ec) return fail(SBR_ERR_ARG, "sbr_integrate_interval: EC tail needs ec%s");
    if (!(T > 0) || n_sub < 1) return fail(SBR_ERR_ARG, "sbr_integrate_interval: T > 0 and n_sub >= 1 required%s");
    IntervalArgs g{n, ld, x, kla, ec, loading, counters, T, n_sub};
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t st = (cudaStream_t)stream;
#define SBR_LAUNCH(TAIL, MODE) sbr_interval_kernel<TAIL, MODE><<<grid, kBlock, 0, st>>>(g, *p, c, t)
    if (mode == SBR_MODE_RK4) {
        if (tail == sbr::TAIL_REACT) SBR_LAUNCH(sbr::TAIL_REACT, SBR_MODE_RK4);
        else if (tail == sbr::TAIL_FILL) SBR_LAUNCH(sbr::TAIL_FILL, SBR_MODE_RK4);
        else if (tail == sbr::TAIL_EC) SBR_LAUNCH(sbr::TAIL_EC, SBR_MODE_RK4);
        else return fail(SBR_ERR_ARG, "sbr_integrate_interval: bad tail%s");
    } else if (mode == SBR_MODE_DP45) {
        if (tail == sbr::TAIL_REACT) SBR_LAUNCH(sbr::TAIL_REACT, SBR_MODE_DP45);
        else if (tail == sbr::TAIL_FILL) SBR_LAUNCH(sbr::TAIL_FILL, SBR_MODE_DP45);
        else if (tail == sbr::TAIL_EC) SBR_LAUNCH(sbr::TAIL_EC, SBR_MODE_DP45);
        else return fail(SBR_ERR_ARG, "sbr_integrate_interval: bad tail%s");
    } else {
        return fail(SBR_ERR_ARG, "sbr_integrate_interval: bad mode%s");
    }
#undef SBR_LAUNCH
    return check_launch("sbr_integrate_interval");
}

int sbr_rhs(int64_t n, int64_t ld, const double* x, const double* kla, const double* ec, const double* loading,
            const SbrParams* p, int tail, double* dx, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if (!x || !kla || !dx) return fail(SBR_ERR_ARG, "sbr_rhs: NULL buffer%s");
    if (tail == sbr::TAIL_FILL && !loading) return fail(SBR_ERR_ARG, "sbr_rhs: fill tail needs loading%s");
    if (tail == sbr::TAIL_EC && !ec) return fail(SBR_ERR_ARG, "sbr_rhs: EC tail needs ec%s");
    RhsArgs g{n, ld, x, kla, ec, loading, dx};
    const sbr::Coef c = sbr::make_coef(*p);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t st = (cudaStream_t)stream;
    if (tail == sbr::TAIL_REACT) sbr_rhs_kernel<sbr::TAIL_REACT><<<grid, kBlock, 0, st>>>(g, *p, c);
    else if (tail == sbr::TAIL_FILL) sbr_rhs_kernel<sbr::TAIL_FILL><<<grid, kBlock, 0, st>>>(g, *p, c);
    else if (tail == sbr::TAIL_EC) sbr_rhs_kernel<sbr::TAIL_EC><<<grid, kBlock, 0, st>>>(g, *p, c);
    else return fail(SBR_ERR_ARG, "sbr_rhs: bad tail%s");
    return check_launch("sbr_rhs");
}

int sbr_os_reset(int64_t n, int64_t ld, const double* x0, const double* influent, const uint8_t* mask,
                 const SbrParams* p, const SbrOsSchedule* s, double* st, double* obs_do, double* obs_ec,
                 uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if ((rc = check_os_schedule(s))) return rc;
    if (!influent || !st || !obs_do || !obs_ec || !done) return fail(SBR_ERR_ARG, "sbr_os_reset: NULL buffer%s");
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "sbr_os_reset: bad mode%s");
    OsArgs g{n, ld, st, x0, influent, mask, obs_do, obs_ec, done, status, counters};
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t cs = (cudaStream_t)stream;
    if (mode == SBR_MODE_RK4) sbr_os_reset_kernel<SBR_MODE_RK4><<<grid, kBlock, 0, cs>>>(g, *p, c, *s, t);
    else sbr_os_reset_kernel<SBR_MODE_DP45><<<grid, kBlock, 0, cs>>>(g, *p, c, *s, t);
    return check_launch("sbr_os_reset");
}

int sbr_os_step_k(int64_t n, int64_t ld, int K, double* st, const double* action, const SbrParams* p,
                  const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                  uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream) {
    return sbr_os_step_traj(n, ld, K, st, action, p, s, obs_do, obs_ec, state, reward, done, status, counters, mode, tol,
                            nullptr, 0, stream);
}

namespace {
struct OsPolicy {
    const float* w1; const float* w2; const float* lo; const float* span; int hidden;
    double* act_log; double* obs_log;
};

int os_step_launch(const char* what, int64_t n, int64_t ld, int K, double* st, double* action, const SbrParams* p,
                   const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                   uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol,
                   double* traj, int traj_cap, const OsPolicy* pol, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if ((rc = check_os_schedule(s))) return rc;
    if (K < 1 || K > 4096) return fail(SBR_ERR_ARG, "%s: K must be in 1..4096", what);
    if (!st || !action || !reward || !done) return fail(SBR_ERR_ARG, "%s: NULL buffer", what);
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "%s: bad mode", what);
    OsStepArgs g;
    g.n = n; g.ld = ld; g.num_tiles = (n + kOsTile - 1) / kOsTile;
    g.st = st; g.action = action; g.obs_do = obs_do; g.obs_ec = obs_ec; g.state = state; g.reward = reward;
    g.done = done; g.status = status; g.counters = counters; g.K = K;
    if (traj && traj_cap < 1) return fail(SBR_ERR_ARG, "%s: traj_cap must be positive", what);
    g.traj = traj; g.traj_cap = traj ? traj_cap : 0;
    g.pol_w1 = nullptr; g.pol_w2 = nullptr; g.pol_lo = nullptr; g.pol_span = nullptr; g.pol_hidden = 0;
    g.act_log = nullptr; g.obs_log = nullptr;
    if (pol) {
        g.pol_w1 = pol->w1; g.pol_w2 = pol->w2; g.pol_lo = pol->lo; g.pol_span = pol->span; g.pol_hidden = pol->hidden;
        g.act_log = pol->act_log; g.obs_log = pol->obs_log;
    }
    // TMA needs 16-byte aligned bases and row pitches (and 32-bit coordinates); anything else takes the plain-load fill
    CUtensorMap tm_st, tm_act, tm_done;
    memset(&tm_st, 0, sizeof(tm_st)); memset(&tm_act, 0, sizeof(tm_act)); memset(&tm_done, 0, sizeof(tm_done));
    g.tma = ((uintptr_t)st % 16 == 0 && (uintptr_t)action % 16 == 0 && (uintptr_t)done % 16 == 0 && ld % 2 == 0 &&
             n < (int64_t)1 << 31) ? 1 : 0;
    if (g.tma)
        g.tma = make_map_2d(&tm_st, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, st, n, SBR_OS_QW, ld, kOsTile, SBR_OS_QW) &&
                make_map_2d(&tm_act, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, action, n, 2, ld, kOsTile, 2) &&
                make_map_2d(&tm_done, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, done, n, 1, n, kOsTile, 1) ? 1 : 0;
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    cudaStream_t cs = (cudaStream_t)stream;
    if (mode == SBR_MODE_RK4)
        sbr_os_step_kernel<SBR_MODE_RK4><<<os_step_grid<SBR_MODE_RK4>(g.num_tiles), kOsBlock, 0, cs>>>(
            g, *p, c, *s, t, tm_st, tm_act, tm_done);
    else
        sbr_os_step_kernel<SBR_MODE_DP45><<<os_step_grid<SBR_MODE_DP45>(g.num_tiles), kOsBlock, 0, cs>>>(
            g, *p, c, *s, t, tm_st, tm_act, tm_done);
    return check_launch(what);
}
}  // namespace

int sbr_os_step_traj(int64_t n, int64_t ld, int K, double* st, const double* action, const SbrParams* p,
                     const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                     uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol,
                     double* traj, int traj_cap, void* stream) {
    return os_step_launch("sbr_os_step", n, ld, K, st, const_cast<double*>(action), p, s, obs_do, obs_ec, state, reward,
                          done, status, counters, mode, tol, traj, traj_cap, nullptr, stream);
}

int sbr_os_rollout_k(int64_t n, int64_t ld, int K, double* st, double* action, const SbrPolicyMlp* policy,
                     const SbrParams* p, const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state,
                     double* reward, uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log,
                     int mode, const SbrTol* tol, void* stream) {
    if (!policy || !policy->w1 || !policy->w2 || !policy->lo || !policy->span)
        return fail(SBR_ERR_ARG, "sbr_os_rollout_k: NULL policy%s");
    if (policy->n_in != 2 * SBR_OS_NOBS || policy->n_out != 2 || policy->hidden < 1 || policy->hidden > kPolMaxHidden)
        return fail(SBR_ERR_ARG, "sbr_os_rollout_k: the policy must map 18 observations to 2 set-points (hidden <= 64)%s");
    const OsPolicy pol{policy->w1, policy->w2, policy->lo, policy->span, policy->hidden, act_log, obs_log};
    return os_step_launch("sbr_os_rollout_k", n, ld, K, st, action, p, s, obs_do, obs_ec, state, reward, done, status,
                          counters, mode, tol, nullptr, 0, &pol, stream);
}

int sbr_os_step(int64_t n, int64_t ld, double* st, const double* action, const SbrParams* p,
                const SbrOsSchedule* s, double* obs_do, double* obs_ec, double* state, double* reward,
                uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream) {
    return sbr_os_step_k(n, ld, 1, st, action, p, s, obs_do, obs_ec, state, reward, done, status, counters, mode, tol,
                         stream);
}

int sbr_v4_reset(int64_t n, int64_t ld, const double* x0, const double* influent, const uint8_t* mask,
                 const SbrParams* p, double* st, double* obs, uint8_t* done, const int32_t* order, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if (!influent || !st || !obs || !done) return fail(SBR_ERR_ARG, "sbr_v4_reset: NULL buffer%s");
    if (ld > 2147483647LL) return fail(SBR_ERR_ARG, "sbr_v4_reset: ld too large%s");
    V4Args g{n, ld, st, x0, influent, mask, nullptr, obs, nullptr, done, nullptr, nullptr, order};
    sbr_v4_reset_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g, *p);
    return check_launch("sbr_v4_reset");
}

namespace {
int v4_step_launch(const char* what, int64_t n, int64_t ld, int K, double* st, const double* influent, const double* action,
                   double* action_io, const SbrPolicyMlp* policy, const SbrParams* p, const SbrOsSchedule* s, double* obs,
                   double* reward, uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log,
                   int mode, const SbrTol* tol, const int32_t* order, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if ((rc = check_os_schedule(s))) return rc;
    if (!st || !influent || !(action || action_io) || !reward || !done) return fail(SBR_ERR_ARG, "%s: NULL buffer", what);
    if (n > 2147483647LL) return fail(SBR_ERR_ARG, "%s: n too large", what);
    if (K < 1 || K > 4096) return fail(SBR_ERR_ARG, "%s: K must be in 1..4096", what);
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "%s: bad mode", what);
    V4Args g{n, ld, st, nullptr, influent, nullptr, action, obs, reward, done, status, counters, order};
    g.K = K;
    g.pol_w1 = nullptr; g.pol_w2 = nullptr; g.pol_lo = nullptr; g.pol_span = nullptr; g.pol_hidden = 0;
    g.action_io = action_io; g.act_log = act_log; g.obs_log = obs_log;
    if (policy) {
        g.pol_w1 = policy->w1; g.pol_w2 = policy->w2; g.pol_lo = policy->lo; g.pol_span = policy->span;
        g.pol_hidden = policy->hidden;
    }
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t cs = (cudaStream_t)stream;
    if (policy) {
        if (mode == SBR_MODE_RK4) sbr_v4_step_kernel<SBR_MODE_RK4, true><<<grid, kBlock, 0, cs>>>(g, *p, c, *s, t);
        else sbr_v4_step_kernel<SBR_MODE_DP45, true><<<grid, kBlock, 0, cs>>>(g, *p, c, *s, t);
    } else {
        if (mode == SBR_MODE_RK4) sbr_v4_step_kernel<SBR_MODE_RK4, false><<<grid, kBlock, 0, cs>>>(g, *p, c, *s, t);
        else sbr_v4_step_kernel<SBR_MODE_DP45, false><<<grid, kBlock, 0, cs>>>(g, *p, c, *s, t);
    }
    return check_launch(what);
}
}  // namespace

int sbr_v4_step(int64_t n, int64_t ld, double* st, const double* influent, const double* action,
                const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward, uint8_t* done,
                int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, const int32_t* order, void* stream) {
    if (!obs) return fail(SBR_ERR_ARG, "sbr_v4_step: NULL buffer%s");
    return v4_step_launch("sbr_v4_step", n, ld, 1, st, influent, action, nullptr, nullptr, p, s, obs, reward, done, status,
                          counters, nullptr, nullptr, mode, tol, order, stream);
}

int sbr_v4_rollout_k(int64_t n, int64_t ld, int K, double* st, const double* influent, double* action,
                     const SbrPolicyMlp* policy, const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward,
                     uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log, int mode,
                     const SbrTol* tol, void* stream) {
    if (!policy || !policy->w1 || !policy->w2 || !policy->lo || !policy->span)
        return fail(SBR_ERR_ARG, "sbr_v4_rollout_k: NULL policy%s");
    if (policy->n_in != SBR_NX || policy->n_out != 1 || policy->hidden < 1 || policy->hidden > kPolMaxHidden)
        return fail(SBR_ERR_ARG, "sbr_v4_rollout_k: the policy must map 14 observations to 1 action (hidden <= 64)%s");
    return v4_step_launch("sbr_v4_rollout_k", n, ld, K, st, influent, nullptr, action, policy, p, s, obs, reward, done, status,
                          counters, act_log, obs_log, mode, tol, nullptr, stream);
}

int sbr_cnt_obs_rows(int kind) {
    return (kind >= 0 && kind < SBR_CNT_KINDS) ? sbr::cnt_obs_rows(kind) : -1;
}

static int check_cnt_config(const SbrCntConfig* q) {
    if (!q) return fail(SBR_ERR_ARG, "SbrCntConfig is NULL%s");
    if (q->kind < 0 || q->kind >= SBR_CNT_KINDS) return fail(SBR_ERR_ARG, "SbrCntConfig: bad kind%s");
    if (!(q->tauI_DO != 0.0)) return fail(SBR_ERR_ARG, "SbrCntConfig: tauI_DO must be non-zero%s");
    if (sbr::cnt_has_ec(q->kind) && !(q->tauI_EC != 0.0)) return fail(SBR_ERR_ARG, "SbrCntConfig: tauI_EC must be non-zero%s");
    if ((q->kind == SBR_CNT_V1 || q->kind == SBR_CNT_V2) && !(q->tm2_1 > q->tm2_0 && q->tm4_0 > q->tm2_1))
        return fail(SBR_ERR_ARG, "SbrCntConfig: phase stamps tm2_0 < tm2_1 < tm4_0 required%s");
    return SBR_OK;
}

int sbr_cnt_reset(int64_t n, int64_t ld, const SbrCntConfig* cfg, const double* x0, const double* influent,
                  const uint8_t* mask, const SbrParams* p, const SbrOsSchedule* s, double* st, double* obs,
                  uint8_t* done, int32_t* status, uint32_t* counters, int mode, const SbrTol* tol, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if ((rc = check_os_schedule(s))) return rc;
    if ((rc = check_cnt_config(cfg))) return rc;
    if (!influent || !st || !obs || !done) return fail(SBR_ERR_ARG, "sbr_cnt_reset: NULL buffer%s");
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "sbr_cnt_reset: bad mode%s");
    CntArgs g{n, ld, st, x0, influent, mask, nullptr, obs, nullptr, done, status, counters};
    g.K = 1;
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const sbr::CntCfg q = sbr::make_cnt_cfg(*cfg);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t cs = (cudaStream_t)stream;
    if (mode == SBR_MODE_RK4) sbr_cnt_reset_kernel<SBR_MODE_RK4><<<grid, kBlock, 0, cs>>>(g, q, *p, c, *s, t);
    else sbr_cnt_reset_kernel<SBR_MODE_DP45><<<grid, kBlock, 0, cs>>>(g, q, *p, c, *s, t);
    return check_launch("sbr_cnt_reset");
}

namespace {
int cnt_step_launch(const char* what, int64_t n, int64_t ld, int K, const SbrCntConfig* cfg, double* st, const double* action,
                    double* action_io, const SbrPolicyMlp* policy, const SbrParams* p, const SbrOsSchedule* s, double* obs,
                    double* reward, uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log,
                    int mode, const SbrTol* tol, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if ((rc = check_os_schedule(s))) return rc;
    if ((rc = check_cnt_config(cfg))) return rc;
    if (!st || !(action || action_io) || !reward || !done) return fail(SBR_ERR_ARG, "%s: NULL buffer", what);
    if (K < 1 || K > 4096) return fail(SBR_ERR_ARG, "%s: K must be in 1..4096", what);
    if (mode != SBR_MODE_RK4 && mode != SBR_MODE_DP45) return fail(SBR_ERR_ARG, "%s: bad mode", what);
    CntArgs g{n, ld, st, nullptr, nullptr, nullptr, action, obs, reward, done, status, counters};
    g.K = K;
    g.pol_w1 = nullptr; g.pol_w2 = nullptr; g.pol_lo = nullptr; g.pol_span = nullptr; g.pol_hidden = 0;
    g.action_io = action_io; g.act_log = act_log; g.obs_log = obs_log;
    if (policy) {
        if (!policy->w1 || !policy->w2 || !policy->lo || !policy->span) return fail(SBR_ERR_ARG, "%s: NULL policy", what);
        if (policy->n_in != cnt_policy_inputs(cfg->kind) || policy->n_out != (cfg->kind == SBR_CNT_OS2 ? 2 : 1) ||
            policy->hidden < 1 || policy->hidden > kPolMaxHidden)
            return fail(SBR_ERR_ARG, "%s: the policy must map the kind's observation rows (7 / 5 / 18) to 1 (SBROS-v2: 2) "
                                     "actions, hidden <= 64", what);
        g.pol_w1 = policy->w1; g.pol_w2 = policy->w2; g.pol_lo = policy->lo; g.pol_span = policy->span;
        g.pol_hidden = policy->hidden;
    }
    const SbrTol t = tol_or_default(tol);
    const sbr::Coef c = sbr::make_coef(*p);
    const sbr::CntCfg q = sbr::make_cnt_cfg(*cfg);
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    cudaStream_t cs = (cudaStream_t)stream;
    if (policy) {
        if (mode == SBR_MODE_RK4) sbr_cnt_step_kernel<SBR_MODE_RK4, true><<<grid, kBlock, 0, cs>>>(g, q, *p, c, *s, t);
        else sbr_cnt_step_kernel<SBR_MODE_DP45, true><<<grid, kBlock, 0, cs>>>(g, q, *p, c, *s, t);
    } else {
        if (mode == SBR_MODE_RK4) sbr_cnt_step_kernel<SBR_MODE_RK4, false><<<grid, kBlock, 0, cs>>>(g, q, *p, c, *s, t);
        else sbr_cnt_step_kernel<SBR_MODE_DP45, false><<<grid, kBlock, 0, cs>>>(g, q, *p, c, *s, t);
    }
    return check_launch(what);
}
}  // namespace

int sbr_cnt_step(int64_t n, int64_t ld, const SbrCntConfig* cfg, double* st, const double* action, const SbrParams* p,
                 const SbrOsSchedule* s, double* obs, double* reward, uint8_t* done, int32_t* status,
                 uint32_t* counters, int mode, const SbrTol* tol, void* stream) {
    return cnt_step_launch("sbr_cnt_step", n, ld, 1, cfg, st, action, nullptr, nullptr, p, s, obs, reward, done, status,
                           counters, nullptr, nullptr, mode, tol, stream);
}

int sbr_cnt_rollout_k(int64_t n, int64_t ld, int K, const SbrCntConfig* cfg, double* st, double* action,
                      const SbrPolicyMlp* policy, const SbrParams* p, const SbrOsSchedule* s, double* obs, double* reward,
                      uint8_t* done, int32_t* status, uint32_t* counters, double* act_log, double* obs_log, int mode,
                      const SbrTol* tol, void* stream) {
    if (!policy) return fail(SBR_ERR_ARG, "sbr_cnt_rollout_k: NULL policy%s");
    return cnt_step_launch("sbr_cnt_rollout_k", n, ld, K, cfg, st, nullptr, action, policy, p, s, obs, reward, done, status,
                           counters, act_log, obs_log, mode, tol, stream);
}

int sbr_policy_mlp(int64_t n, int64_t ld, const double* obs_a, int rows_a, const double* obs_b, int rows_b,
                   const float* w1, const float* w2, const float* lo, const float* span, int hidden, int n_out,
                   double* action, void* stream) {
    if (n <= 0) return fail(SBR_ERR_ARG, "n must be positive%s");
    if (ld < n) return fail(SBR_ERR_ARG, "ld must be >= n%s");
    if (!obs_a || !w1 || !w2 || !lo || !span || !action || (rows_b > 0 && !obs_b))
        return fail(SBR_ERR_ARG, "sbr_policy_mlp: NULL buffer%s");
    if (rows_a < 1 || rows_b < 0 || rows_a + rows_b > kPolMaxIn || hidden < 1 || hidden > kPolMaxHidden || n_out < 1 ||
        n_out > kPolMaxOut)
        return fail(SBR_ERR_ARG, "sbr_policy_mlp: sizes out of range (inputs <= 40, hidden <= 64, outputs <= 4)%s");
    PolicyArgs g{n, ld, obs_a, obs_b, w1, w2, lo, span, action, rows_a, rows_b, hidden, n_out};
    sbr_policy_mlp_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g);
    return check_launch("sbr_policy_mlp");
}

int sbr_influent_mix(int64_t n, int64_t ld, const double* rnd, const double* mean, const double* std,
                     double* influent, void* stream) {
    if (n <= 0) return fail(SBR_ERR_ARG, "n must be positive%s");
    if (ld < n) return fail(SBR_ERR_ARG, "ld must be >= n%s");
    if (!rnd || !mean || !std || !influent) return fail(SBR_ERR_ARG, "sbr_influent_mix: NULL buffer%s");
    const unsigned grid = (unsigned)((n + 127) / 128);
    sbr_influent_mix_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n, ld, rnd, mean, std, influent);
    return check_launch("sbr_influent_mix");
}

int sbr_influent_sample(int64_t n, int64_t ld, uint64_t seed, int64_t env_offset, int64_t* epoch, int64_t epoch0,
                        int scenario, const double* mean, const double* std, const uint8_t* mask, double* influent,
                        int32_t* scenario_out, void* stream) {
    if (n <= 0) return fail(SBR_ERR_ARG, "n must be positive%s");
    if (ld < n) return fail(SBR_ERR_ARG, "ld must be >= n%s");
    if (!mean || !std || !influent) return fail(SBR_ERR_ARG, "sbr_influent_sample: NULL buffer%s");
    if (scenario < -1 || scenario > 7) return fail(SBR_ERR_ARG, "sbr_influent_sample: scenario must be -1..7%s");
    if (env_offset < 0) return fail(SBR_ERR_ARG, "sbr_influent_sample: env_offset must be >= 0%s");
    SampleArgs g{n, ld, env_offset, epoch0, seed, epoch, mask, mean, std, influent, scenario_out, scenario};
    sbr_influent_sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g);
    return check_launch("sbr_influent_sample");
}

int sbr_philox_normals(int64_t n, int64_t ld, uint64_t seed, int64_t env_offset, int64_t epoch0, double* z,
                       void* stream) {
    if (n <= 0 || ld < n || !z || env_offset < 0) return fail(SBR_ERR_ARG, "sbr_philox_normals: bad arguments%s");
    sbr_philox_normals_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, ld, seed, env_offset,
                                                                                               epoch0, z);
    return check_launch("sbr_philox_normals");
}

int sbr_permute_rows(int64_t n, const int64_t* perm, int nbuf, const void* const* src, void* const* dst,
                     const int64_t* ld_src, const int64_t* ld_dst, const int32_t* rows, const int32_t* elem_bytes,
                     int scatter, void* stream) {
    if (n <= 0 || !perm || !src || !dst || !ld_src || !ld_dst || !rows || !elem_bytes)
        return fail(SBR_ERR_ARG, "sbr_permute_rows: bad arguments%s");
    if (nbuf < 1 || nbuf > SBR_PERMUTE_MAX) return fail(SBR_ERR_ARG, "sbr_permute_rows: 1..SBR_PERMUTE_MAX buffers%s");
    PermuteArgs g;
    memset(&g, 0, sizeof(g));
    g.n = n; g.perm = perm; g.nbuf = nbuf; g.scatter = scatter ? 1 : 0;
    int total = 0;
    for (int k = 0; k < nbuf; ++k) {
        if (!src[k] || !dst[k] || rows[k] < 1 || ld_src[k] < n || ld_dst[k] < n || (elem_bytes[k] != 4 && elem_bytes[k] != 8))
            return fail(SBR_ERR_ARG, "sbr_permute_rows: bad buffer descriptor%s");
        g.src[k] = src[k]; g.dst[k] = dst[k]; g.ld_src[k] = ld_src[k]; g.ld_dst[k] = ld_dst[k];
        g.elem[k] = elem_bytes[k]; g.row0[k] = total;
        total += rows[k];
    }
    g.row0[nbuf] = total;
    if (total > 65535) return fail(SBR_ERR_ARG, "sbr_permute_rows: too many rows%s");
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)total);
    sbr_permute_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    return check_launch("sbr_permute_rows");
}

int sbr_settle_draw(int64_t n, int64_t ld, double* x, double settle_time, const SbrParams* p, double* sX, double* out,
                    int32_t* status, void* stream) {
    int rc = check_common(n, ld, p);
    if (rc) return rc;
    if (!x || !sX || !out) return fail(SBR_ERR_ARG, "sbr_settle_draw: NULL buffer%s");
    if (!(settle_time > 0.0)) return fail(SBR_ERR_ARG, "sbr_settle_draw: settle_time must be positive%s");
    sbr_settle_draw_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, ld, x, settle_time, *p, sX,
                                                                                          out, status);
    return check_launch("sbr_settle_draw");
}

int sbr_reward_stats_init(double* stats, void* stream) {
    if (!stats) return fail(SBR_ERR_ARG, "sbr_reward_stats_init: NULL buffer%s");
    sbr_reward_stats_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stats);
    return check_launch("sbr_reward_stats_init");
}

int sbr_reward_stats(int64_t n, const double* reward, const int32_t* status, double* stats, void* stream) {
    if (n <= 0 || !reward || !stats) return fail(SBR_ERR_ARG, "sbr_reward_stats: bad arguments%s");
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    sbr_reward_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, reward, status, stats);
    return check_launch("sbr_reward_stats");
}

int sbr_fp64_probe(int blocks, int threads, int iters, double* sink, double* flops, void* stream) {
    if (blocks < 1 || threads < 1 || threads > 1024 || iters < 1 || !sink)
        return fail(SBR_ERR_ARG, "sbr_fp64_probe: bad arguments%s");
    sbr_fp64_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    if (flops) *flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    return check_launch("sbr_fp64_probe");
}

}  // extern "C"
