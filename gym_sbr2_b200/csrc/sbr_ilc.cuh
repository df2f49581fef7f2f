// sbr_ilc.cuh -- per-environment arithmetic of the reference's batch-to-batch (iterative-learning) feed-forward KLa
// path, the research feature of `SBR-v0`:
//   cycle 0 (no feed-forward)      SBR_model_PID_on.run         + sub_phases_PID_on.py:178-271, 406-500
//   cycle k (feed-forward + PID)   SBR_model_batchPID_fbPID.run + sub_phases_batchPID_fbPID.py:139-253, 388-500
//   batch-to-batch controller      module_batch_PID.batch_PID     (module_batch_PID.py:7-275)
// Same kinetics, tails, RK4 stepper and settler as the cycle-per-step path (sbr_core.cuh).  What differs:
//   * the DO controller samples So at EVERY output point of the reference's odeint grid (9, phase 5: 10 per PID
//     interval) into a per-env memory -- the batch-to-batch controller works on those.  MODE DP45 (default): adaptive
//     Dormand-Prince over the PID interval, the memory filled from the method's continuous extension
//     (dp45_interval_dense); MODE RK4: one RK4 step per output point -- cycle 0 starts aerated at KLa 240 and So then
//     collapses from 7 to 0.004 g/m3 within a few points of the fill phase, where RK4 on the grid is 2e-5 g/m3 off in
//     the memory although the end state agrees;
//   * KLa of an interval = feedback PID + clamp(u_batch + KLa memory of cycle 0) read at list position 9 i + 1 (:230);
//   * the feedback bias starts from 0 (feed-forward cycles, :173-232: Kla[0] is never seeded) or from the incoming KLa
//     (cycle 0, sub_phases_PID_on.py:218);
//   * Qw follows in closed form from the biomass set-point (SBR_model_batchPID_fbPID.py:283-291) and the decant
//     rescales the particulates to the mixed residual solids (sub_phases_batchPID_fbPID.py:784-809).
// `SbrEnv.step` itself cannot run in the reference (float linspace counts; a seven-argument call of the ten-parameter
// reward, gym_SBR_env0.py:203), so parity is pinned function by function (oracle/make_golden_ilc.py) and the reward
// written here is module_reward.sbr_reward's formula on this cycle's applied KLa -- by construction, not pinned.
// Reference behaviour is cited as file:line relative to /root/reference/gym_SBR/envs.
#pragma once
#include "sbr_core.cuh"

namespace sbr {

struct IlcIo {
    Column so;         // out [S]: So at every output sample, the six PID-controlled phases concatenated (off[])
    Column kla_mem;    // out [S] (may be NULL): cycle 0: the feedback KLa per sample (the later feed-forward base);
                       //                        feed-forward cycles: the clamped feed-forward profile (Kla_memory, :173-194)
    Column kla_base;   // in  [S]: KLa memory of cycle 0 (feed-forward cycles only)
    Column u;          // in  [S]: u_batch of the batch-to-batch controller (feed-forward cycles only)
};

struct IlcOut {
    double Qeff, Qw, reward, OCI;
    double kla_mean[3];    // mean applied KLa of phases 3, 5, 8
    int status;
};

// Dormand-Prince 5(4) over ONE PID interval [0, T] with the So memory filled from the method's own continuous extension
// (Hairer, Norsett & Wanner II.6, the 4th-order interpolant of dopri5: local error O(h^5) like the step itself, measured
// against LSODA at 1e-12 in tests/test_twin_parity_ilc.py): output point p = 0 .. m-1 at time (p + 1) T / m lands in
// so[j0 + p].  Same controller and constants as dp45_interval (sbr_core.cuh), whose stage vectors are all alive at the
// accept -- this copy adds the six-term sum of the So component and the Horner evaluation per output point, so that the
// stepper no longer stops at every point (7.3 right-hand sides per point, first-same-as-last lost each time).
struct DpDense {
    double d1, d3, d4, d5, d6, d7;
};
#define SBR_DP_DENSE                                                                                              \
    {-12715105075.0 / 11282082432.0, 87487479700.0 / 32700410799.0, -10690763975.0 / 1880347072.0,                 \
     701980252875.0 / 199316789632.0, -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0}

template <int TAIL>
SBR_HD int dp45_interval_dense(double (&x)[SBR_NX], double T, int m, const Flow& f, const Coef& c, const TailArgs& a,
                               const SbrTol& tol, Dp45State& st, double& xpq, const Column& so, int j0) {
#ifdef __CUDA_ARCH__
    const DpTab& tb = kDpTab;
#else
    const DpTab tb = SBR_DP_TABLEAU;
#endif
    const DpDense dd = SBR_DP_DENSE;
    double k1[SBR_NX], k2[SBR_NX], k3[SBR_NX], k4[SBR_NX], k5[SBR_NX], k6[SBR_NX], y[SBR_NX];
    double t = 0.0;
    double h = st.h * SBR_DP_FIRST;
    int status = 0, steps = 0, p = 0;
    const double h_out = T / (double)m;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) y[i] = x[i];
    double g1 = stage<TAIL>(y, k1, 0.0, f, c, a);
    st.n_rhs += 1;
    while (t < T) {
        if (steps >= tol.max_steps) { status = SBR_ST_STEPLIMIT; break; }
        ++steps;
        const double rem = T - t;
        const float n_f = ceilf((float)rem * frcp_fast((float)h) * SBR_DP_NGUARD);
        const bool last = !(n_f > 1.0f);
        const double hs = last ? rem : rem * (double)frcp_fast(n_f);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) y[i] = fma(hs * tb.a21, k1[i], x[i]);
        stage<TAIL>(y, k2, fma(tb.c2, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) y[i] = fma(hs * tb.a32, k2[i], fma(hs * tb.a31, k1[i], x[i]));
        const double g3 = stage<TAIL>(y, k3, fma(tb.c3, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) y[i] = fma(hs * tb.a43, k3[i], fma(hs * tb.a42, k2[i], fma(hs * tb.a41, k1[i], x[i])));
        const double g4 = stage<TAIL>(y, k4, fma(tb.c4, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i))
                y[i] = fma(hs * tb.a54, k4[i], fma(hs * tb.a53, k3[i], fma(hs * tb.a52, k2[i], fma(hs * tb.a51, k1[i], x[i]))));
        const double g5 = stage<TAIL>(y, k5, fma(tb.c5, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i))
                y[i] = fma(hs * tb.a65, k5[i], fma(hs * tb.a64, k4[i], fma(hs * tb.a63, k3[i],
                       fma(hs * tb.a62, k2[i], fma(hs * tb.a61, k1[i], x[i])))));
        const double g6 = stage<TAIL>(y, k6, t + hs, f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i))
                y[i] = fma(hs * tb.b6, k6[i], fma(hs * tb.b5, k5[i], fma(hs * tb.b4, k4[i],
                       fma(hs * tb.b3, k3[i], fma(hs * tb.b1, k1[i], x[i])))));
        const double g7 = stage<TAIL>(y, k2, t + hs, f, c, a);      // k7 = f(5th-order solution) lands in k2
        st.n_rhs += 6;
        double en3[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) {
                const double err = fma(tb.e7, k2[i], fma(tb.e6, k6[i], fma(tb.e5, k5[i], fma(tb.e4, k4[i],
                                   fma(tb.e3, k3[i], tb.e1 * k1[i])))));
                const double sc = fma(tol.rtol, fabs(y[i]), tol.atol * tol_scale(i));
                const double q = err * rcp_rough(sc);
                en3[aidx(i) % 3] = fma(q, q, en3[aidx(i) % 3]);
            }
        const double en = ((en3[0] + en3[1]) + en3[2]) * (hs * hs * (1.0 / 9));
        const bool finite = en < 1e300;
        if (en <= 1.0 || !finite) {
            const double t_new = (last || !finite) ? T : t + hs;
            // continuous extension of the So component over [t, t + hs]
            const double r2 = y[iSo] - x[iSo];
            const double r3 = fma(hs, k1[iSo], -r2);
            const double r4 = r2 - hs * k2[iSo] - r3;
            const double r5 = hs * fma(dd.d7, k2[iSo], fma(dd.d6, k6[iSo], fma(dd.d5, k5[iSo], fma(dd.d4, k4[iSo],
                              fma(dd.d3, k3[iSo], dd.d1 * k1[iSo])))));
            const double inv_hs = 1.0 / hs;
            while (p < m) {
                const double tp = (double)(p + 1) * h_out;
                if (p + 1 < m ? tp > t_new : t_new < T) break;        // the interval's last point goes with its last step
                const double th = p + 1 < m ? (tp - t) * inv_hs : 1.0;
                const double th1 = 1.0 - th;
                so.set(j0 + p, fma(th, fma(th1, fma(th, fma(th1, r5, r4), r3), r2), x[iSo]));
                ++p;
            }
            t = t_new;
            xpq = fma(hs, fma(tb.b6, g6, fma(tb.b5, g5, fma(tb.b4, g4, fma(tb.b3, g3, tb.b1 * g1)))), xpq);
            g1 = g7;
#pragma unroll
            for (int i = 0; i < SBR_NX; ++i)
                if (active(i)) { x[i] = y[i]; k1[i] = k2[i]; }
        } else {
            st.n_rej += 1;
        }
        float fac = SBR_DP_MAXGROW;
        if (en > 1e-20) {
            fac = SBR_DP_SAFETY * pow_m01((float)en);
            fac = fminf(SBR_DP_MAXGROW, fmaxf(0.2f, fac));
        }
        if (en > 1.0) fac = fminf(fac, 1.0f);
        if (!(last && en <= 1.0)) h = hs * (double)fac;
        else h = fmax(h, hs * (double)fac);
    }
    // an interval given up at the step limit still fills its memory (with the last state): the caller flags the env
    while (p < m) { so.set(j0 + p, x[iSo]); ++p; }
    st.h = h;
    return status;
}

// One PID-controlled phase on the reference's output grid.  ff: feed-forward cycle (else cycle 0).  kla_carry: in = KLa the
// phase starts from (cycle 0 only), out = the phase's last feedback KLa.  off = sample offset of the phase in the memories.
template <int TAIL, int MODE>
SBR_HD int ilc_phase(double (&x)[SBR_NX], int n_int, int m, double T, double sp, bool ff, double& kla_carry,
                     const Coef& c, TailArgs a, const PidA& pid, const SbrTol& tol, Dp45State& st, const IlcIo& io,
                     int off, double& kla_sum) {
    double bias = ff ? 0.0 : kla_carry, ie = 0.0, so_prev = 0.0, so_i = x[iSo];
    double ksum = 0.0, kla_fb = kla_carry;
    int status = 0;
    io.so.set(off, x[iSo]);
    io.kla_mem.set(off, ff ? io.kla_base.get(off) : kla_carry);
    const double h = T / (double)m;
    // Feed-forward KLa of interval i = Kla_memory[9 i + 1] of a list that holds m entries per interval: entry q - 1 = (i', ii')
    // was built from index 9 i' + ii' + 1 of u_batch and of the cycle-0 memory (sub_phases_batchPID_fbPID.py:177-194, 230).
    // The two loads of interval i + 1 are issued before the solve of interval i, off the PID -> KLa -> first stage chain
    // (measured: 27.28 against 27.23 ms per 2^17 cycles -- the other resident warp already covered them; kept, it costs
    // nothing).
    auto feed_forward = [&](int i) {
        const int q = 9 * i;
        const int v = off + 9 * (q / m) + (q % m) + 1;
        return clip_keep_nan(io.u.get(v) + io.kla_base.get(v), pid.lo, pid.hi);
    };
    double ff_i = ff ? feed_forward(0) : 0.0;
    for (int i = 0; i < n_int; ++i) {
        kla_fb = pid_a_update(pid, sp, so_i, so_prev, i == 0, ie, bias);
        const double kla = kla_fb + ff_i;
        if (ff && i + 1 < n_int) ff_i = feed_forward(i + 1);
        a.kla = kla;
        a.kla_sat = kla * c.so_sat;
        const Flow f{x[iV], TAIL == TAIL_REACT ? 0.0 : a.q};
        const double snh0 = x[iSnh], sno0 = x[iSno];
        double xpq = 0.0;
        if (MODE == SBR_MODE_RK4) {
            for (int s = 0; s < m; ++s) {
                rk4_step<TAIL>(x, (double)s * h, h, f, c, a, xpq);
                st.n_rhs += 4;
                io.so.set(off + 1 + i * m + s, x[iSo]);
            }
        } else {
            status |= dp45_interval_dense<TAIL>(x, T, m, f, c, a, tol, st, xpq, io.so, off + 1 + i * m);
        }
        if (io.kla_mem.p) {
            for (int s = 0; s < m; ++s) {
                const int v = off + 9 * i + s + 1;
                io.kla_mem.set(off + 1 + i * m + s,
                               ff ? clip_keep_nan(io.u.get(v) + io.kla_base.get(v), pid.lo, pid.hi) : kla_fb);
            }
        }
        // passive components of the interval (closed forms, see sbr_core.cuh `active`)
        if (TAIL == TAIL_REACT) {
            x[iXp] = fma(c.ixp, xpq, x[iXp]);
            x[iSalk] += ((x[iSnh] - snh0) - (x[iSno] - sno0)) * c.c136;
        } else {
            const double V1 = f.V(T), qT = f.q * T;
            const double w = f.V0 / V1, u = qT / V1;
            const double D0 = x[iSalk];
            x[iV] = V1;
            x[iSi] = fma(x[iSi], w, cin<TAIL>(a, iSi) * u);
            x[iXi] = fma(x[iXi], w, cin<TAIL>(a, iXi) * u);
            x[iXp] = fma(x[iXp], w, fma(cin<TAIL>(a, iXp), u, c.ixp * xpq / V1));
            const double n0 = (snh0 - sno0) * c.c136;
            const double n_in = (cin<TAIL>(a, iSnh) - cin<TAIL>(a, iSno)) * c.c136;
            const double D1 = fma(D0 - n0, w, (cin<TAIL>(a, iSalk) - n_in) * u);
            x[iSalk] = D1 + (x[iSnh] - x[iSno]) * c.c136;
        }
        ksum += kla;
        so_prev = so_i;
        so_i = x[iSo];
    }
    kla_carry = kla_fb;
    kla_sum = ksum;
    return status;
}

// drawing.sim_drawing of the PID_on / batchPID model files (sub_phases_batchPID_fbPID.py:784-809).
SBR_HD void draw_fixed_qw(double (&x)[SBR_NX], const double (&sX)[10], double Xf, double Qeff, double Qw) {
    const double V0 = x[iV];
    const double V = V0 - Qeff - Qw;
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) tot += sX[i];
    const double sX2 = (tot * V0 / 10 - Qw * sX[0] - Qeff * sX[9]) / V;
    x[iV] = V;
    x[iXs] = (0.75 * x[iXs] / Xf) * sX2;
    x[iXp] = (0.75 * x[iXp] / Xf) * sX2;
    x[iXi] = (0.75 * x[iXi] / Xf) * sX2;
    x[iXbh] = (0.75 * x[iXbh] / Xf) * sX2;
    x[iXba] = (0.75 * x[iXba] / Xf) * sX2;
}

// Whole cycle.  sp8: DO set-points of the 8 phases; off[6]: sample offsets of phases 1, 2, 3, 4, 5, 8; t_fill: length
// of the fill phase (Qin = q_in t_fill, SBR_model_batchPID_fbPID.py:25-27).  p carries this path's controller and plant
// constants (SbrIlcVecEnv: Kc 0.5/1.18, tauI 0.0015, tauD 0.005, PID dt 0.05, biomass set-point 5400, IV 0.66).
template <int MODE>
SBR_HD void cycle_ilc(double (&x)[SBR_NX], const double (&sp8)[8], Loading load, double q_fill, double t_fill, bool ff,
                      const SbrParams& p, const Coef& c, const SbrSchedule& s, const SbrTol& tol, Dp45State& st,
                      const IlcIo& io, const int (&off)[6], IlcOut& o) {
    const PidA pid = make_pid_a(p, c);
    TailArgs a;
    a.kla = 0.0; a.q = q_fill; a.ec_conc = 0.0; a.load = load;
    double kla = p.kla_max;                         // SBR_model_PID_on.py:146: kla0 = DO_control_par[5]
    double ksum = 0.0, kla5 = 0.0;
    int status = ilc_phase<TAIL_FILL, MODE>(x, s.n_int[0], s.n_sub[0], s.interval[0], sp8[0], ff, kla, c, a, pid, tol, st,
                                            io, off[0], ksum);
    a.q = 0.0;
    for (int j = 0; j < 5; ++j) {
        const int ph = j < 4 ? j + 1 : 7;
        if (j == 4) {
            kla5 = kla;
            double sX[10], Xf;
            settle_closed_form(x, s.settle_time, p.settler_area, p.settler_vmax, sX, Xf);
            double tot = 0.0;
#pragma unroll
            for (int i = 0; i < 10; ++i) tot += sX[i];
            const double qin = (p.WV - p.IV) / t_fill;
            o.Qw = (tot * p.WV / 10 - p.biomass_setpoint * (p.WV - qin * t_fill) - qin * t_fill * sX[9]) / (sX[0] - sX[9]);
            o.Qeff = qin * t_fill - o.Qw;
            draw_fixed_qw(x, sX, Xf, o.Qeff, o.Qw);
        }
        double kc = j == 4 ? kla5 : kla;
        status |= ilc_phase<TAIL_REACT, MODE>(x, s.n_int[ph], s.n_sub[ph], s.interval[ph], sp8[ph], ff, kc, c, a, pid, tol,
                                              st, io, off[j + 1], ksum);
        if (j < 4) kla = kc;
        const double mean = ksum / (double)s.n_int[ph];
        if (j == 1) o.kla_mean[0] = mean;
        if (j == 3) o.kla_mean[1] = mean;
        if (j == 4) o.kla_mean[2] = mean;
    }
    // module_reward.sbr_reward's formula (module_reward.py:4-51) on this cycle's applied KLa, effluent flow and the end
    // state's ammonia -- NOT pinned: the reference's call site cannot run (header)
    const double ME = 0.005 * 1.32 * 24 + 0.005 * 1.32 * 24;
    const double AE = p.so_sat / (1.8 * 1000) * (1.32 * o.kla_mean[0] + 1.32 * o.kla_mean[1] + (1.32 - o.Qw) * o.kla_mean[2]);
    const double PE = 0.004 * (p.WV - p.IV) + 0.05 * o.Qw + 0.004 * o.Qeff;
    o.OCI = AE + PE + ME;
    o.reward = (5 - o.OCI) + (x[iSnh] < 4 ? 0.0 : -20.0);
    bool finite = true;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) finite = finite && (fabs(x[i]) < 1e300);
    o.status = status | (finite ? 0 : SBR_ST_NONFINITE);
}

// ---------------------------------------------------------------------------------------------------------
// Batch-to-batch controller of ONE phase of one env (module_batch_PID.py:20-52, 214-270).
//   E(t) = sum_{j in [t, min(t + tp, n))} (sp - So_j) w_j dt / D(t),   D(t) = sum w_j dt (host, the reference's own sums)
//   u(t) = Kc E(t) + (Kc / tauI) sum_over_cycles E(t) + Kc tauD (E(t) - E_previous_cycle(t))
// The window sum runs as a backward recursion N(t) = N(t+1) + term(t) - term(t + tp): the weights decay like
// exp(-t / tau) with tp = 3 tau / dt, so going backward the sum grows and what is subtracted is e^-3 of it -- no
// cancellation (checked against the reference's direct sums in tests/).  w, D: shared by all envs.
// ---------------------------------------------------------------------------------------------------------
SBR_HD void ilc_update_phase(int n, int tp, double sp, double dt, const double* w, const double* D, const Column& so,
                             const Column& e_sum, const Column& e_last, const Column& u, double Kc, double KcI,
                             double KcD) {
    double N = 0.0;
    for (int t = n - 1; t >= 0; --t) {
        N += ((sp - so.get(t)) * w[t]) * dt;
        if (t + tp < n) N -= ((sp - so.get(t + tp)) * w[t + tp]) * dt;
        const double E = N / D[t];
        const double prev = e_last.get(t);
        const double acc = e_sum.get(t) + E;
        e_sum.set(t, acc);
        e_last.set(t, E);
        u.set(t, Kc * E + KcI * acc + KcD * (E - prev));
    }
}

}  // namespace sbr
