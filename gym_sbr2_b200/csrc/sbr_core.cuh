// sbr_core.cuh -- per-environment arithmetic of the batched SBR stepper (one env = one thread).
//
// Everything here is a pure function of registers: the kinetic right-hand side with its three tails, the
// RK4 and Dormand-Prince steppers, the DO->KLa PID, the settler closed form, the draw/waste algebra and the
// reward/observation epilogues.  The kernels in sbr_kernels.cu only add SoA loads/stores around these.
// The functions are SBR_HD (host+device) so that tests can compile the same source with g++ into a CPU twin
// (oracle/twin/, test infrastructure) and debug the logic without a GPU; the product only ships the CUDA build.
//
// Reference behaviour is cited as file:line relative to /root/reference/gym_SBR/envs.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/sbr_b200.h"

#ifdef __CUDACC__
#define SBR_HD __host__ __device__ __forceinline__
#define SBR_HD_NOINLINE __host__ __device__
#else
#define SBR_HD inline
#define SBR_HD_NOINLINE inline
#endif

namespace sbr {

enum { TAIL_REACT = 0, TAIL_FILL = 1, TAIL_EC = 2 };
enum { iV = 0, iSi, iSs, iXi, iXs, iXbh, iXba, iXp, iSo, iSno, iSnh, iSnd, iXnd, iSalk };

// ---------------------------------------------------------------------------------------------------------
// Derived constants: stoichiometric coefficients folded on the host with the reference's own expressions
// (sub_phases_FB.py:307-343).  Passed by value as a kernel argument (constant bank), never a global symbol.
// ---------------------------------------------------------------------------------------------------------
struct Coef {
    double Ks, Koh, Kno, Knh, Koa, Kx;                 // half-saturation constants
    double muh, etag_Koh, etah_Koh, mua_Ya, bh, ba, ka, kh;   // rate constants (products folded on the host)
    double Ya;            // rho3 = Ya * rho3s
    double n_invYh;       // nu2_1 = nu2_2 = -1/Yh
    double one_m_ixp;     // nu4_4 = nu4_5 = 1 - ixp   (sic: ixp where ASM1 has fp)
    double ixp;           // nu7_4 = nu7_5
    double c81, c83_Ya;   // So:  -(1-Yh)/Yh , -(4.57-Ya)/Ya * Ya
    double c92;           // Sno: -(1-Yh)/(2.86 Yh)   (nu9_3 = 1/Ya is folded into rho3s)
    double n_ixb, c103_Ya;   // Snh: -ixb , (-ixb - 1/Ya) * Ya
    double c124;          // Xnd: ixb - fp*ixp
    double c136;          // Salk: 1/14 -- nu13_k == (nu10_k - nu9_k)/14 for every process k (charge balance)
    double so_sat;
    // controller gains folded on the host (every env needs them in every launch; an IEEE divide is ~15 instructions
    // and a ~150-cycle dependency chain on the device): Kc/tauI, Kc*tauD and 1/dt of the three PIDs
    double pidA_KcI, pidA_KcD, pidA_inv_dt;                    // cycle-per-step DO-PID (also SBR-v4's gains)
    double os_KcI_DO, os_KcD_DO, os_KcI_EC, os_KcD_EC, os_inv_dt;   // SBROS-v1 DO- and NO3-PID; os_inv_dt also SBR-v4
};

inline Coef make_coef(const SbrParams& p) {
    Coef c;
    c.Ks = p.Ks; c.Koh = p.Koh; c.Kno = p.Kno; c.Knh = p.Knh; c.Koa = p.Koa; c.Kx = p.Kx;
    c.muh = p.muh; c.etag_Koh = p.etag * p.Koh; c.etah_Koh = p.etah * p.Koh; c.mua_Ya = p.mua * (1 / p.Ya);
    c.bh = p.bh; c.ba = p.ba; c.ka = p.ka; c.kh = p.kh;
    c.Ya = p.Ya;
    c.n_invYh = -1 / p.Yh;
    c.one_m_ixp = 1 - p.ixp;
    c.ixp = p.ixp;
    c.c81 = -(1 - p.Yh) / p.Yh;
    c.c83_Ya = (-(4.57 - p.Ya) / p.Ya) * p.Ya;
    c.c92 = -((1 - p.Yh) / (2.86 * p.Yh));
    c.n_ixb = -p.ixb;
    c.c103_Ya = (-p.ixb - 1 / p.Ya) * p.Ya;
    c.c124 = p.ixb - p.fp * p.ixp;
    c.c136 = 1.0 / 14;
    c.so_sat = p.so_sat;
    c.pidA_KcI = p.pid_Kc / p.pid_tauI; c.pidA_KcD = p.pid_Kc * p.pid_tauD; c.pidA_inv_dt = 1.0 / p.pid_dt;
    c.os_KcI_DO = p.os_Kc_DO / p.os_tauI_DO; c.os_KcD_DO = p.os_Kc_DO * p.os_tauD_DO;
    c.os_KcI_EC = p.os_Kc_EC / p.os_tauI_EC; c.os_KcD_EC = p.os_Kc_EC * p.os_tauD_EC;
    c.os_inv_dt = 1.0 / p.os_pid_dt;
    return c;
}

// ---------------------------------------------------------------------------------------------------------
// FP64 reciprocal: MUFU.RCP64H seed (~2^-19) + one Newton step (2 DFMA) -> relative error <= ~2^-36.
// An IEEE divide costs ~10 FP64-pipe slots; the RHS has 6-7 of them per evaluation.
// SBR_RCP_NEWTON 2 (default): quadratic correction r (1 + e).  Measured on 2^20 whole cycles against the cubic
//   variant (profiles/r01f_ab_cycle_rcp_variants.log): RHS relative error 2e-11, x_last moves by at most 2.4e-11 relative =
//   2e-6 of ONE parity tolerance unit, 2500x below the RK4 truncation error on the reference grid (0.005 units) and
//   below the rounding noise of the reference's own LSODA run -- for 4 % less kernel time (88.5 -> 85.0 ms).
// SBR_RCP_NEWTON 3: cubic correction r (1 + e + e^2), 3 DFMA, ~2^-57.
// Tried and rejected: a float seed built with integer instructions + MUFU.RCP (2^-22, then ONE Newton step reaches
//   9e-13): beside a saturated FP64 pipe every other instruction still costs ~half an issue cycle and lengthens
//   the dependency chain, so the ~13 extra integer instructions per reciprocal cost more than the DFMA they save
//   (88.5 -> 113.7 ms; tools/issue_probe.cu).
// ---------------------------------------------------------------------------------------------------------
#ifndef SBR_RCP_NEWTON
#define SBR_RCP_NEWTON 2
#endif
SBR_HD double rcp(double d) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
#if SBR_RCP_NEWTON >= 3
    double t = fma(e, e, e);      // cubic: r (1 + e + e^2), error e^3
    return fma(r, t, r);
#else
    return fma(r, e, r);          // quadratic: r (1 + e), error e^2
#endif
#else
    return 1.0 / d;
#endif
}

// Raw MUFU.RCP64H reciprocal (~2^-20 relative): enough for ratios that only steer the step-size controller.
SBR_HD double rcp_rough(double d) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
#else
    return 1.0 / d;
#endif
}

// ---------------------------------------------------------------------------------------------------------
// What the stepper carries.  Only NINE components are dynamically coupled ("active"):
//     Ss, Xs, Xbh, Xba, So, Sno, Snh, Snd, Xnd
// The other five never feed a rate and have closed forms or are plain quadratures, in every tail:
//   V     dV/dt = q (fill flow, carbon-dosing flow, or 0)             -> V(t) = V0 + q t
//   Si,Xi no kinetics (sub_phases_FB.py:348,352): d(x V)/dt = q c_in   -> x(t) = (x0 V0 + q c_in t) / V(t)
//   Salk  every Salk stoichiometric coefficient is (nu_Snh - nu_Sno)/14 (sub_phases_FB.py:337-343 vs :329-336, the
//         charge balance of ASM1), so D = Salk - (Snh - Sno)/14 has no kinetics either: D(t) as Si above, and
//         Salk(t) = D(t) + (Snh(t) - Sno(t))/14
//   Xp    d(Xp V)/dt = ixp (rho4 + rho5) V + q c_in: a quadrature of the active solution, accumulated with the
//         Runge-Kutta weights (needs no stage values)
// Runge-Kutta methods integrate linear invariants and quadratures exactly as they would these components, so the
// result is what carrying all 14 would give up to truncation error of the (now exact) passive parts -- while the
// register footprint of a stage drops from 14 to 9 doubles and the fill / dosing tails cost 9, not 14, dilution FMAs.
// ---------------------------------------------------------------------------------------------------------
SBR_HD constexpr bool active(int i) {
    return i == iSs || i == iXs || i == iXbh || i == iXba || i == iSo || i == iSno || i == iSnh || i == iSnd
        || i == iXnd;
}

// Packed index of an active component (0..8).
SBR_HD constexpr int aidx(int i) {
    return i == iSs ? 0 : i == iXs ? 1 : i == iXbh ? 2 : i == iXba ? 3 : i == iSo ? 4 : i == iSno ? 5
         : i == iSnh ? 6 : i == iSnd ? 7 : 8;
}

// Influent loading accessor: component i at p[i * stride] (shared memory column on the GPU).
struct Loading {
    const double* p;
    int stride;
    SBR_HD double operator()(int i) const { return p[i * stride]; }
};

struct TailArgs {
    double kla;        // oxygen transfer coefficient, constant over one PID interval
    double kla_sat;    // kla * So_sat -- set by integrate_interval()
    double q;          // FILL: inflow m3/d ; EC: carbon dosing flow m3/d
    double ec_conc;    // EC: dosing concentration gCOD/m3
    Loading load;      // FILL: influent concentrations
};

// Inflow concentration of component i: influent (fill), pure carbon into Ss (dosing), nothing (react).
template <int TAIL>
SBR_HD double cin(const TailArgs& a, int i) {
    return TAIL == TAIL_FILL ? a.load(i) : (TAIL == TAIL_EC && i == iSs ? a.ec_conc : 0.0);
}

// ---------------------------------------------------------------------------------------------------------
// Kinetic rates of the active components (sub_phases_FB.py:278-372) + aeration.  CSE-minimal form: 6 Monod
// denominators -> 3 reciprocals (they are only needed as products of two or three); hydrolysis written as
// Xs*g / Xnd*g so no divide by Xs or Xbh is needed ((Xs/Xbh)/(Kx+Xs/Xbh) == Xs/(Kx*Xbh+Xs)).  51 FP64 + 3 MUFU.
// Returns s45 = rho4 + rho5 (the Xp production rate is ixp * s45).
// ---------------------------------------------------------------------------------------------------------
SBR_HD double kinetics(const double (&y)[SBR_NX], double (&k)[SBR_NX], const Coef& c, const TailArgs& a) {
    const double Ss = y[iSs], Xs = y[iXs], Xbh = y[iXbh], Xba = y[iXba], So = y[iSo], Sno = y[iSno],
                 Snh = y[iSnh], Snd = y[iSnd], Xnd = y[iXnd];
    // two pairs of Monod denominators only ever appear as products, so each pair shares ONE reciprocal:
    //   rho1, rho2 ~ 1/((Ks+Ss)(Koh+So))   and   rho3 ~ 1/((Knh+Snh)(Koa+So))
    //   hydrolysis ~ 1/((Koh+So)(Kx Xbh+Xs)): folded into the first pair's reciprocal, R = 1/(d1 d2 d6)
    const double d1 = c.Ks + Ss, d2 = c.Koh + So, d4 = c.Knh + Snh, d5 = c.Koa + So, d6 = fma(c.Kx, Xbh, Xs);
    const double R = rcp((d1 * d2) * d6);
    // both uses of R come with a factor Xbh: RX d6 = Xbh/((Ks+Ss)(Koh+So)), RX d1 = Xbh/((Koh+So)(Kx Xbh+Xs))
    const double RX = R * Xbh;
    const double r45 = rcp(d4 * d5);
    const double r3 = rcp(c.Kno + Sno);
    const double mNo = Sno * r3;                             // Sno/(Kno+Sno)
    const double C = (c.muh * Ss) * (RX * d6);
    const double rho1 = C * So;                              // muh Ss/(Ks+Ss) So/(Koh+So) Xbh
    const double rho2 = (C * c.etag_Koh) * mNo;              // muh Ss/(Ks+Ss) Koh/(Koh+So) Sno/(Kno+Sno) etag Xbh
    // rho3 scaled by nu9_3 = 1/Ya on the host (mua_Ya = mua/Ya), so that d(Sno)/dt takes it without a multiply
    const double rho3s = ((c.mua_Ya * Snh) * So) * (Xba * r45);
    const double rho6 = (c.ka * Snd) * Xbh;
    // hydrolysis: kh (Xs/Xbh)/(Kx+Xs/Xbh) [So/(Koh+So) + etah Koh/(Koh+So) Sno/(Kno+Sno)] Xbh  ==  Xs * g
    const double g = (c.kh * (RX * d1)) * fma(c.etah_Koh, mNo, So);
    const double rho7 = Xs * g;
    const double rho8 = Xnd * g;
    const double s12 = rho1 + rho2;
    const double dXba = c.ba * Xba;                          // rho5
    const double s45 = fma(c.bh, Xbh, dXba);                 // rho4 + rho5 (decay of Xbh and Xba)
    k[iSs] = fma(c.n_invYh, s12, rho7);
    k[iXs] = fma(c.one_m_ixp, s45, -rho7);
    k[iXbh] = fma(-c.bh, Xbh, s12);                          // rho1 + rho2 - rho4
    k[iXba] = fma(c.Ya, rho3s, -dXba);                       // rho3 - rho5
    // aeration KLa (So_sat - So): KLa * So_sat is constant over the PID interval (a.kla_sat)
    k[iSo] = fma(-a.kla, So, fma(c.c81, rho1, fma(c.c83_Ya, rho3s, a.kla_sat)));
    k[iSno] = fma(c.c92, rho2, rho3s);
    k[iSnh] = fma(c.n_ixb, s12, fma(c.c103_Ya, rho3s, rho6));
    k[iSnd] = rho8 - rho6;
    k[iXnd] = fma(c.c124, s45, -rho8);
    return s45;
}

// Volume bookkeeping of one interval: V(t) = V0 + q t.
struct Flow {
    double V0, q;
    SBR_HD double V(double t) const { return fma(q, t, V0); }
};

// Stage derivative of the active components at time offset t: kinetics + dilution (q/V(t)) (c_in - y)
// (sub_phases_FB.py:146-176; gym_SBR_oneshot.py:1757-1787).  Returns the Xp quadrature integrand
// ixp-free: s45 * V(t) (react: s45).
template <int TAIL>
SBR_HD double stage(const double (&y)[SBR_NX], double (&k)[SBR_NX], double t, const Flow& f, const Coef& c,
                    const TailArgs& a) {
    const double s45 = kinetics(y, k, c, a);
    if (TAIL == TAIL_REACT) return s45;
    const double Vt = f.V(t);
    const double dil = f.q * rcp(Vt);
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i)
        if (active(i)) {
            if (TAIL == TAIL_FILL) k[i] = fma(dil, cin<TAIL>(a, i) - y[i], k[i]);
            else k[i] = fma(-dil, y[i], k[i]);
        }
    if (TAIL == TAIL_EC) k[iSs] = fma(dil, a.ec_conc, k[iSs]);
    // an env that does not dose (q == 0) inside a warp that runs the dosing tail must get the react tail's result
    // bit for bit (its trajectory must not depend on its 31 neighbours): every dilution FMA above is exact for
    // dil == 0, and the quadrature integrand stays s45 as in the react tail (see integrate_interval's epilogue)
    if (TAIL == TAIL_EC && f.q == 0.0) return s45;
    return s45 * Vt;
}

// Full 14-component derivative (the reference's dxdt), for the stage-level entry sbr_rhs and its tests only.
template <int TAIL>
SBR_HD void rhs(const double (&y)[SBR_NX], double (&k)[SBR_NX], const Coef& c, const TailArgs& a) {
    const Flow f{y[iV], TAIL == TAIL_REACT ? 0.0 : a.q};
    const double s45 = kinetics(y, k, c, a);
    const double kin_salk = (k[iSnh] - k[iSno]) * c.c136;
    const double dil = TAIL == TAIL_REACT ? 0.0 : f.q / y[iV];
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i)
        if (active(i) && TAIL != TAIL_REACT) k[i] = fma(dil, cin<TAIL>(a, i) - y[i], k[i]);
    k[iV] = f.q;
    k[iSi] = dil * (cin<TAIL>(a, iSi) - y[iSi]);
    k[iXi] = dil * (cin<TAIL>(a, iXi) - y[iXi]);
    k[iXp] = fma(dil, cin<TAIL>(a, iXp) - y[iXp], c.ixp * s45);
    k[iSalk] = fma(dil, cin<TAIL>(a, iSalk) - y[iSalk], kin_salk);
}

// ---------------------------------------------------------------------------------------------------------
// Classical RK4, one step of size h starting at time offset t.  Low-storage form: x (state), acc (weighted
// sum), y (stage input), k.  xpq accumulates the Xp quadrature sum_i b_i h g_i.
// ---------------------------------------------------------------------------------------------------------
template <int TAIL>
SBR_HD void rk4_step(double (&x)[SBR_NX], double t, double h, const Flow& f, const Coef& c, const TailArgs& a,
                     double& xpq) {
    double y[SBR_NX], k[SBR_NX], acc[SBR_NX];
    const double h2 = 0.5 * h, h6 = h * (1.0 / 6.0), h3 = h * (1.0 / 3.0);
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) y[i] = x[i];
    double g = stage<TAIL>(y, k, t, f, c, a);
    double q = h6 * g;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i)
        if (active(i)) { acc[i] = fma(h6, k[i], x[i]); y[i] = fma(h2, k[i], x[i]); }
    g = stage<TAIL>(y, k, t + h2, f, c, a);
    q = fma(h3, g, q);
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i)
        if (active(i)) { acc[i] = fma(h3, k[i], acc[i]); y[i] = fma(h2, k[i], x[i]); }
    g = stage<TAIL>(y, k, t + h2, f, c, a);
    q = fma(h3, g, q);
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i)
        if (active(i)) { acc[i] = fma(h3, k[i], acc[i]); y[i] = fma(h, k[i], x[i]); }
    g = stage<TAIL>(y, k, t + h, f, c, a);
    xpq += fma(h6, g, q);
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i)
        if (active(i)) x[i] = fma(h6, k[i], acc[i]);
}

// ---------------------------------------------------------------------------------------------------------
// Dormand-Prince 5(4) with FSAL, per-env adaptive step.
// ---------------------------------------------------------------------------------------------------------
// Controller constants (measured on 2^20 random envs): safety 0.85, growth <= 3 and a 0.7
// discount on the first step of every PID interval cut the rejected steps per cycle from 112 to 31 at equal
// kernel time and tighten the 99.9th-percentile error 4x.
#ifndef SBR_DP_LATE_H
#define SBR_DP_LATE_H 0       // 1: multiply the stage sum by h once (fewer live h*a_ij products, one more op per stage)
#endif
#ifndef SBR_DP_SAFETY
#define SBR_DP_SAFETY 0.85f
#endif
#ifndef SBR_DP_MAXGROW
#define SBR_DP_MAXGROW 3.0f
#endif
#ifndef SBR_DP_FIRST
#define SBR_DP_FIRST 0.7
#endif
#ifndef SBR_DP_NGUARD
// Step count of what is left of an interval = ceil(rem / proposal * guard).  A guard below 1 lets the equal steps exceed the
// controller's proposal by up to 1 / guard when that saves a whole step (6 RHS): the predicted error of such a step is
// (1 / guard)^5 x the controller's target of safety^5 = 0.44, i.e. 0.57 of the tolerance at 0.95.  Measured on 2^20 whole cycles
// at rtol 1e-7 (profiles/r02t_*): guard 0.99999 / 0.97 / 0.95 / 0.90 -> 8381 / 8213 / 8119 / 7983 RHS per env, 68.8 / - /
// 66.4 / 65.4 ms, deviation from a converged solution p99.9 0.016 / 0.018 / 0.019 / 0.023 tolerance units, rejects 36.7 / - /
// 42.1 / 57.0 per env; 0.85 starts to trip the step limit.
#define SBR_DP_NGUARD 0.95f
#endif
#ifndef SBR_DP_K7_IN_K1
#define SBR_DP_K7_IN_K1 0
#endif
// Step-size factor safety * en^(-1/10): it only steers the controller, so the device uses the MUFU lg2/ex2 pair
// (2 instructions) instead of the ~30-instruction software log2f.
#ifndef SBR_DP_FLOAT_H
#define SBR_DP_FLOAT_H 1      // 1: the cycle path carries its step-size proposal in FP32 and uses the raw MUFU lg2 / ex2
                              // (measured: 70.55 -> 68.69 ms per 2^20 cycles, identical step counts, profiles/r02p_*)
#endif
SBR_HD float pow_m01(float en) {
#ifdef __CUDA_ARCH__
#if SBR_DP_FLOAT_H
    // en is in [1e-20, inf): no denormal / range fix-ups needed around the two MUFU instructions
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(en));
    l *= -0.1f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return r;
#else
    return exp2f(-0.1f * __log2f(en));
#endif
#else
    return exp2f(-0.1f * log2f(en));
#endif
}
SBR_HD float __frcp_rn_compat(float v) {
#ifdef __CUDA_ARCH__
    return __frcp_rn(v);
#else
    return 1.0f / v;
#endif
}
// 1/v for a step COUNT: the raw MUFU.RCP approximation (1 instruction; __frcp_rn is MUFU + two FFMA + a slow-path
// call).  The 0.99999 factor at the call site absorbs its 1-ulp error.
SBR_HD float frcp_fast(float v) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return 1.0f / v;
#endif
}
struct DpTab {
    double c2, c3, c4, c5, a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65;
    double b1, b3, b4, b5, b6, e1, e3, e4, e5, e6, e7;
};
#define SBR_DP_TABLEAU                                                                                              \
    {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0 / 5, 3.0 / 40, 9.0 / 40, 44.0 / 45, -56.0 / 15, 32.0 / 9,              \
     19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 9017.0 / 3168, -355.0 / 33, 46732.0 / 5247,      \
     49.0 / 176, -5103.0 / 18656, 35.0 / 384, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 71.0 / 57600,    \
     -71.0 / 16695, 71.0 / 1920, -17253.0 / 339200, 22.0 / 525, -1.0 / 40}
#ifdef __CUDACC__
static __constant__ DpTab kDpTab = SBR_DP_TABLEAU;
#endif
struct Dp45State {
    double h;          // current step-size proposal, carried across PID intervals
    uint32_t n_rhs;    // RHS evaluations (accepted + rejected)
    uint32_t n_rej;    // rejected steps
};

// Per-component absolute-tolerance scale: atol_i = atol * scale_i, scale from the reference's own
// normalisation vector x_1_state (gym_SBR_oneshot.py:153) so that So ~ 1e-15 in anoxic phases does not
// drive the step to zero.
SBR_HD constexpr double tol_scale(int i) {
    return i == iV ? 1.32 : i == iSi ? 30.0 : i == iSs ? 30.0 : i == iXi ? 1500.0 : i == iXs ? 150.0
         : i == iXbh ? 3000.0 : i == iXba ? 2000.0 : i == iXp ? 600.0 : i == iSo ? 8.0 : i == iSno ? 20.0
         : i == iSnh ? 20.0 : i == iSnd ? 10.0 : i == iXnd ? 10.0 : 10.0;
}

// Integrate the active components over [0, T] with constant tail arguments; xpq accumulates the Xp quadrature.
// Returns status bits (0 or SBR_ST_STEPLIMIT).  The error norm is the RMS over the 9 active components.
template <int TAIL>
SBR_HD int dp45_interval(double (&x)[SBR_NX], double T, const Flow& f, const Coef& c, const TailArgs& a,
                         const SbrTol& tol, Dp45State& st, double& xpq) {
    // Butcher tableau (Dormand & Prince 1980): on the device the coefficients are operands straight from the
    // constant bank (kDpTab) -- as literals the compiler rebuilds 24 of them with two UMOVs each in every step
#ifdef __CUDA_ARCH__
    const DpTab& tb = kDpTab;
#else
    const DpTab tb = SBR_DP_TABLEAU;
#endif
    const double c2 = tb.c2, c3 = tb.c3, c4 = tb.c4, c5 = tb.c5;
    const double a21 = tb.a21;
    const double a31 = tb.a31, a32 = tb.a32;
    const double a41 = tb.a41, a42 = tb.a42, a43 = tb.a43;
    const double a51 = tb.a51, a52 = tb.a52, a53 = tb.a53, a54 = tb.a54;
    const double a61 = tb.a61, a62 = tb.a62, a63 = tb.a63, a64 = tb.a64, a65 = tb.a65;
    const double b1 = tb.b1, b3 = tb.b3, b4 = tb.b4, b5 = tb.b5, b6 = tb.b6;
    const double e1 = tb.e1, e3 = tb.e3, e4 = tb.e4, e5 = tb.e5, e6 = tb.e6, e7 = tb.e7;
    double k1[SBR_NX], k2[SBR_NX], k3[SBR_NX], k4[SBR_NX], k5[SBR_NX], k6[SBR_NX], y[SBR_NX];
    double t = 0.0;
    double h = st.h * SBR_DP_FIRST;   // KLa has just jumped: the carried proposal is discounted for the first step
    int status = 0;
    int steps = 0;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) y[i] = x[i];
    double g1 = stage<TAIL>(y, k1, 0.0, f, c, a);
    st.n_rhs += 1;
    while (t < T) {
        if (steps >= tol.max_steps) { status = SBR_ST_STEPLIMIT; break; }
        ++steps;
        // spread what is left of the interval over equal steps no longer than the controller's proposal: a
        // proposal that does not divide the interval would otherwise end it with a sliver step (6 RHS for nothing)
        const double rem = T - t;
        const float n_f = ceilf((float)rem * frcp_fast((float)h) * SBR_DP_NGUARD);   // float is plenty for a count
        const bool last = !(n_f > 1.0f);
        const double hs = last ? rem : rem * (double)frcp_fast(n_f);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) y[i] = fma(hs * a21, k1[i], x[i]);
        stage<TAIL>(y, k2, fma(c2, hs, t), f, c, a);   // b2 = 0: no quadrature contribution
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) y[i] = SBR_DP_LATE_H ? fma(hs, fma(a32, k2[i], a31 * k1[i]), x[i])
                                                : fma(hs * a32, k2[i], fma(hs * a31, k1[i], x[i]));
        const double g3 = stage<TAIL>(y, k3, fma(c3, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) y[i] = SBR_DP_LATE_H ? fma(hs, fma(a43, k3[i], fma(a42, k2[i], a41 * k1[i])), x[i])
                                                : fma(hs * a43, k3[i], fma(hs * a42, k2[i], fma(hs * a41, k1[i], x[i])));
        const double g4 = stage<TAIL>(y, k4, fma(c4, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i))
                y[i] = SBR_DP_LATE_H ? fma(hs, fma(a54, k4[i], fma(a53, k3[i], fma(a52, k2[i], a51 * k1[i]))), x[i])
                     : fma(hs * a54, k4[i], fma(hs * a53, k3[i], fma(hs * a52, k2[i], fma(hs * a51, k1[i], x[i]))));
        const double g5 = stage<TAIL>(y, k5, fma(c5, hs, t), f, c, a);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i))
                y[i] = SBR_DP_LATE_H ? fma(hs, fma(a65, k5[i], fma(a64, k4[i], fma(a63, k3[i], fma(a62, k2[i], a61 * k1[i])))), x[i])
                     : fma(hs * a65, k5[i], fma(hs * a64, k4[i], fma(hs * a63, k3[i],
                       fma(hs * a62, k2[i], fma(hs * a61, k1[i], x[i])))));
        const double g6 = stage<TAIL>(y, k6, t + hs, f, c, a);
        // 5th-order solution (into y) ; k2 is free from here on and receives k7 = f(y)
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i))
                y[i] = SBR_DP_LATE_H ? fma(hs, fma(b6, k6[i], fma(b5, k5[i], fma(b4, k4[i], fma(b3, k3[i], b1 * k1[i])))), x[i])
                     : fma(hs * b6, k6[i], fma(hs * b5, k5[i], fma(hs * b4, k4[i],
                       fma(hs * b3, k3[i], fma(hs * b1, k1[i], x[i])))));
        const double g7 = stage<TAIL>(y, k2, t + hs, f, c, a);
        st.n_rhs += 6;
        // error estimate, RMS norm over the active components
        // (h is factored out of the 9 error components; the per-component scale only steers the controller: it is
        // taken from the new solution alone -- |y| is an operand modifier, where max(|x|, |y|) costs ten integer
        // instructions per component -- and its reciprocal is the raw MUFU approximation; three partial sums instead
        // of one serial DFMA chain)
        double en3[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i)
            if (active(i)) {
                const double err = fma(e7, k2[i], fma(e6, k6[i], fma(e5, k5[i], fma(e4, k4[i],
                                   fma(e3, k3[i], e1 * k1[i])))));
                const double sc = fma(tol.rtol, fabs(y[i]), tol.atol * tol_scale(i));
                const double q = err * rcp_rough(sc);
                en3[aidx(i) % 3] = fma(q, q, en3[aidx(i) % 3]);
            }
        const double en = ((en3[0] + en3[1]) + en3[2]) * (hs * hs * (1.0 / 9));   // mean square
        const bool finite = en < 1e300;   // false for NaN/Inf
        if (en <= 1.0 || !finite) {
            // accept (a non-finite state is accepted so that the loop terminates; flagged by the caller)
            t = last ? T : t + hs;
            xpq = fma(hs, fma(b6, g6, fma(b5, g5, fma(b4, g4, fma(b3, g3, b1 * g1)))), xpq);
            g1 = g7;
#pragma unroll
            for (int i = 0; i < SBR_NX; ++i)
                if (active(i)) { x[i] = y[i]; k1[i] = k2[i]; }
            if (!finite) { t = T; }
        }
        else {
            st.n_rej += 1;
        }
        // step-size controller: h *= clamp(safety * en^(-1/10), 0.2, max growth)  (en is the SQUARED norm).  A
        // rejected step costs the whole warp 6 RHS evaluations (the other 31 envs wait), so the constants lean
        // conservative: see DESIGN.md section 4 for the measured trade-off.
        float fac = SBR_DP_MAXGROW;
        if (en > 1e-20) {
            fac = SBR_DP_SAFETY * pow_m01((float)en);
            fac = fminf(SBR_DP_MAXGROW, fmaxf(0.2f, fac));
        }
        if (en > 1.0) fac = fminf(fac, 1.0f);
        if (!(last && en <= 1.0)) h = hs * (double)fac;   // a truncated final step does not shrink the carry
        else h = fmax(h, hs * (double)fac);
    }
    st.h = h;
    return status;
}

// ---------------------------------------------------------------------------------------------------------
// One PID interval of length T: RK4 with n_sub equal sub-steps, or DP45, on the active components; then the
// closed forms of the passive ones (see the table above).
// ---------------------------------------------------------------------------------------------------------
template <int TAIL, int MODE>
SBR_HD int integrate_interval(double (&x)[SBR_NX], double T, int n_sub, const Coef& c,
                              const TailArgs& a_in, const SbrTol& tol, Dp45State& st) {
    TailArgs a = a_in;
    a.kla_sat = a.kla * c.so_sat;
    const Flow f{x[iV], TAIL == TAIL_REACT ? 0.0 : a.q};
    const double snh0 = x[iSnh], sno0 = x[iSno];
    double xpq = 0.0;
    int status = 0;
    if (MODE == SBR_MODE_RK4) {
        const double h = T / (double)n_sub;
        for (int s = 0; s < n_sub; ++s) rk4_step<TAIL>(x, (double)s * h, h, f, c, a, xpq);
        st.n_rhs += 4u * (uint32_t)n_sub;
    } else {
        status = dp45_interval<TAIL>(x, T, f, c, a, tol, st, xpq);
    }
    const double dN = ((x[iSnh] - snh0) - (x[iSno] - sno0)) * c.c136;
    if (TAIL == TAIL_REACT || (TAIL == TAIL_EC && f.q == 0.0)) {
        x[iXp] = fma(c.ixp, xpq, x[iXp]);
        x[iSalk] += dN;
    } else {
        const double V1 = f.V(T), qT = f.q * T;
        const double w = f.V0 * rcp(V1), u = qT * rcp(V1);       // x(T) = x0 V0/V1 + c_in q T / V1
        const double D0 = x[iSalk];                               // D = Salk - (Snh - Sno)/14, shifted by a constant
        x[iV] = V1;
        x[iSi] = fma(x[iSi], w, cin<TAIL>(a, iSi) * u);
        x[iXi] = fma(x[iXi], w, cin<TAIL>(a, iXi) * u);
        x[iXp] = fma(x[iXp], w, fma(cin<TAIL>(a, iXp), u, c.ixp * xpq * rcp(V1)));
        // D'(t) = (q/V)(D_in - D) with D = Salk - (Snh - Sno)/14  =>  closed form for D, then add the nitrogen part back
        const double n0 = (snh0 - sno0) * c.c136;
        const double n_in = (cin<TAIL>(a, iSnh) - cin<TAIL>(a, iSno)) * c.c136;
        const double D1 = fma(D0 - n0, w, (cin<TAIL>(a, iSalk) - n_in) * u);
        x[iSalk] = D1 + (x[iSnh] - x[iSno]) * c.c136;
        (void)dN;
    }
    return status;
}

// ---------------------------------------------------------------------------------------------------------
// DO -> KLa positional PID of the cycle-per-step path (sub_phases_FB.py:233-250).
// ---------------------------------------------------------------------------------------------------------
struct PidA {
    double Kc, Kc_tauI, Kc_tauD, dt, inv_dt, lo, hi;
    bool bypass;       // SBR_FLAG_RAW_KLA: the "set-point" IS the phase's KLa, no controller
};

// Clipping as the reference does it (np.clip / min(max(a, lo), hi) / if-elif chains): a NaN action is NOT
// sanitised, it propagates into the state and ends up in status[i] (SBR_ST_NONFINITE), exactly as the reference's
// outputs turn NaN; CUDA's fmin/fmax would silently turn it into a bound.
SBR_HD double clip_keep_nan(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

SBR_HD PidA make_pid_a(const SbrParams& p, const Coef& c) {
    PidA q;
    q.Kc = p.pid_Kc; q.Kc_tauI = c.pidA_KcI; q.Kc_tauD = c.pidA_KcD;
    q.dt = p.pid_dt; q.inv_dt = c.pidA_inv_dt; q.lo = p.kla_min; q.hi = p.kla_max;
    q.bypass = false;
    return q;
}

// One update of the positional DO->KLa PID (sub_phases_FB.py:233-250).  `first` = interval 0 of a phase: no
// derivative / integral update, and the clamped output becomes the bias of the phase's later intervals (:218,243).
// Two independent clamp checks, each undoing the integral update (:245-250).  The derivative uses the gain folded
// on the host (1/dt, at most 1 ulp from the reference's division), like the interval-per-step PIDs.
SBR_HD double pid_a_update(const PidA& pid, double sp, double so_i, double so_prev, bool first, double& ie,
                           double& bias) {
    if (pid.bypass) return sp;                   // raw-KLa action: constant over the phase
    const double e = sp - so_i;
    double dcv = 0.0;
    if (!first) {
        dcv = (so_i - so_prev) * pid.inv_dt;
        ie = ie + e * pid.dt;
    }
    double kla = pid.Kc * e + pid.Kc_tauI * ie + pid.Kc_tauD * dcv + bias;
    if (kla > pid.hi) { kla = pid.hi; ie = ie - e * pid.dt; }
    if (kla < pid.lo) { kla = pid.lo; ie = ie - e * pid.dt; }
    if (first) bias = kla;
    return kla;
}

struct PhaseOut {
    double kla_sum;
    double kla_last;
};

// One PID-controlled phase = filling.sim_rxn / rxn.sim_rxn (sub_phases_FB.py:178-271, 406-500), interval by
// interval: the fixed-step (RK4) form of the cycle path.  So is sampled at interval starts.
template <int TAIL, int MODE>
SBR_HD int pid_phase(double (&x)[SBR_NX], int n_int, int n_sub, double T, double sp, double kla_in,
                     const Coef& c, TailArgs a, const PidA& pid, const SbrTol& tol, Dp45State& st,
                     PhaseOut& out) {
    double bias = kla_in, ie = 0.0, so_prev = 0.0, so_i = x[iSo];
    double ksum = 0.0, kla = kla_in;
    int status = 0;
    for (int i = 0; i < n_int; ++i) {
        kla = pid_a_update(pid, sp, so_i, so_prev, i == 0, ie, bias);
        a.kla = kla;
        status |= integrate_interval<TAIL, MODE>(x, T, n_sub, c, a, tol, st);
        ksum += kla;
        so_prev = so_i;
        so_i = x[iSo];
    }
    out.kla_sum = ksum;
    out.kla_last = kla;
    return status;
}

// ---------------------------------------------------------------------------------------------------------
// Adaptive (Dormand-Prince) form of the cycle path: one SEGMENT = a run of consecutive PID-controlled phases that
// share a tail (fill | react phases 2-5 | idle), integrated by one loop nest "for interval: while (t < T) step".
//
// First-same-as-last is kept across PID intervals and phase switches of a segment: the last stage of an interval
// is f(x_end); the next interval starts from the same state with only KLa changed by the PID, and KLa enters the
// right-hand side linearly and only in d(So)/dt -- so the carried stage is corrected by dKLa (So_sat - So).  dKLa
// is formed from the PID output and the KLa still held in `kla` BEFORE `kla` is overwritten: round 1 kept a
// second loop-carried copy (Fsal::kla) and ptxas 12.9 coalesced that copy with the new value ("lost copy":
// `DADD R12, R74, -R74` -- the correction compiled to zero, profiles/r02_fsal_lost_copy_sass.txt).
//
// Register budget (this is what bounds occupancy): the classical formulation keeps x, y and six stage vectors
// live (72 doubles).  Here the 5th-order solution and the error estimate are accumulated as soon as their inputs
// exist, so that k2..k5 die when the input of stage 6 is formed; the passive components (V, Si, Xi, Xp, Salk) leave
// the loop entirely: closed forms / one quadrature over the whole segment (see the table above `active`).  The step
// loop of the react segment then runs in 254 registers without spills (8 warps/SM).
//
// Tried and rejected, each measured on 2^20 envs at rtol 1e-7, envs sorted by first set-point (profiles/r02*_ab_*):
//  * k3 and k4 parked in a shared-memory column between their uses (45 LDS/STS per step; 168 registers, 12 warps/SM,
//    no spills): 105.6 ms against 76.6 ms -- beside the FP64 pipe every LSU instruction costs far more than its
//    issue slot, and the third warp per sub-partition does not buy it back (compiler spills at 168 registers: 86.1 ms).
//  * a single flat loop whose body is "one step attempt" with the interval / phase bookkeeping as a predicated block
//    inside it, so that the lanes of a warp run through their intervals at their own pace and only re-converge at the
//    end of the segment.  The per-env step count is a property of the env (its first DO set-point: corr -0.96 below
//    1 g/m3), not of the interval, so max_lane(sum of steps) is no better than sum_interval(max_lane steps) in env
//    order (13.8k vs 14.3k RHS per warp against a mean of 8.4k), and with envs sorted by set-point the 3 % it gains
//    (30.6 instead of 29.6 of 32 lanes) is eaten by the bookkeeping block running in every iteration.
//  * the step-size controller run once per planned run of equal steps instead of once per step (it is ~300 cycles of
//    dependent latency per step: the same body without any controller takes 1300 cycles per warp-step, with it 1860):
//    the graded steps after each KLa jump are lost and the RHS count rises 37 %.
// Kept: the error norm as three partial sums and the raw MUFU reciprocal for the step count (-4 %).
// ---------------------------------------------------------------------------------------------------------


// Scratch column of one env outside the register file: slot j at p[j * stride] (shared memory on the device,
// conflict-free for consecutive threads; a local array in the CPU twin).  Holds what is written once per phase and
// read once per cycle (the per-phase KLa sums), so that it costs no registers inside the step loop.
enum { PARK_KSUM = 0, PARK_SLOTS = 4 };
struct Park {
    double* p;
    int stride;
    SBR_HD void put(int j, double v) const { p[j * stride] = v; }
    SBR_HD double get(int j) const { return p[j * stride]; }
};

// ph0: schedule index of the segment's first phase; NPH phases follow each other in the schedule; sp[]: this env's
// DO set-point per phase.  On return: x = state at the end of the segment (all 14 components), kla_last = KLa of the
// last interval, park slots PARK_KSUM + j = sum of the per-interval KLa of phase j.  Returns status bits.
template <int TAIL, int NPH>
SBR_HD int dp45_segment(double (&x)[SBR_NX], const SbrSchedule& s, int ph0, const double (&sp)[NPH], double kla_in,
                        const Coef& c, TailArgs a, const PidA& pid, const SbrTol& tol, Dp45State& st,
                        const Park& park, double& kla_last) {
#ifdef __CUDA_ARCH__
    const DpTab& tb = kDpTab;
#else
    const DpTab tb = SBR_DP_TABLEAU;
#endif
    const double V_start = x[iV], snh0 = x[iSnh], sno0 = x[iSno];
    Flow f{V_start, TAIL == TAIL_REACT ? 0.0 : a.q};      // f.V0 = volume at the start of the current interval
    double xpq = 0.0;                                      // Xp quadrature over the whole segment
    // controller state of the current phase
    int ph = 0, i_int = 0, n_int = s.n_int[ph0];
    double T = s.interval[ph0], spc = sp[0];
    double bias = kla_in, ie = 0.0, so_prev = 0.0, so_i = x[iSo], ksum = 0.0;
    double kla = pid_a_update(pid, spc, so_i, so_prev, true, ie, bias);
    a.kla = kla;
    a.kla_sat = kla * c.so_sat;
    double k1[SBR_NX], k2[SBR_NX], k3[SBR_NX], k4[SBR_NX], k5[SBR_NX], y[SBR_NX], sol[SBR_NX], err[SBR_NX];
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) y[i] = x[i];
    double g1 = stage<TAIL>(y, k1, 0.0, f, c, a);
    st.n_rhs += 1;
#if SBR_DP_FLOAT_H
    // the proposal only steers the step count: carried in FP32, so that the serial tail of a step attempt (error norm ->
    // factor -> proposal -> count -> step) holds one conversion instead of three and no FP64 multiply
    float hf = (float)st.h;
#else
    double h = st.h;
#endif
    int status = 0;
    for (;;) {
        // ---- one PID interval: step attempts until t reaches T ----
        double t = 0.0;
        int steps = 0;
#if SBR_DP_FLOAT_H
        hf = hf * (float)SBR_DP_FIRST;
#else
        h = h * SBR_DP_FIRST;          // KLa has just jumped: the carried proposal is discounted for the first step
#endif
        while (t < T) {
            // work bound of an env that has left the physical regime: give the interval up (state flagged)
            if (steps >= tol.max_steps) { status |= SBR_ST_STEPLIMIT; break; }
            ++steps;
            // spread what is left of the interval over equal steps no longer than the controller's proposal: a
            // proposal that does not divide the interval would otherwise end it with a sliver step (6 RHS for nothing)
            const double rem = T - t;
#if SBR_DP_FLOAT_H
            const float rem_f = (float)rem;
            const float n_f = ceilf(rem_f * frcp_fast(hf) * SBR_DP_NGUARD);
            const bool last = !(n_f > 1.0f);
            const float inv_n = frcp_fast(n_f);
            const double hs = last ? rem : rem * (double)inv_n;
            const float hs_f = last ? rem_f : rem_f * inv_n;
#else
            const float n_f = ceilf((float)rem * frcp_fast((float)h) * SBR_DP_NGUARD);   // float is plenty for a count
            const bool last = !(n_f > 1.0f);
            const double hs = last ? rem : rem * (double)frcp_fast(n_f);
#endif
            {
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i)) y[i] = fma(hs * tb.a21, k1[i], x[i]);
                stage<TAIL>(y, k2, fma(tb.c2, hs, t), f, c, a);   // b2 = 0: no quadrature contribution
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i)) y[i] = fma(hs * tb.a32, k2[i], fma(hs * tb.a31, k1[i], x[i]));
                const double g3 = stage<TAIL>(y, k3, fma(tb.c3, hs, t), f, c, a);
                double gq = fma(tb.b3, g3, tb.b1 * g1);
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i)) y[i] = fma(hs * tb.a43, k3[i], fma(hs * tb.a42, k2[i], fma(hs * tb.a41, k1[i], x[i])));
                const double g4 = stage<TAIL>(y, k4, fma(tb.c4, hs, t), f, c, a);
                gq = fma(tb.b4, g4, gq);
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i))
                        y[i] = fma(hs * tb.a54, k4[i], fma(hs * tb.a53, k3[i], fma(hs * tb.a52, k2[i],
                               fma(hs * tb.a51, k1[i], x[i]))));
                const double g5 = stage<TAIL>(y, k5, fma(tb.c5, hs, t), f, c, a);
                gq = fma(tb.b5, g5, gq);
                // input of stage 6, and everything of the 5th-order solution and of the error estimate that k1..k5
                // contribute: k2..k5 are dead after this block
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i)) {
                        y[i] = fma(hs * tb.a65, k5[i], fma(hs * tb.a64, k4[i], fma(hs * tb.a63, k3[i],
                               fma(hs * tb.a62, k2[i], fma(hs * tb.a61, k1[i], x[i])))));
                        sol[i] = fma(hs * tb.b5, k5[i], fma(hs * tb.b4, k4[i], fma(hs * tb.b3, k3[i],
                                 fma(hs * tb.b1, k1[i], x[i]))));
                        err[i] = fma(tb.e5, k5[i], fma(tb.e4, k4[i], fma(tb.e3, k3[i], tb.e1 * k1[i])));
                    }
                const double g6 = stage<TAIL>(y, k2, t + hs, f, c, a);       // k6 lands in k2's registers
                gq = fma(tb.b6, g6, gq);
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i)) {
                        sol[i] = fma(hs * tb.b6, k2[i], sol[i]);
                        err[i] = fma(tb.e6, k2[i], err[i]);
                    }
#if SBR_DP_K7_IN_K1
                // k1..k5 are dead (their shares of sol / err were taken above), so k7 = f(5th-order solution) lands in
                // k1's registers: an accepted step then needs no k1 <- k7 copy (18 moves); a rejected one (2-3 % of the
                // attempts) re-evaluates k1 = f(x) instead
                double (&k7)[SBR_NX] = k1;
#else
                double (&k7)[SBR_NX] = k2;
#endif
                const double g7 = stage<TAIL>(sol, k7, t + hs, f, c, a);     // k7 = f(5th-order solution)
                st.n_rhs += 6;
                // error estimate, RMS norm over the active components (h is factored out of the 9 components; the
                // per-component scale only steers the controller, so its reciprocal is the raw MUFU approximation).
                // Three partial sums: the 9 squares would otherwise be one serial DFMA chain at the very point
                // where the step has no other work left to overlap with it.
                double en3[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int i = 0; i < SBR_NX; ++i)
                    if (active(i)) {
                        const double e_i = fma(tb.e7, k7[i], err[i]);
                        const double sc = fma(tol.rtol, fabs(sol[i]), tol.atol * tol_scale(i));
                        const double q = e_i * rcp_rough(sc);
                        en3[aidx(i) % 3] = fma(q, q, en3[aidx(i) % 3]);
                    }
                const double en = ((en3[0] + en3[1]) + en3[2]) * (hs * hs * (1.0 / 9));   // mean square
                const bool finite = en < 1e300;   // false for NaN/Inf
                if (en <= 1.0 || !finite) {
                    // accept (a non-finite state is accepted so that the loop terminates; flagged by the caller)
                    t = (last || !finite) ? T : t + hs;
                    xpq = fma(hs, gq, xpq);
                    g1 = g7;
#pragma unroll
                    for (int i = 0; i < SBR_NX; ++i)
                        if (active(i)) { x[i] = sol[i]; if (!SBR_DP_K7_IN_K1) k1[i] = k2[i]; }
                } else {
                    st.n_rej += 1;
#if SBR_DP_K7_IN_K1
#pragma unroll
                    for (int i = 0; i < SBR_NX; ++i) y[i] = x[i];
                    stage<TAIL>(y, k1, t, f, c, a);
                    st.n_rhs += 1;
#endif
                }
                // step-size controller: h *= clamp(safety * en^(-1/10), 0.2, max growth)  (en is the SQUARED norm).
                // A rejected step costs the whole warp 6 RHS evaluations (the other 31 envs wait), so the constants
                // lean conservative: see DESIGN.md section 4 for the measured trade-off.
#if SBR_DP_FLOAT_H
                // branch-free: an error below 1e-20 (or NaN) maps to the floor, whose factor clamps to the maximum growth
                float fac = SBR_DP_SAFETY * pow_m01(fmaxf((float)en, 1e-20f));
                fac = fminf(SBR_DP_MAXGROW, fmaxf(0.2f, fac));
#else
                float fac = SBR_DP_MAXGROW;
                if (en > 1e-20) {
                    fac = SBR_DP_SAFETY * pow_m01((float)en);
                    fac = fminf(SBR_DP_MAXGROW, fmaxf(0.2f, fac));
                }
#endif
                if (en > 1.0) fac = fminf(fac, 1.0f);
#if SBR_DP_FLOAT_H
                if (!(last && en <= 1.0)) hf = hs_f * fac;        // a truncated final step does not shrink the carry
                else hf = fmaxf(hf, hs_f * fac);
#else
                if (!(last && en <= 1.0)) h = hs * (double)fac;   // a truncated final step does not shrink the carry
                else h = fmax(h, hs * (double)fac);
#endif
            }
        }
        // ---- end of a PID interval (sub_phases_FB.py:226-265): sample So, next PID output, KLa jump ----
        ksum += kla;
        so_prev = so_i;
        so_i = x[iSo];
        if (TAIL != TAIL_REACT) f.V0 = f.V(T);
        ++i_int;
        bool first = false;
        if (i_int == n_int) {
            // ---- end of a phase (SBR_model_FB.py:88-172): the next one starts from this phase's last KLa ----
            park.put(PARK_KSUM + ph, ksum);
            ksum = 0.0;
            ++ph;
            if (ph == NPH) break;
            i_int = 0;
            n_int = s.n_int[ph0 + ph];
            T = s.interval[ph0 + ph];
#pragma unroll
            for (int j = 1; j < NPH; ++j)
                if (ph == j) spc = sp[j];
            ie = 0.0; bias = kla; first = true;
        }
        const double kla_new = pid_a_update(pid, spc, so_i, so_prev, first, ie, bias);
        // first-same-as-last across the KLa jump: dKLa from the OLD kla, then overwrite it (see the header)
        k1[iSo] = fma(kla_new - kla, c.so_sat - x[iSo], k1[iSo]);
        kla = kla_new;
        a.kla = kla_new;
        a.kla_sat = kla_new * c.so_sat;
    }
#if SBR_DP_FLOAT_H
    st.h = (double)hf;
#else
    st.h = h;
#endif
    kla_last = kla;
    // ---- passive components over the whole segment (closed forms; see the table above `active`) ----
    if (TAIL == TAIL_REACT) {
        x[iXp] = fma(c.ixp, xpq, x[iXp]);
        x[iSalk] += ((x[iSnh] - snh0) - (x[iSno] - sno0)) * c.c136;
    } else {
        const double V_end = f.V0;
        const double w = V_start / V_end, u = (V_end - V_start) / V_end;   // x_end = x_start V_start/V_end + c_in dV/V_end
        x[iV] = V_end;
        x[iSi] = fma(x[iSi], w, cin<TAIL>(a, iSi) * u);
        x[iXi] = fma(x[iXi], w, cin<TAIL>(a, iXi) * u);
        x[iXp] = fma(x[iXp], w, fma(cin<TAIL>(a, iXp), u, c.ixp * xpq / V_end));
        // D = Salk - (Snh - Sno)/14 is a pure dilution variable (charge balance): closed form, nitrogen part added back
        const double n0 = (snh0 - sno0) * c.c136;
        const double n_in = (cin<TAIL>(a, iSnh) - cin<TAIL>(a, iSno)) * c.c136;
        const double D1 = fma(x[iSalk] - n0, w, (cin<TAIL>(a, iSalk) - n_in) * u);
        x[iSalk] = D1 + (x[iSnh] - x[iSno]) * c.c136;
    }
    return status;
}

// ---------------------------------------------------------------------------------------------------------
// Settler: the reference's 10-layer flux model takes max(vmax, exp-exp) (sub_phases_FB.py:677-686), and
// exp(-rh d) - exp(-rp d) < 1 for every real d, so v == vmax in every layer and the ODE is the linear chain
//   s0' = a s1 ; s_i' = a (s_{i+1} - s_i) (1<=i<=8) ; s9' = -a s9 ,  a = vmax * As / V, equal initial layers Xf.
// Closed form: s_{9-k}(T) = Xf e^{-aT} sum_{j<=k} (aT)^j / j!  (k = 0..8), s0 = 10 Xf - sum_{i>=1} s_i
// (matches the reference's odeint to 6e-8).  Layer 0 = bottom.
// ---------------------------------------------------------------------------------------------------------
SBR_HD void settle_closed_form(const double (&x)[SBR_NX], double T, double area, double vmax, double (&sX)[10],
                               double& Xf) {
    Xf = 0.75 * (x[iXi] + x[iXs] + x[iXbh] + x[iXba] + x[iXp]);   // sub_phases_FB.py:730
    const double aT = vmax * area / x[iV] * T;
    const double ex = exp(-aT) * Xf;
    double term = 1.0, part = 1.0, rest = 0.0;
    sX[9] = ex;
    rest = sX[9];
#pragma unroll
    for (int kk = 1; kk <= 8; ++kk) {
        term = term * aT / (double)kk;
        part += term;
        sX[9 - kk] = ex * part;
        rest += sX[9 - kk];
    }
    sX[0] = 10.0 * Xf - rest;
}

// drawing.cal_eq (sub_phases_FB.py:868-915) == the EQI block of module_reward_EQIOCI.sbr_reward
// (module_reward_EQIOCI.py:28-47).  eff = [0.66, Ntot, COD, Snh, BOD5, Sno].
SBR_HD double effluent_quality(const double (&xe)[SBR_NX], double (&eff)[6]) {
    const double Snkj = xe[iSnh] + xe[iSnd] + xe[iXnd] + 0.08 * (xe[iXbh] + xe[iXba]) + 0.06 * (xe[iXp] + xe[iXi]);
    const double Ntot = xe[iSno] + Snkj;
    const double SS = 0.75 * (xe[iXs] + xe[iXi] + xe[iXbh] + xe[iXba] + xe[iXp]);
    const double BOD5 = 0.25 * (xe[iSs] + xe[iXs] + (1 - 0.08) * (xe[iXbh] + xe[iXba]));
    const double COD = xe[iSs] + xe[iSi] + xe[iXs] + xe[iXi] + xe[iXbh] + xe[iXba] + xe[iXp];
    eff[0] = 0.66; eff[1] = Ntot; eff[2] = COD; eff[3] = xe[iSnh]; eff[4] = BOD5; eff[5] = xe[iSno];
    return (2 * SS + 1 * COD + 30 * Snkj + 10 * xe[iSno] + 2 * BOD5) * (1.0 / 1000) * 0.66;
}

struct DrawOut {
    double Qw, EQI;
    double eff[6];   // [0.66, Ntot, COD, Snh, BOD5, Sno]
    int status;
};

// drawing.sim_drawing + cal_eq (sub_phases_FB.py:780-915).  x is replaced by the post-draw reactor state.
SBR_HD void draw_and_waste(double (&x)[SBR_NX], const double (&sX)[10], double Xf, double Qeff, double biomass_sp,
                           DrawOut& o) {
    const double V0 = x[iV];
    const double lv = V0 / 10;
    double resV = V0 - Qeff;
    // m = int(ceil(round(Qeff / lv)))  -- Python round = half-to-even = rint
    const double mr = rint(Qeff / lv);
    int m = (int)mr;
    o.status = 0;
    if (!(mr >= 1.0 && mr <= 9.0)) { o.status |= SBR_ST_LAYERS; m = m < 1 ? 1 : 9; }
    // effluent solids: layers [10-m, 8]; the top layer is left out by the reference's [-m:-1] slice (:794)
    double sX_eff = 0.0;
#pragma unroll
    for (int i = 1; i < 9; ++i)
        if (i >= 10 - m) sX_eff += sX[i] * lv;
    double xe[SBR_NX];
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) xe[i] = x[i];
    xe[iXs] = x[iXs] * (1 / 0.75) * sX_eff / Xf;
    xe[iXp] = x[iXp] * (1 / 0.75) * sX_eff / Xf;
    xe[iXi] = x[iXi] * (1 / 0.75) * sX_eff / Xf;
    xe[iXbh] = x[iXbh] * (1 / 0.75) * sX_eff / Xf;
    xe[iXba] = x[iXba] * (1 / 0.75) * sX_eff / Xf;
    // waste sludge bottom-up (:805-836)
    double wl[10];
    double total = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        wl[i] = (i < 10 - m) ? lv * sX[i] : 0.0;
        if (i < 10 - m) total += wl[i];
    }
    double waste = total - biomass_sp * resV;
    double Qw = NAN;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (i < 10 - m && !found) {
            const double left = waste - wl[i];
            if (left > 0) {
                waste = left;
                wl[i] = 0.0;
                resV -= lv;
            } else {
                Qw = waste / (sX[i] - biomass_sp);
                wl[i] = wl[i] - Qw * sX[i];
                resV -= Qw;
                found = true;
            }
        }
    }
    if (!found) o.status |= SBR_ST_WASTE;
    double rem = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i)
        if (i < 10 - m) rem += wl[i];
    const double sX2 = rem / resV;
    x[iV] = resV;
    x[iXs] = x[iXs] * (1 / 0.75) * sX2 / Xf;
    x[iXp] = x[iXp] * (1 / 0.75) * sX2 / Xf;
    x[iXi] = x[iXi] * (1 / 0.75) * sX2 / Xf;
    x[iXbh] = x[iXbh] * (1 / 0.75) * sX2 / Xf;
    x[iXba] = x[iXba] * (1 / 0.75) * sX2 / Xf;
    // cal_eq (:868-915) on the effluent
    o.EQI = effluent_quality(xe, o.eff);
    o.Qw = Qw;
}

// module_reward.sbr_reward (module_reward.py:4-51): means of the per-interval KLa of phases 3, 5, 8.
SBR_HD void reward_v2(const SbrParams& p, double kla3_mean, double kla5_mean, double kla8_mean, double Qw, double Snh,
                      double& reward, double& OCI) {
    const double ME = 0.005 * 1.32 * 24 + 0.005 * 1.32 * 24;
    const double AE = p.so_sat / (1.8 * 1000) * (1.32 * kla3_mean + 1.32 * kla5_mean + (1.32 - Qw) * kla8_mean);
    const double PE = 0.004 * p.Qin + 0.05 * Qw + 0.004 * p.Qeff;
    OCI = AE + PE + ME;
    reward = (5 - OCI) + (Snh < 4 ? 0.0 : -20.0);
}

struct CycleOut {
    double obs[3];
    double reward;
    double aux[SBR_AUX_ROWS];
    int status;
};

// Epilogue shared by both integrator forms: reward (module_reward.py:4-51), observation and aux rows
// (gym_SBR_env2.py:156-171).
SBR_HD void cycle_epilogue(const double (&x)[SBR_NX], const SbrParams& p, const DrawOut& d, const double (&kla_mean)[3],
                           int status, CycleOut& o) {
    double reward, OCI;
    reward_v2(p, kla_mean[0], kla_mean[1], kla_mean[2], d.Qw, d.eff[3], reward, OCI);
    bool finite = true;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) finite = finite && (fabs(x[i]) < 1e300);
    if (!finite || !(fabs(reward) < 1e300)) status |= SBR_ST_NONFINITE;
    o.obs[0] = p.Qeff; o.obs[1] = d.eff[2]; o.obs[2] = d.eff[3] / 30;
    o.reward = reward;
    o.aux[SBR_AUX_OCI] = OCI; o.aux[SBR_AUX_QW] = d.Qw; o.aux[SBR_AUX_EQI] = d.EQI;
#pragma unroll
    for (int j = 0; j < 6; ++j) o.aux[SBR_AUX_EFF_Q + j] = d.eff[j];
    o.aux[SBR_AUX_KLA3_MEAN] = kla_mean[0]; o.aux[SBR_AUX_KLA5_MEAN] = kla_mean[1];
    o.aux[SBR_AUX_KLA8_MEAN] = kla_mean[2];
    o.status = status;
}

// Whole cycle = SBR_model_FB.run (SBR_model_FB.py:8-295) + SbrEnv2.step epilogue (gym_SBR_env2.py:131-171).
// x: in = start state, out = state after the idle phase.  action: raw, clipped here.
// MODE RK4: interval by interval (pid_phase); MODE DP45: three segments (dp45_segment), `park` required.
template <int MODE>
SBR_HD void cycle_v2(double (&x)[SBR_NX], const double (&action)[3], Loading load, double q_fill,
                     const SbrParams& p, const Coef& c, const SbrSchedule& s, const SbrTol& tol,
                     Dp45State& st, CycleOut& o, const Park& park) {
    PidA pid = make_pid_a(p, c);
    // SBR_FLAG_RAW_KLA (BASELINE configs[1] "random KLa actions"): the three actions are the KLa of phases 3, 5 and 8
    // as fractions of kla_max, the other reacting phases run unaerated, and no DO controller is in the loop
    pid.bypass = (tol.flags & SBR_FLAG_RAW_KLA) != 0;
    double sp3[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)                                                                   // np.clip (:133)
        sp3[j] = clip_keep_nan(action[j], 0.0, 1.0) * (pid.bypass ? p.kla_max : p.action_scale);
    TailArgs a;
    a.kla = 0.0; a.q = q_fill; a.ec_conc = 0.0; a.load = load;
    int status = 0;
    double kla_mean[3] = {0.0, 0.0, 0.0};
    DrawOut d;
    if (MODE == SBR_MODE_DP45) {
        double kla = 0.0;
        // phase 1: fill, set-point 0 (gym_SBR_env2.py:54)
        const double sp_fill[1] = {0.0};
        status |= dp45_segment<TAIL_FILL, 1>(x, s, 0, sp_fill, p.kla0, c, a, pid, tol, st, park, kla);
        a.q = 0.0;
        // phases 2..5 react with set-points [0, sp3[0], 0, sp3[1]], each biased by the previous phase's last KLa
        // (SBR_model_FB.py:94,120,146,172)
        const double sp_react[4] = {0.0, sp3[0], 0.0, sp3[1]};
        status |= dp45_segment<TAIL_REACT, 4>(x, s, 1, sp_react, kla, c, a, pid, tol, st, park, kla);
        kla_mean[0] = park.get(PARK_KSUM + 1) / (double)s.n_int[2];
        kla_mean[1] = park.get(PARK_KSUM + 3) / (double)s.n_int[4];
        double sX[10], Xf;
        settle_closed_form(x, s.settle_time, p.settler_area, p.settler_vmax, sX, Xf);
        draw_and_waste(x, sX, Xf, p.Qeff, p.biomass_setpoint, d);
        status |= d.status;
        // phase 8 idle with set-point sp3[2], biased by phase 5's last KLa (:266)
        const double sp_idle[1] = {sp3[2]};
        status |= dp45_segment<TAIL_REACT, 1>(x, s, 7, sp_idle, kla, c, a, pid, tol, st, park, kla);
        kla_mean[2] = park.get(PARK_KSUM + 0) / (double)s.n_int[7];
        cycle_epilogue(x, p, d, kla_mean, status, o);
        return;
    }
    PhaseOut po;
    // phase 1: fill, set-point 0 (gym_SBR_env2.py:54)
    status |= pid_phase<TAIL_FILL, MODE>(x, s.n_int[0], s.n_sub[0], s.interval[0], 0.0, p.kla0, c, a, pid, tol, st, po);
    double kla = po.kla_last;
    a.q = 0.0;
    // phases 2..5 react with set-points [0, sp3[0], 0, sp3[1]], each biased by the previous phase's last KLa
    // (SBR_model_FB.py:94,120,146,172); then settle + draw; then phase 8 idle with set-point sp3[2], biased
    // by phase 5's last KLa (:266).  One loop so that the react stepper is instantiated once.
    for (int j = 0; j < 5; ++j) {
        const int ph = j < 4 ? j + 1 : 7;
        if (j == 4) {
            double sX[10], Xf;
            settle_closed_form(x, s.settle_time, p.settler_area, p.settler_vmax, sX, Xf);
            draw_and_waste(x, sX, Xf, p.Qeff, p.biomass_setpoint, d);
            status |= d.status;
        }
        const double sp = j == 1 ? sp3[0] : (j == 3 ? sp3[1] : (j == 4 ? sp3[2] : 0.0));
        status |= pid_phase<TAIL_REACT, MODE>(x, s.n_int[ph], s.n_sub[ph], s.interval[ph], sp, kla, c, a, pid, tol,
                                              st, po);
        const double mean = po.kla_sum / (double)s.n_int[ph];
        if (j < 4) kla = po.kla_last;
        if (j == 1) kla_mean[0] = mean;
        if (j == 3) kla_mean[1] = mean;
        if (j == 4) kla_mean[2] = mean;
    }
    cycle_epilogue(x, p, d, kla_mean, status, o);
}


// One SoA column of an env: element j at p[j * stride].
struct Column {
    double* p;           // NULL: an output the caller did not ask for (set() is then a no-op)
    int64_t stride;
    SBR_HD double get(int j) const { return p[(int64_t)j * stride]; }
    SBR_HD void set(int j, double v) const { if (p) p[(int64_t)j * stride] = v; }
};

// ---------------------------------------------------------------------------------------------------------
// The same cycle with a TRAJECTORY record (sbr_cycle_v2_traj): what SBR_model_FB.run stacks into its `t`, `x` return
// values (SBR_model_FB.py:71-86 and the like after every phase) and its kla arrays, sampled at the END of every PID
// interval: record k = [t, x[14], KLa of the interval], k counting the 528 intervals of phases 1-5 and 8 in order, plus one
// record for the post-draw state (KLa row 0).  A separate function on purpose: the timed kernels above stay untouched.
// It always runs interval by interval (pid_phase's form, both integrators) -- the passive components only exist per
// interval there -- so in adaptive mode its step sequence is not the three-segment kernel's; both agree within the
// tolerance.  traj: record r, row j at traj.p[(r * SBR_TRAJ2_ROWS + j) * traj.stride].
// ---------------------------------------------------------------------------------------------------------
template <int TAIL, int MODE>
SBR_HD int pid_phase_traj(double (&x)[SBR_NX], int n_int, int n_sub, double T, double sp, double kla_in, const Coef& c,
                          TailArgs a, const PidA& pid, const SbrTol& tol, Dp45State& st, PhaseOut& out, double t0,
                          const Column& traj, int& rec) {
    double bias = kla_in, ie = 0.0, so_prev = 0.0, so_i = x[iSo];
    double ksum = 0.0, kla = kla_in;
    int status = 0;
    for (int i = 0; i < n_int; ++i) {
        kla = pid_a_update(pid, sp, so_i, so_prev, i == 0, ie, bias);
        a.kla = kla;
        status |= integrate_interval<TAIL, MODE>(x, T, n_sub, c, a, tol, st);
        ksum += kla;
        so_prev = so_i;
        so_i = x[iSo];
        const int base = rec * SBR_TRAJ2_ROWS;
        traj.set(base + SBR_TRAJ2_T, fma((double)(i + 1), T, t0));
#pragma unroll
        for (int j = 0; j < SBR_NX; ++j) traj.set(base + SBR_TRAJ2_X + j, x[j]);
        traj.set(base + SBR_TRAJ2_KLA, kla);
        ++rec;
    }
    out.kla_sum = ksum;
    out.kla_last = kla;
    return status;
}

template <int MODE>
SBR_HD void cycle_v2_traj(double (&x)[SBR_NX], const double (&action)[3], Loading load, double q_fill,
                          const SbrParams& p, const Coef& c, const SbrSchedule& s, const SbrTol& tol, Dp45State& st,
                          CycleOut& o, const double (&t_start)[SBR_NPHASE], const Column& traj) {
    PidA pid = make_pid_a(p, c);
    pid.bypass = (tol.flags & SBR_FLAG_RAW_KLA) != 0;
    double sp3[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        sp3[j] = clip_keep_nan(action[j], 0.0, 1.0) * (pid.bypass ? p.kla_max : p.action_scale);
    TailArgs a;
    a.kla = 0.0; a.q = q_fill; a.ec_conc = 0.0; a.load = load;
    int status = 0, rec = 0;
    double kla_mean[3] = {0.0, 0.0, 0.0};
    DrawOut d;
    PhaseOut po;
    status |= pid_phase_traj<TAIL_FILL, MODE>(x, s.n_int[0], s.n_sub[0], s.interval[0], 0.0, p.kla0, c, a, pid, tol, st, po,
                                              t_start[0], traj, rec);
    double kla = po.kla_last;
    a.q = 0.0;
    for (int j = 0; j < 5; ++j) {
        const int ph = j < 4 ? j + 1 : 7;
        if (j == 4) {
            double sX[10], Xf;
            settle_closed_form(x, s.settle_time, p.settler_area, p.settler_vmax, sX, Xf);
            draw_and_waste(x, sX, Xf, p.Qeff, p.biomass_setpoint, d);
            status |= d.status;
            // the post-draw state, stamped with the start of the idle phase
            const int base = rec * SBR_TRAJ2_ROWS;
            traj.set(base + SBR_TRAJ2_T, t_start[7]);
#pragma unroll
            for (int k = 0; k < SBR_NX; ++k) traj.set(base + SBR_TRAJ2_X + k, x[k]);
            traj.set(base + SBR_TRAJ2_KLA, 0.0);
            ++rec;
        }
        const double sp = j == 1 ? sp3[0] : (j == 3 ? sp3[1] : (j == 4 ? sp3[2] : 0.0));
        status |= pid_phase_traj<TAIL_REACT, MODE>(x, s.n_int[ph], s.n_sub[ph], s.interval[ph], sp, kla, c, a, pid, tol, st,
                                                   po, t_start[ph], traj, rec);
        const double mean = po.kla_sum / (double)s.n_int[ph];
        if (j < 4) kla = po.kla_last;
        if (j == 1) kla_mean[0] = mean;
        if (j == 3) kla_mean[1] = mean;
        if (j == 4) kla_mean[2] = mean;
    }
    cycle_epilogue(x, p, d, kla_mean, status, o);
}


// =========================================================================================================
// Path B: the interval-per-step env SbrOS (gym_SBR_oneshot.py).  One env.step = one (at phase boundaries two)
// 72-s PID interval; DO-PID -> KLa in aerobic phases, NO3-PID -> external-carbon flow EC in anoxic phases; the
// last step also runs settle + draw + idle.  The reference keeps So / Sno / Kla / EC as ever-growing Python
// lists in module globals; what the arithmetic actually reads back is So[-2:], Sno[-2:], EC[-1], the two PID
// integrals and the last 10 entries of Kla -- that is the persistent per-env state (SBR_OS_* rows).
// =========================================================================================================

// IEEE round-to-nearest add / sub / div that the compiler may neither contract nor re-associate: the number of
// output points of an interval, L = int(((t + t_delta) - t) / dt), flips between 9 and 10 with the rounding of
// the running time (gym_SBR_oneshot.py:1339,1384) and feeds the reward (module_reward_EQIOCI.py:70,79).
SBR_HD double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
SBR_HD double sub_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b;
    return r;
#endif
}
SBR_HD double div_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}

// true if the predicate holds for any env of the warp (CPU twin: for this env).  Used to pick the cheaper
// react tail when no env of the warp doses carbon: with ec == 0 the EC tail IS the react tail.
SBR_HD bool warp_any(bool pred) {
#ifdef SBR_SINGLE_TAIL      // A/B build: always the dosing tail (one stepper instance instead of two, less code to stream):
                            // measured slower, 0.250 against 0.224 ms per SBROS-v1 step (profiles/r02av_*), not taken
    return true;
#endif
#ifdef __CUDA_ARCH__
    return __any_sync(__activemask(), pred) != 0;
#else
    return pred;
#endif
}


// The env's KLa history (the last 10 entries of the reference's ever-growing `Kla` list) as a CIRCULAR buffer: the
// entry of the k-th interval since the reset sits in slot k % 10, so that a step writes one slot (two at a phase
// switch) instead of shifting ten.  The slot of the next push follows from the running time (every interval advances
// it by t_delta and pushes exactly one entry), so no head pointer is stored.  `in`: where the entries are read from
// (the state rows in global memory, or their staged copy in shared memory); `out`: the state rows in global memory.
// A push goes to both, so that later steps of the same launch see it.
struct KlaRing {
    Column in, out;
    int head;                                                          // slot of the next push = oldest entry
    SBR_HD double back(int m) const {                                  // m = 1: newest ... m = 10: oldest
        int j = head - m;
        j += j < 0 ? 10 : 0;
        return in.get(j);
    }
    SBR_HD void push(double v) {
        out.set(head, v);
        if (in.p != out.p) in.set(head, v);
        head = head == 9 ? 0 : head + 1;
    }
};
SBR_HD int os_interval_count(double t, const SbrOsSchedule& s) {       // intervals run since the reset
    const int k = (int)((t - s.t_fill) / s.t_delta + 0.5);
    return k > 0 ? k : 0;
}
SBR_HD int os_ring_head(double t, const SbrOsSchedule& s) { return os_interval_count(t, s) % 10; }

// Optional trajectory dump (the reference's trajectory(), gym_SBR_oneshot.py:1275-1288, sampled at the ENDS of the
// PID intervals): record k = the state after the k-th interval since the reset, rows SBR_TRAJ_* of include/sbr_b200.h;
// the terminal step appends two more records (after settle + draw, after the idle phase).  p == NULL: off (the timed
// path).  Records beyond `cap` are dropped.
struct OsTraj {
    double* p;           // this env's column of traj[cap][SBR_TRAJ_ROWS][ld]
    int64_t ld;
    int cap;
    SBR_HD bool on(int k) const { return p != nullptr && k >= 0 && k < cap; }
    SBR_HD void put(int k, int row, double v) const { p[((int64_t)k * SBR_TRAJ_ROWS + row) * ld] = v; }
    SBR_HD void state(int k, double t, const double (&x)[SBR_NX], double kla, double ec, double u_do, double u_ec) const {
        if (!on(k)) return;
        put(k, SBR_TRAJ_T, t);
#pragma unroll
        for (int i = 0; i < SBR_NX; ++i) put(k, SBR_TRAJ_X + i, x[i]);
        put(k, SBR_TRAJ_KLA, kla); put(k, SBR_TRAJ_EC, ec); put(k, SBR_TRAJ_U_DO, u_do); put(k, SBR_TRAJ_U_EC, u_ec);
        put(k, SBR_TRAJ_REWARD, NAN); put(k, SBR_TRAJ_EQI, NAN); put(k, SBR_TRAJ_OCI, NAN);
        put(k, SBR_TRAJ_AE, NAN); put(k, SBR_TRAJ_ECO, NAN);
    }
};

struct OsPid {
    double Kc_DO, KcI_DO, KcD_DO, Kc_EC, KcI_EC, KcD_EC, dt, inv_dt, kla_lo, kla_hi, ec_lo, ec_hi;
};

SBR_HD OsPid make_os_pid(const SbrParams& p, const Coef& c) {
    OsPid q;
    q.Kc_DO = p.os_Kc_DO; q.KcI_DO = c.os_KcI_DO; q.KcD_DO = c.os_KcD_DO;
    q.Kc_EC = p.os_Kc_EC; q.KcI_EC = c.os_KcI_EC; q.KcD_EC = c.os_KcD_EC;
    q.dt = p.os_pid_dt; q.inv_dt = c.os_inv_dt; q.kla_lo = p.kla_min; q.kla_hi = p.kla_max; q.ec_lo = p.ec_min; q.ec_hi = p.ec_max;
    return q;
}

// Controller scalars of one env.  So[-1] is always x[8] and is passed in.
struct OsCtrl {
    double t, so_prev, sno_last, sno_prev, ie_do, ie_ec, ec_last, kla_last;
};

// DO-PID (gym_SBR_oneshot.py:1889-1909; anoxic variant :1975-1990 advances the integral but sets Kla := 0).
// Incremental bias Kla[-1]; PID dt = 0.002/24 (NOT the 72-s control interval); two independent clamp checks,
// each undoing the integral update.  `first` = the t_start == 0 branch taken only by reset.
SBR_HD double os_pid_do(OsCtrl& c, double so_last, double sp, bool first, bool aerobic, const OsPid& q) {
    const double e = sp - so_last;
    double dcv = 0.0;
    if (!first) {
        dcv = (so_last - c.so_prev) * q.inv_dt;      // the reference divides by dt: <= 1 ulp apart
        c.ie_do = c.ie_do + e * q.dt;
    } else {
        c.ie_do = 0.0;
    }
    double kla = aerobic ? q.Kc_DO * e + q.KcI_DO * c.ie_do + q.KcD_DO * dcv + c.kla_last : 0.0;
    if (kla > q.kla_hi) { kla = q.kla_hi; c.ie_do = c.ie_do - e * q.dt; }
    if (kla < q.kla_lo) { kla = q.kla_lo; c.ie_do = c.ie_do - e * q.dt; }
    return kla;
}

// NO3-PID -> external carbon flow (gym_SBR_oneshot.py:1917-1948; aerobic variant :1925-1937 sets EC := 0).
// Error sign reversed (Sno - sp); lower clamp first, `elif` upper.
SBR_HD double os_pid_ec(OsCtrl& c, double sp, bool dosing, const OsPid& q) {
    const double e = c.sno_last - sp;
    const double dcv = (c.sno_last - c.sno_prev) * q.inv_dt;
    c.ie_ec = c.ie_ec + e * q.dt;
    double ec = dosing ? q.Kc_EC * e + q.KcI_EC * c.ie_ec + q.KcD_EC * dcv + c.ec_last : 0.0;
    if (ec < q.ec_lo) { ec = q.ec_lo; c.ie_ec = c.ie_ec - e * q.dt; }
    else if (ec > q.ec_hi) { ec = q.ec_hi; c.ie_ec = c.ie_ec - e * q.dt; }
    return ec;
}

SBR_HD double clip1(double v) { return v > 1.0 ? 1.0 : (v < -1.0 ? -1.0 : v); }

// The six components whose change over the step enters the observations: Ss, Xbh, Xba, So, Sno, Snh.
struct ObsRef { double Ss, Xbh, Xba, So, Sno, Snh; };
SBR_HD ObsRef obs_ref(const double (&x)[SBR_NX]) {
    ObsRef r;
    r.Ss = x[iSs]; r.Xbh = x[iXbh]; r.Xba = x[iXba]; r.So = x[iSo]; r.Sno = x[iSno]; r.Snh = x[iSnh];
    return r;
}

// Observation scales are fixed API constants of the reference (gym_SBR_oneshot.py:150-156, 1069-1112); the
// kernels multiply by their reciprocals (an IEEE double division is ~40 SASS instructions, and the epilogue has 42
// of them) -- at most 1 ulp away from the reference's x / scale.
#define SBR_INV(v) (1.0 / (v))

// Clipped deltas shared by the reset and step observations (gym_SBR_oneshot.py:388-429, 1069-1112).
SBR_HD void os_emit_deltas(const double (&x)[SBR_NX], const ObsRef& f, const Column& obs_do, const Column& obs_ec) {
    const double dXbh = clip1((x[iXbh] - f.Xbh) * SBR_INV(4000.0)), dSnh = clip1((x[iSnh] - f.Snh) * SBR_INV(50.0));
    obs_do.set(5, dXbh);
    obs_do.set(6, clip1((x[iXba] - f.Xba) * SBR_INV(500.0)));
    obs_do.set(7, clip1((x[iSo] - f.So) * SBR_INV(8.0)));
    obs_do.set(8, dSnh);
    obs_ec.set(5, clip1((x[iSs] - f.Ss) * SBR_INV(50.0)));
    obs_ec.set(6, dXbh);
    obs_ec.set(7, clip1((x[iSno] - f.Sno) * SBR_INV(50.0)));
    obs_ec.set(8, dSnh);
}

// 1 / x_1_state[1:] (gym_SBR_oneshot.py:153)
SBR_HD constexpr double inv_state_scale(int i) {
    return i == iV ? SBR_INV(1.32) : i == iSi ? SBR_INV(30.0) : i == iSs ? SBR_INV(30.0) : i == iXi ? SBR_INV(1500.0)
         : i == iXs ? SBR_INV(150.0) : i == iXbh ? SBR_INV(3000.0) : i == iXba ? SBR_INV(2000.0)
         : i == iXp ? SBR_INV(600.0) : i == iSo ? SBR_INV(8.0) : i == iSno ? SBR_INV(20.0) : i == iSnh ? SBR_INV(20.0)
         : i == iSnd ? SBR_INV(10.0) : i == iXnd ? SBR_INV(10.0) : SBR_INV(10.0);
}

// Step observation epilogue (gym_SBR_oneshot.py:1015-1112): obs_DO = [t, Xbh, Xba, So, Snh] / x_1_DO,
// obs_EC = [t, Ss, Xbh, Sno, Snh] / x_1_EC (:150-156), each followed by 4 clipped deltas; state = [t, x] / x_1_state.
SBR_HD void os_emit_obs(double t, const double (&x)[SBR_NX], const ObsRef& first, const Column& obs_do,
                        const Column& obs_ec, const Column& state) {
    const double tn = t * 2.0, xbh = x[iXbh] * SBR_INV(2000.0), snh = x[iSnh] * SBR_INV(10.0);
    obs_do.set(0, tn); obs_do.set(1, xbh); obs_do.set(2, x[iXba] * SBR_INV(500.0));
    obs_do.set(3, x[iSo] * SBR_INV(8.0)); obs_do.set(4, snh);
    obs_ec.set(0, tn); obs_ec.set(1, x[iSs] * SBR_INV(30.0)); obs_ec.set(2, xbh);
    obs_ec.set(3, x[iSno] * SBR_INV(10.0)); obs_ec.set(4, snh);
    os_emit_deltas(x, first, obs_do, obs_ec);
    state.set(0, tn);
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) state.set(i + 1, x[i] * inv_state_scale(i));
}

// SbrOS.reset (gym_SBR_oneshot.py:168-438) + Sim_filling (:1585-1654).  x: in = x0, out = state after the fill.
template <int MODE>
SBR_HD int os_reset_env(double (&x)[SBR_NX], const Loading& load, const SbrParams& p, const Coef& coef,
                        const SbrOsSchedule& s, const SbrTol& tol, Dp45State& dp, OsCtrl& c, const Column& ring,
                        const Column& obs_do, const Column& obs_ec) {
    const OsPid pid = make_os_pid(p, coef);
    const ObsRef x0r = obs_ref(x);
    const double so0 = x[iSo], sno0 = x[iSno];
    c.so_prev = so0; c.sno_prev = sno0; c.sno_last = sno0;
    c.ie_do = 0.0; c.ie_ec = 0.0; c.ec_last = 0.0; c.kla_last = 0.0;
    // one DO-PID update at set-point 0 on the t_start == 0 branch (:1597-1617): e < 0, KLa clamps to 0 and the
    // anti-windup leaves ie_DO = So0 * dt; the fill-phase EC controller is forced to 0 (:1620-1645)
    const double kla = os_pid_do(c, so0, 0.0, true, true, pid);
    TailArgs a;
    a.kla = kla; a.q = load(0); a.ec_conc = 0.0; a.load = load;
    const int n_sub = s.rk4_sub_fill > 0 ? s.rk4_sub_fill : s.fill_pts - 1;
    SbrTol tl = tol;
    tl.max_steps = tol.max_steps * (int)ceil(s.t_fill / s.t_delta);     // one solve over ~25 control intervals
    const int status = integrate_interval<TAIL_FILL, MODE>(x, s.t_fill, n_sub, coef, a, tl, dp);
    c.so_prev = so0;                 // So  = [x0[8], x_fill[8]]
    c.sno_prev = sno0;               // Sno = [x0[9], x_fill[2]]  -- sic, Ss stored as Sno (:1652)
    c.sno_last = x[iSs];
    c.kla_last = kla;
    c.t = s.t_fill;
    // `Kla` = [0, kla] replicated to the length of the fill trajectory (:320-324)
#pragma unroll
    for (int j = 0; j < 10; ++j) ring.set(j, (j & 1) ? kla : 0.0);
    // reset observation: flow-weighted mix of influent and reactor content (:347-364)
    const double Qin = p.Qin, IV = p.IV;
    const double iden = 1.0 / (Qin + IV);
    const double mXbh = (Qin * load(iXbh) + x[iXbh] * IV) * iden, mSnh = (Qin * load(iSnh) + x[iSnh] * IV) * iden;
    obs_do.set(0, c.t * 2.0);
    obs_do.set(1, mXbh * SBR_INV(2000.0));
    obs_do.set(2, (Qin * load(iXba) + x[iXba] * IV) * iden * SBR_INV(500.0));
    obs_do.set(3, (Qin * load(iSo) + x[iSo] * IV) * iden * SBR_INV(8.0));
    obs_do.set(4, mSnh * SBR_INV(10.0));
    obs_ec.set(0, c.t * 2.0);
    obs_ec.set(1, (Qin * load(iSs) + x[iSs] * IV) * iden * SBR_INV(30.0));
    obs_ec.set(2, mXbh * SBR_INV(2000.0));
    obs_ec.set(3, (Qin * load(iSno) + x[iSno] * IV) * iden * SBR_INV(10.0));
    obs_ec.set(4, mSnh * SBR_INV(10.0));
    os_emit_deltas(x, x0r, obs_do, obs_ec);
    return status;
}

struct OsStepOut {
    double reward;
    double Qw;       // only meaningful when done
    ObsRef first;    // reference state of the observation deltas (start of the step's last interval)
    int done;
    int status;
};

// SbrOS.step (gym_SBR_oneshot.py:843-1273) without the observation epilogue (the caller emits it with os_emit_obs
// from o.first: a launch that advances K steps stores only the last one's).  ring: the env's 10-entry Kla history
// (read after the integration, so that it costs no registers while the stepper runs).  Passes 0..3 are the reference's four NON-exclusive ifs on
// the running time (:860,896,931,963: anoxic / aerobic / anoxic / aerobic -- a step that crosses a phase boundary
// runs two intervals); pass 4 computes the reward and, at the end of the react phases, settle + draw + the idle
// solve.  All five share ONE stepper call site per tail so the kernel stays inside the instruction cache.
template <int MODE>
SBR_HD void os_step_env(double (&x)[SBR_NX], OsCtrl& c, KlaRing& ring, double a_do, double a_ec,
                        const SbrParams& p, const Coef& coef, const SbrOsSchedule& s, const SbrTol& tol,
                        Dp45State& dp, OsStepOut& o, const OsTraj& traj) {
    const OsPid pid = make_os_pid(p, coef);
    int status = 0, L = 10;
    double span = s.t_delta, u_do = 0.0, u_ec_rec = 0.0, ec_before = c.ec_last;
    ObsRef first = obs_ref(x);
    TailArgs a;
    a.kla = 0.0; a.q = 0.0; a.ec_conc = p.ec_conc; a.load = Loading{nullptr, 0};
    o.done = 0;
    o.Qw = NAN;
    o.reward = 0.0;
    SbrTol tl = tol;
    for (int pass = 0; pass < 5; ++pass) {
        const double t = c.t;
        double T, so_start = x[iSo], t_next;
        int n_sub;
        if (pass < 4) {
            const bool cond = pass == 0 ? (t < s.tm3_0)
                            : pass == 1 ? (t >= s.tm3_0 && t <= s.tm3_1)
                            : pass == 2 ? (t > s.tm3_1 && t <= s.tm4_1)
                                        : (t > s.tm4_1);
            if (!cond) continue;
            const bool aerobic = (pass & 1) != 0;
            u_do = aerobic ? clip_keep_nan(a_do, 0.0, p.do_sp_max) : 0.0;               // :862-870, 898-906
            const double u_ec = aerobic ? 0.0 : clip_keep_nan(a_ec, 0.0, p.no_sp_max);
            u_ec_rec = u_ec;
            // run_aero_step / run_anaero_step (:1331-1419)
            t_next = add_rn(t, s.t_delta);
            span = sub_rn(t_next, t);
            L = (int)div_rn(span, s.dt);
            first = obs_ref(x);
            a.kla = os_pid_do(c, so_start, u_do, false, aerobic, pid);
            ec_before = c.ec_last;
            a.q = os_pid_ec(c, u_ec, !aerobic, pid);
            T = span;
            n_sub = s.rk4_sub_interval > 0 ? s.rk4_sub_interval : (L > 1 ? L - 1 : 1);
        } else {
            // reward = module_reward_EQIOCI.sbr_reward (module_reward_EQIOCI.py:4-115) on the post-interval state:
            // `Kla` holds ONE entry per interval, so Kla[-L:-1] sums the previous L-1 intervals and leaves the
            // current one out (:70-71); `EC` holds L-1 copies per interval, so EC[-L:-1] is the previous
            // interval's flow once plus the current one L-2 times (:79).  Summed oldest first, like sum() does.
            double ksum = 0.0;
            for (int j = L; j >= 2; --j) ksum += ring.back(j);
            double esum = 0.0 + ec_before;
            for (int j = 0; j < L - 2; ++j) esum += c.ec_last;
            double eff[6];
            const double EQI2 = effluent_quality(x, eff) * 0.1;
            const double ispan = rcp(span);
            const double AE = (8 / (1.8 * 1000)) * ispan * (1.32 * ksum * p.os_pid_dt);
            const double ECO = p.ec_conc * esum * p.os_pid_dt * (ispan * 1e-3);
            const double OCI = AE + ECO;
            o.reward = (1 - (EQI2 * EQI2 + OCI * OCI)) * (1.0 / 473);
            if (traj.p) {
                const int k = os_interval_count(c.t, s) - 1;
                if (traj.on(k)) {
                    // the diagnostics the reference appends to reward_EQI_t / reward_OCI_t / reward_AE_t / reward_EC_t:
                    // EQI / 10 and the two cost terms normalised by their maxima (module_reward_EQIOCI.py:72,80,96,109-112)
                    const double dt = p.os_pid_dt;
                    const double AE_max = 1.32 * (240 * 11) * dt * (8 / ((dt * 11) * 1.8 * 1000));
                    const double EC_max = p.ec_conc * (0.0005 * 11) * dt / ((dt * 11) * 1000);
                    traj.put(k, SBR_TRAJ_REWARD, o.reward); traj.put(k, SBR_TRAJ_EQI, EQI2);
                    traj.put(k, SBR_TRAJ_OCI, AE / AE_max + ECO / EC_max);
                    traj.put(k, SBR_TRAJ_AE, AE / AE_max); traj.put(k, SBR_TRAJ_ECO, ECO / EC_max);
                }
            }
            const bool terminal = c.t >= s.tm5_1;
            // end of the react phases (:1122): Sim_Settling_Drawing (:2264-2420) + Sim_idle (:2554-2597) in this
            // step; the reward stays the pre-settle one, obs/state are recomputed from the post-idle state with
            // deltas taken against the end-of-react state (:1167-1261)
            double kla_idle = 0.0;
            T = s.t_delta; n_sub = 1; t_next = s.t_cycle;
            if (terminal) {
                o.done = 1;
                first = obs_ref(x);
                const double t_set_end = add_rn(c.t, s.settle_len);
                const double T_set = sub_rn(t_set_end, c.t);
                double sX[10], Xf;
                settle_closed_form(x, T_set, p.settler_area, p.settler_vmax, sX, Xf);
                DrawOut d;
                draw_and_waste(x, sX, Xf, p.Qeff, p.biomass_setpoint, d);
                status |= d.status;
                o.Qw = d.Qw;
                const double t_draw_end = add_rn(t_set_end, s.draw_len);
                if (traj.p) traj.state(os_interval_count(c.t, s), t_draw_end, x, 0.0, 0.0, u_do, 0.0);
                so_start = x[iSo];
                c.so_prev = so_start;              // So padded with the frozen value over settle + draw (:2415-2416)
                T = sub_rn(s.t_cycle, t_draw_end);
                tl.max_steps = tol.max_steps * (int)ceil(T / s.t_delta);    // one solve over ~36 control intervals
                const int pts = (int)div_rn(T, s.dt);
                n_sub = s.rk4_sub_idle > 0 ? s.rk4_sub_idle : (pts > 1 ? pts - 1 : 1);
                kla_idle = os_pid_do(c, so_start, u_do, false, true, pid);
                a.kla = kla_idle; a.q = 0.0;
                ring.push(kla_idle);
            }
            if (!terminal) break;
        }
        if (warp_any(a.q != 0.0)) status |= integrate_interval<TAIL_EC, MODE>(x, T, n_sub, coef, a, tl, dp);
        else status |= integrate_interval<TAIL_REACT, MODE>(x, T, n_sub, coef, a, tl, dp);
        c.so_prev = so_start;
        c.sno_prev = c.sno_last;
        c.sno_last = x[iSno];
        c.kla_last = a.kla;
        if (traj.p) {
            // interval k ends here; the idle solve of the terminal step is the record after the post-draw one
            const int k = pass < 4 ? os_interval_count(t_next, s) - 1 : os_interval_count(c.t, s) + 1;
            traj.state(k, t_next, x, a.kla, a.q, u_do, pass < 4 ? u_ec_rec : 0.0);
        }
        c.t = t_next;
        if (pass < 4) {
            c.ec_last = a.q;
            ring.push(a.kla);
        }
    }
    bool finite = fabs(o.reward) < 1e300;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) finite = finite && (fabs(x[i]) < 1e300);
    if (!finite) status |= SBR_ST_NONFINITE;
    o.first = first;
    o.status = status;
}

// =========================================================================================================
// SBR-v4 (SbrEnv4, gym_SBR_env4.py): interval-per-step env with the fill phase stepped inside step() and a 1-D
// "change of DO set-point" action.  Same physics pieces as above; the controller is the incremental-bias DO-PID
// with the cycle-per-step tuning (Kc 5, tauI 0.00035, tauD 0.005, :61-69) and PID dt = 0.002/24.
// =========================================================================================================
struct V4Ctrl {
    double t, u, so_prev, ie, kla_last, kla_sum;
};

struct V4Out {
    double reward, Qw;
    int done, status;
};

// 1 / x_1 (gym_SBR_env4.py:91): the observation is x / x_1.
SBR_HD constexpr double inv_x1_v4(int i) {
    return i == 0 ? SBR_INV(1.32000000e+00) : i == 1 ? SBR_INV(3.00000000e+01) : i == 2 ? SBR_INV(3.81606587e+01)
         : i == 3 ? SBR_INV(6.94658685e+02) : i == 4 ? SBR_INV(1.07772100e+02) : i == 5 ? SBR_INV(1.22613841e+03)
         : i == 6 ? SBR_INV(7.88460027e+01) : i == 7 ? SBR_INV(2.57616136e+02) : i == 8 ? SBR_INV(1.01108024e+00)
         : i == 9 ? SBR_INV(6.24510635e+00) : i == 10 ? SBR_INV(1.78877937e+01) : i == 11 ? SBR_INV(3.95743344e+00)
         : i == 12 ? SBR_INV(5.70432163e+00) : SBR_INV(5.50185509e+00);
}

// Controller block shared by Sim_filling / Sim_rxn / Sim_idle (gym_SBR_env4.py:497-524, 667-697, 1202-1234).
SBR_HD double v4_pid(V4Ctrl& c, double so_last, bool first, const SbrParams& p, const Coef& coef) {
    const double dt = p.os_pid_dt;
    const double e = c.u - so_last;
    double dcv = 0.0;
    if (!first) { dcv = (so_last - c.so_prev) * coef.os_inv_dt; c.ie = c.ie + e * dt; }
    else c.ie = 0.0;
    double kla = p.pid_Kc * e + coef.pidA_KcI * c.ie + coef.pidA_KcD * dcv + c.kla_last;
    if (kla > p.kla_max) { kla = p.kla_max; c.ie = c.ie - e * dt; }
    if (kla < p.kla_min) { kla = p.kla_min; c.ie = c.ie - e * dt; }
    return kla;
}

// SbrEnv4.step + run_step (gym_SBR_env4.py:200-358).  load: the env's influent column (used in the fill phase).
template <int MODE>
SBR_HD void v4_step_env(double (&x)[SBR_NX], V4Ctrl& c, double action, const Loading& load, const SbrParams& p,
                        const Coef& coef, const SbrOsSchedule& s, const SbrTol& tol, Dp45State& dp,
                        const Column& obs, V4Out& o) {
    const bool first = c.t == 0.0;
    if (first) { c.u = 0.0; c.kla_last = 0.0; c.kla_sum = 0.0; }       // u = 0 (:209); So = [x[8]], Kla = [0] (:272-277)
    c.u = c.u + action;
    c.u = c.u < 0.0 ? 0.0 : (c.u > p.do_sp_max ? p.do_sp_max : c.u);    // :213-218
    const double t = c.t;
    // batch type from the running time (:256-268): 0 fill, 1 react (phases 2-5), 2 settle + draw + idle
    const int bt = (t >= 0.0 && t < s.t_fill) ? 0 : (t < s.tm5_1 ? 1 : 2);
    TailArgs a;
    a.kla = 0.0; a.q = 0.0; a.ec_conc = 0.0; a.load = load;
    int status = 0;
    double so_start = x[iSo], t_next, T;
    int n_sub;
    SbrTol tl = tol;
    o.Qw = NAN;
    double eff_snh = 0.0;
    if (bt < 2) {
        t_next = add_rn(t, s.t_delta);
        T = sub_rn(t_next, t);
        const int L = (int)div_rn(T, s.dt);                              // len(t_range) (:286)
        n_sub = s.rk4_sub_interval > 0 ? s.rk4_sub_interval : (L > 1 ? L - 1 : 1);
        a.kla = v4_pid(c, so_start, first, p, coef);
        if (bt == 0) a.q = load(0);
    } else {
        // Sim_Settling_Drawing (:919-1070; its `dt` argument is the control interval) then Sim_idle (:1202-1242)
        const double t_set_end = add_rn(t, s.settle_len);
        double sX[10], Xf;
        settle_closed_form(x, sub_rn(t_set_end, t), p.settler_area, p.settler_vmax, sX, Xf);
        DrawOut d;
        draw_and_waste(x, sX, Xf, p.Qeff, p.biomass_setpoint, d);
        status |= d.status;
        o.Qw = d.Qw;
        eff_snh = d.eff[3];
        const double t_draw_end = add_rn(t_set_end, s.draw_len);
        so_start = x[iSo];
        c.so_prev = so_start;                                            // So extended with the frozen value (:1063-1064)
        T = sub_rn(s.t_cycle, t_draw_end);
        tl.max_steps = tol.max_steps * (int)ceil(T / s.t_delta);
        const int pts = (int)div_rn(T, s.dt);
        n_sub = s.rk4_sub_idle > 0 ? s.rk4_sub_idle : (pts > 1 ? pts - 1 : 1);
        a.kla = v4_pid(c, so_start, false, p, coef);
        t_next = s.t_cycle;
    }
    if (bt == 0) status |= integrate_interval<TAIL_FILL, MODE>(x, T, n_sub, coef, a, tl, dp);
    else status |= integrate_interval<TAIL_REACT, MODE>(x, T, n_sub, coef, a, tl, dp);
    c.so_prev = so_start;
    c.kla_last = a.kla;
    c.kla_sum += a.kla;
    c.t = t_next;
    // module_reward_continuous.sbr_reward (module_reward_continuous.py:4-65)
    const double tdl = 0.002 / 24;
    double PE, AE_dT, r_snh = 0.0;
    if (bt == 0) { PE = 0.004 * p.Qin; AE_dT = 1.32 * a.kla * tdl; }
    else if (bt == 1) { PE = 0.0; AE_dT = 1.32 * a.kla * tdl; }
    else {
        PE = 0.05 * o.Qw + 0.004 * p.Qeff;
        AE_dT = 1.32 * c.kla_sum * tdl;
        r_snh = eff_snh < 4 ? 0.0 : -246.0;
    }
    const double AE = p.so_sat / (1.8 * 1000) * AE_dT;
    o.reward = (0.5 - (AE + PE)) + r_snh;
    o.done = (bt == 2 && c.t >= s.t_cycle) ? 1 : 0;
    bool finite = fabs(o.reward) < 1e300;
#pragma unroll
    for (int i = 0; i < SBR_NX; ++i) { finite = finite && (fabs(x[i]) < 1e300); obs.set(i, x[i] * inv_x1_v4(i)); }
    if (!finite) status |= SBR_ST_NONFINITE;
    o.status = status;
}

// SbrEnv4.reset observation (gym_SBR_env4.py:185-191): x_2[0] = Qin + IV, x_2[i] = (Qin c_in,i + x0_i IV)/(Qin + IV).
SBR_HD void v4_reset_obs(const double (&x0)[SBR_NX], const Loading& load, const SbrParams& p, const Column& obs) {
    const double Qin = p.Qin, IV = p.IV;
    const double iden = 1.0 / (Qin + IV);
    obs.set(0, (Qin + IV) * inv_x1_v4(0));
#pragma unroll
    for (int i = 1; i < SBR_NX; ++i) obs.set(i, (Qin * load(i) + x0[i] * IV) * iden * inv_x1_v4(i));
}

}  // namespace sbr
