// family.cuh -- registered device-side integrand family and the infinity_handling.jl variable
// transforms (src/infinity_handling.jl:91-125, 201-237, 312-356), evaluated on the device.
//
// The evaluation order of every family is the definition shared with the CPU oracle
// (oracle/oracle.c, orc_family_eval): explicit fma() where the definition fuses, plain * and +
// elsewhere (this library is compiled with -fmad=false so nvcc never contracts on its own).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/integrals_b200.h"

namespace ib {

__host__ __device__ constexpr int fam_nparam(int fam, int d) {
    return (fam == IB200_FAM_SINXP || fam == IB200_FAM_GLTEST) ? 1
         : (fam >= IB200_FAM_GENZ_OSC && fam <= IB200_FAM_GENZ_DISC) ? 2 * d
         : (fam == IB200_FAM_COSPROD || fam == IB200_FAM_EXPLIN) ? d
         : (fam == IB200_FAM_GAUSSPOLY) ? 3 * d + 1 : -1;
}
__host__ __device__ constexpr bool fam_dim_ok(int fam, int d) {
    return d >= 1 && d <= IB200_MAX_DIM && ((fam != IB200_FAM_SINXP && fam != IB200_FAM_GLTEST) || d == 1);
}

#define IB_TWO_PI 6.283185307179586476925286766559
#define IB_HALF_PI 1.5707963267948966

// ------------------------------------------------------------------------------------------------
// Branch-free exp for the integrand inner loops.  CUDA's exp() carries a data-dependent branch
// (the |x| > 708 slow path) in every call, which fences the instruction scheduler: consecutive,
// independent evaluations of a rule cannot be interleaved and the FP64 pipe idles on the 13-deep
// Horner chain (ncu: "wait" stalls, profiles/r01_*).  Same construction as libm (Cody-Waite
// reduction by ln2, degree-13 polynomial, exponent reconstruction in two halves so that gradual
// underflow is honoured), extremes handled by selects.  Maximum error 0.99 ulp over [-750, 700]
// (measured against long double, 2e7 samples); exp_bf(-inf) = 0, exp_bf(+inf) = inf, NaN -> NaN.
// ------------------------------------------------------------------------------------------------
// The constants of the branch-free exp in CONSTANT memory: an FP64 instruction cannot carry a 64-bit immediate (every literal is two MOVs
// into a register pair before the DFMA), but it can take a constant-bank operand directly.  Used only where it measured faster -- the
// exponentials of the rolled 6..8-D rule (exp_bfN<N, true>); the unrolled 2..5-D rules and the trigonometric kernels keep literals
// (measured on BASELINE configs 1 and 4: constant operands cost them 7-10 %).
__constant__ double c_exp[18] = {
    6755399441055744.0, 1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
    2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
    1.6666666666666666e-01, 0.5, -745.2, 709.79 };
#ifndef IB200_LIBM_EXP
__device__ __forceinline__ double exp_bf(double x) {
    const double MG = 6755399441055744.0;                       // 1.5 * 2^52: round to nearest integer
    const double t = fma(x, 1.4426950408889634074, MG);
    int k = __double2loint(t);
    const double kf = t - MG;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double q = 1.6059043836821613e-10;                          // 1/13! ... 1/2!
    q = fma(q, r, 2.08767569878681e-09);
    q = fma(q, r, 2.505210838544172e-08);
    q = fma(q, r, 2.755731922398589e-07);
    q = fma(q, r, 2.7557319223985893e-06);
    q = fma(q, r, 2.48015873015873e-05);
    q = fma(q, r, 1.984126984126984e-04);
    q = fma(q, r, 1.388888888888889e-03);
    q = fma(q, r, 8.333333333333333e-03);
    q = fma(q, r, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    const double p = fma(q, r * r, r) + 1.0;
    k = max(-2000, min(2000, k));
    const int k1 = k >> 1, k2 = k - k1;
    double res = (p * __hiloint2double((k1 + 1023) << 20, 0)) * __hiloint2double((k2 + 1023) << 20, 0);
    res = x < -745.2 ? 0.0 : res;
    res = x > 709.79 ? __longlong_as_double(0x7ff0000000000000LL) : res;
    return res;
}
#else
__device__ __forceinline__ double exp_bf(double x) { return exp(x); }
#endif

// N independent evaluations in lock step.  ptxas schedules a lone Horner chain as back-to-back
// dependent DFMAs (each waits out the FP64 latency); writing the N chains side by side in the source
// gives it N independent instructions between dependent ones (ncu: "wait" stall, profiles/r01_*).
template <int N, bool CM = false>
__device__ __forceinline__ void exp_bfN(const double (&x)[N], double (&out)[N]) {
#ifdef IB200_LIBM_EXP
#pragma unroll
    for (int j = 0; j < N; j++) out[j] = exp(x[j]);
#else
    // CM: the constants come from constant memory (c_exp) instead of 64-bit literals.  Measured both ways: the rolled 6..8-D rule
    // (gm_rule_thread3.cuh), short of registers, is faster with constant operands; the unrolled 2..5-D rules lose 7-9 % with them
    // (Gaussian 414 -> 444 ms, C0 324 -> 354 ms on BASELINE config 4), so literals stay the default.
    const double MG = CM ? c_exp[0] : 6755399441055744.0;
    double t[N], kf[N], r[N], q[N];
#pragma unroll
    for (int j = 0; j < N; j++) t[j] = fma(x[j], CM ? c_exp[1] : 1.4426950408889634074, MG);
#pragma unroll
    for (int j = 0; j < N; j++) kf[j] = t[j] - MG;
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], CM ? c_exp[2] : -6.93147180369123816490e-01, x[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], CM ? c_exp[3] : -1.90821492927058770002e-10, r[j]);
#pragma unroll
    for (int j = 0; j < N; j++) q[j] = fma(CM ? c_exp[4] : 1.6059043836821613e-10, r[j], CM ? c_exp[5] : 2.08767569878681e-09);
#define IB_STEP(i, c) _Pragma("unroll") for (int j = 0; j < N; j++) q[j] = fma(q[j], r[j], CM ? c_exp[i] : (c));
    IB_STEP(6, 2.505210838544172e-08) IB_STEP(7, 2.755731922398589e-07) IB_STEP(8, 2.7557319223985893e-06)
    IB_STEP(9, 2.48015873015873e-05) IB_STEP(10, 1.984126984126984e-04) IB_STEP(11, 1.388888888888889e-03)
    IB_STEP(12, 8.333333333333333e-03) IB_STEP(13, 4.1666666666666664e-02) IB_STEP(14, 1.6666666666666666e-01) IB_STEP(15, 0.5)
#undef IB_STEP
    double pp[N];
#pragma unroll
    for (int j = 0; j < N; j++) pp[j] = r[j] * r[j];
#pragma unroll
    for (int j = 0; j < N; j++) pp[j] = fma(q[j], pp[j], r[j]);
#pragma unroll
    for (int j = 0; j < N; j++) pp[j] = pp[j] + 1.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        int k = max(-2000, min(2000, __double2loint(t[j])));
        const int k1 = k >> 1, k2 = k - k1;
        double res = (pp[j] * __hiloint2double((k1 + 1023) << 20, 0)) * __hiloint2double((k2 + 1023) << 20, 0);
        res = x[j] < (CM ? c_exp[16] : -745.2) ? 0.0 : res;
        out[j] = x[j] > (CM ? c_exp[17] : 709.79) ? __longlong_as_double(0x7ff0000000000000LL) : res;
    }
#endif
}

// Branch-free cos for |x| < 2^30 (3-term Cody-Waite reduction by pi/2 with FMA, fdlibm kernel
// polynomials chosen per quadrant by selects; max error 1.5 ulp measured against long double); larger
// arguments take libm's Payne-Hanek path.
template <int N>
__device__ __forceinline__ void cos_bfN(const double (&x)[N], double (&out)[N]) {
#ifdef IB200_LIBM_EXP
#pragma unroll
    for (int j = 0; j < N; j++) out[j] = cos(x[j]);
#else
    const double MG = 6755399441055744.0;
    double t[N], kf[N], r[N], z[N], p[N];
    bool big = false;
#pragma unroll
    for (int j = 0; j < N; j++) { t[j] = fma(x[j], 0.63661977236758134308, MG); big = big || !(fabs(x[j]) < 1073741824.0); }
#pragma unroll
    for (int j = 0; j < N; j++) kf[j] = t[j] - MG;
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], -1.5707963267948966, x[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], -6.123233995736766e-17, r[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], 1.4973849048591698e-33, r[j]);
    int iq[N];
#pragma unroll
    for (int j = 0; j < N; j++) { iq[j] = __double2loint(t[j]) + 1; z[j] = r[j] * r[j]; }
#pragma unroll
    for (int j = 0; j < N; j++) p[j] = (iq[j] & 1) ? -1.13596475577881948265e-11 : 1.58969099521155010221e-10;
#define IB_STEP(cc, cs) _Pragma("unroll") for (int j = 0; j < N; j++) p[j] = fma(p[j], z[j], (iq[j] & 1) ? (cc) : (cs));
    IB_STEP(2.08757232129817482790e-09, -2.50507602534068634195e-08)
    IB_STEP(-2.75573143513906633035e-07, 2.75573137070700676789e-06)
    IB_STEP(2.48015872894767294178e-05, -1.98412698298579493134e-04)
    IB_STEP(-1.38888888888741095749e-03, 8.33333333332248946124e-03)
    IB_STEP(4.16666666666666019037e-02, -1.66666666666666324348e-01)
#undef IB_STEP
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double sres = fma(r[j] * z[j], p[j], r[j]);
        const double hz = 0.5 * z[j], w = 1.0 - hz;
        const double cres = w + (((1.0 - w) - hz) + (z[j] * z[j]) * p[j]);
        const double res = (iq[j] & 1) ? cres : sres;
        out[j] = (iq[j] & 2) ? -res : res;
    }
    if (big) {
#pragma unroll
        for (int j = 0; j < N; j++) out[j] = cos(x[j]);
    }
#endif
}

// Branch-free sin of N arguments in lock step: cos_bfN with the quadrant counted from 0 instead of 1 (sin x = cos(x - pi/2)): the same
// reduction, one polynomial per argument with the coefficients chosen by the quadrant.  |x| >= 2^30 takes libm's sin.  Max error 1.5 ulp.
// sin_bfN_fast: the branch-free part alone; true when some |x| >= 2^30 (or NaN) and the caller has to redo the group through libm
// (see sincos_bfN_fast: one out-of-line slow path per caller instead of one inlined per group).
template <int N>
__device__ __forceinline__ bool sin_bfN_fast(const double (&x)[N], double (&out)[N]) {
    const double MG = 6755399441055744.0;
    double t[N], kf[N], r[N], z[N], p[N];
    bool big = false;
#pragma unroll
    for (int j = 0; j < N; j++) { t[j] = fma(x[j], 0.63661977236758134308, MG); big = big || !(fabs(x[j]) < 1073741824.0); }
#pragma unroll
    for (int j = 0; j < N; j++) kf[j] = t[j] - MG;
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], -1.5707963267948966, x[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], -6.123233995736766e-17, r[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], 1.4973849048591698e-33, r[j]);
    int iq[N];
#pragma unroll
    for (int j = 0; j < N; j++) { iq[j] = __double2loint(t[j]); z[j] = r[j] * r[j]; }
#pragma unroll
    for (int j = 0; j < N; j++) p[j] = (iq[j] & 1) ? -1.13596475577881948265e-11 : 1.58969099521155010221e-10;
#define IB_STEP(cc, cs) _Pragma("unroll") for (int j = 0; j < N; j++) p[j] = fma(p[j], z[j], (iq[j] & 1) ? (cc) : (cs));
    IB_STEP(2.08757232129817482790e-09, -2.50507602534068634195e-08)
    IB_STEP(-2.75573143513906633035e-07, 2.75573137070700676789e-06)
    IB_STEP(2.48015872894767294178e-05, -1.98412698298579493134e-04)
    IB_STEP(-1.38888888888741095749e-03, 8.33333333332248946124e-03)
    IB_STEP(4.16666666666666019037e-02, -1.66666666666666324348e-01)
#undef IB_STEP
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double sres = fma(r[j] * z[j], p[j], r[j]);
        const double hz = 0.5 * z[j], w = 1.0 - hz;
        const double cres = w + (((1.0 - w) - hz) + (z[j] * z[j]) * p[j]);
        const double res = (iq[j] & 1) ? cres : sres;
        out[j] = (iq[j] & 2) ? -res : res;
    }
    return big;
}
template <int N>
__device__ __forceinline__ void sin_bfN(const double (&x)[N], double (&out)[N]) {
    if (sin_bfN_fast<N>(x, out)) {
#pragma unroll
        for (int j = 0; j < N; j++) out[j] = sin(x[j]);
    }
}
// libm's sin of n arguments, out of line (the cold path of callers of sin_bfN_fast)
static __device__ __noinline__ void sin_libm_n(const double *x, double *out, int n) {
    for (int j = 0; j < n; j++) out[j] = sin(x[j]);
}

// Branch-free sin and cos of N arguments in lock step: the reduction and the two fdlibm kernel polynomials of cos_bfN,
// both evaluated, quadrant fix-up by selects.  |x| >= 2^30 takes libm's sincos.  Max error 1.5 ulp each.
// sincos_bfN_fast: the branch-free part alone; returns true when some |x| >= 2^30 (or NaN) and the caller has to redo the group
// through libm.  Callers that evaluate many groups in straight-line code (the angle-addition Genz-Malik rule) collect the flag and
// keep ONE out-of-line slow path, so that libm's Payne-Hanek code does not sit between the groups in the instruction stream.
template <int N>
__device__ __forceinline__ bool sincos_bfN_fast(const double (&x)[N], double (&sn)[N], double (&cs)[N]) {
    const double MG = 6755399441055744.0;
    double t[N], kf[N], r[N], z[N], ps[N], pc[N];
    bool big = false;
#pragma unroll
    for (int j = 0; j < N; j++) { t[j] = fma(x[j], 0.63661977236758134308, MG); big = big || !(fabs(x[j]) < 1073741824.0); }
#pragma unroll
    for (int j = 0; j < N; j++) kf[j] = t[j] - MG;
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], -1.5707963267948966, x[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], -6.123233995736766e-17, r[j]);
#pragma unroll
    for (int j = 0; j < N; j++) r[j] = fma(kf[j], 1.4973849048591698e-33, r[j]);
#pragma unroll
    for (int j = 0; j < N; j++) { z[j] = r[j] * r[j]; ps[j] = 1.58969099521155010221e-10; pc[j] = -1.13596475577881948265e-11; }
#define IB_STEP(cc, css) _Pragma("unroll") for (int j = 0; j < N; j++) { pc[j] = fma(pc[j], z[j], (cc)); ps[j] = fma(ps[j], z[j], (css)); }
    IB_STEP(2.08757232129817482790e-09, -2.50507602534068634195e-08)
    IB_STEP(-2.75573143513906633035e-07, 2.75573137070700676789e-06)
    IB_STEP(2.48015872894767294178e-05, -1.98412698298579493134e-04)
    IB_STEP(-1.38888888888741095749e-03, 8.33333333332248946124e-03)
    IB_STEP(4.16666666666666019037e-02, -1.66666666666666324348e-01)
#undef IB_STEP
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double sres = fma(r[j] * z[j], ps[j], r[j]);
        const double hz = 0.5 * z[j], w = 1.0 - hz;
        const double cres = w + (((1.0 - w) - hz) + (z[j] * z[j]) * pc[j]);
        const int q = __double2loint(t[j]);
        const double s0 = (q & 1) ? cres : sres, c0 = (q & 1) ? sres : cres;
        sn[j] = (q & 2) ? -s0 : s0;
        cs[j] = ((q + 1) & 2) ? -c0 : c0;
    }
    return big;
}
template <int N>
__device__ __forceinline__ void sincos_bfN(const double (&x)[N], double (&sn)[N], double (&cs)[N]) {
    if (sincos_bfN_fast<N>(x, sn, cs)) {
#pragma unroll
        for (int j = 0; j < N; j++) sincos(x[j], &sn[j], &cs[j]);
    }
}
// libm's sincos of n arguments, out of line (the cold path of callers of sincos_bfN_fast)
static __device__ __noinline__ void sincos_libm_n(const double *x, double *sn, double *cs, int n) {
    for (int j = 0; j < n; j++) sincos(x[j], &sn[j], &cs[j]);
}

// ------------------------------------------------------------------------------------------------
// Axis map: one coordinate of the ChangeOfVariables the reference applies to every tuple-domain
// problem (src/common.jl:97-103).  kind: 0 finite, 1 only lb infinite, 2 only ub infinite, 3 both.
// ------------------------------------------------------------------------------------------------
struct AxisMap {
    double lb, ub;
    double den;        // finite: (ub-lb)*0.5
    double sg;         // semi-infinite: flipsign(t, inf bound) == t*sg
    double tlo, thi;   // transformed bounds (u2t, infinity_handling.jl:91-107)
    int kind;
};

__device__ __forceinline__ double flipsign1(double y) { return signbit(y) ? -1.0 : 1.0; }

__device__ inline void make_axis(double lb, double ub, AxisMap &m) {
    const bool il = isinf(lb), iu = isinf(ub);
    m.lb = lb; m.ub = ub; m.den = 0.0; m.sg = 1.0;
    if (il && iu)      { m.kind = 3; m.tlo = flipsign1(lb); m.thi = flipsign1(ub); }
    else if (il)       { m.kind = 1; m.tlo = flipsign1(lb); m.thi = 0.0; m.sg = flipsign1(lb); }
    else if (iu)       { m.kind = 2; m.tlo = 0.0; m.thi = flipsign1(ub); m.sg = flipsign1(ub); }
    else               { m.kind = 0; m.tlo = -1.0; m.thi = 1.0; m.den = (ub - lb) * 0.5; }
}

// t -> (u, du/dt)   t2ujac / t2ujac_tan / t2ujac_cot
__device__ __forceinline__ void t2ujac(int tr, double t, const AxisMap &m, double &u, double &jac) {
    if (m.kind == 0) {                              // infinity_handling.jl:121-124
        u = m.lb + (1.0 + t) * m.den;
        jac = m.den;
        return;
    }
    if (tr == IB200_TR_IF_INF) {
        if (m.kind == 3) {                          // :112-114
            const double den = 1.0 / (1.0 - t * t);
            u = t * den;
            jac = (1.0 + t * t) * (den * den);
        } else {                                    // :115-120
            const double den = 1.0 / (1.0 - t * m.sg);
            u = (m.kind == 1 ? m.ub : m.lb) + t * den;
            jac = den * den;
        }
        return;
    }
    if (tr == IB200_TR_COT && m.kind == 2) {        // :338-351
        const double arg = IB_HALF_PI * (1.0 - t);
        const double c = 1.0 / tan(arg);
        u = m.lb + c;
        jac = IB_HALF_PI * (1.0 + c * c);
        return;
    }
    {                                               // tan map :203-232 (and cot's fall-backs :314-337)
        const double tv = tan(IB_HALF_PI * t);
        u = (m.kind == 3) ? tv : ((m.kind == 1 ? m.ub : m.lb) + tv);
        jac = IB_HALF_PI * (1.0 + tv * tv);
    }
}

// ------------------------------------------------------------------------------------------------
// Families.  Family<FAM,D>:
//   NQ, prepare(par, q)      per-problem prepared parameters (registers, static indexing)
//   eval(x, q)               the integrand, same operation order as orc_family_eval
//   separable form f(x) = finish( init (+|*) term_0(x_0) (+|*) ... (+|*) term_{D-1}(x_{D-1}) ),
//   used by the tensor-rule kernel to hoist per-axis work out of the innermost loop:
//   kProd, sep_init(par), sep_term(k, x, par), sep_finish(acc, par)   (par: raw params, any memory)
// ------------------------------------------------------------------------------------------------
template <int FAM, int D> struct Family;

__device__ __forceinline__ double cos_bf(double x) { double xi[1] = {x}, o[1]; cos_bfN<1>(xi, o); return o[0]; }

// finish on N independent accumulators in lock step (families with a vectorised transcendental override sep_finishN)
template <typename F, int N>
__device__ __forceinline__ void finishN_default(const double (&acc)[N], const double *par, double (&out)[N]) {
#pragma unroll
    for (int j = 0; j < N; j++) out[j] = F::sep_finish(acc[j], par);
}

template <int D> struct Family<IB200_FAM_SINXP, D> {
    static constexpr int NQ = 1; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) { q[0] = par[0]; }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) { return sin(x[0] * q[0]); }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int, double x, const double *par) { return sin(x * par[0]); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_SINXP, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_GLTEST, D> {
    static constexpr int NQ = 1; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) { q[0] = par[0]; }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        return fma(5.0, x[0], sin(x[0])) - q[0] * exp_bf(x[0]);
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int, double x, const double *par) { return fma(5.0, x, sin(x)) - par[0] * exp_bf(x); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GLTEST, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_GENZ_OSC, D> {
    static constexpr int NQ = D + 1; static constexpr bool kProd = false;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
        q[0] = IB_TWO_PI * par[D];
#pragma unroll
        for (int i = 0; i < D; i++) q[1 + i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = q[0];
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(q[1 + i], x[i], s);
        return cos_bf(s);
    }
    __device__ static double sep_init(const double *par) { return IB_TWO_PI * par[D]; }
    __device__ static double sep_term(int k, double x, const double *par) { return par[k] * x; }
    __device__ static double sep_finish(double acc, const double *) { return cos_bf(acc); }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        cos_bfN<N>(acc, out);
    }
};

template <int D> struct Family<IB200_FAM_GENZ_PROD, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < D; i++) { q[i] = 1.0 / (par[i] * par[i]); q[D + i] = par[D + i]; }
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double pr = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) { const double t = x[i] - q[D + i]; pr *= fma(t, t, q[i]); }
        return 1.0 / pr;
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int k, double x, const double *par) {
        const double t = x - par[D + k];
        return 1.0 / fma(t, t, 1.0 / (par[k] * par[k]));
    }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GENZ_PROD, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_GENZ_CORN, D> {
    static constexpr int NQ = D; static constexpr bool kProd = false;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < D; i++) q[i] = par[i];
    }
    __device__ static double pw(double s) {
        const double r = 1.0 / s; double p = r;
#pragma unroll
        for (int i = 0; i < D; i++) p *= r;
        return p;
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(q[i], x[i], s);
        return pw(s);
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int k, double x, const double *par) { return par[k] * x; }
    __device__ static double sep_finish(double acc, const double *) { return pw(acc); }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GENZ_CORN, D>, N>(acc, par, out);
    }
};

// Separable form of the three exponential families: exp(sum_k t_k) = prod_k exp(t_k).  The rule kernels tabulate the
// per-axis FACTORS exp(t_k) at the few abscissae a rule uses on each axis (7 for Genz-Malik, n for a tensor rule) and
// form every point as a product of table entries: 7 D exponentials per Genz-Malik box instead of 2^D + 2 D^2 + 2 D + 1
// (SURVEY 8(d) strength-reduction caveat: the integrand is still taken at every point of the reference's rule; the
// value of a point differs from exp-of-the-sum by a few ulp).  eval() keeps the reference's form for the kernels that
// evaluate points one by one (QuadGK, explicit node lists).
template <int D> struct Family<IB200_FAM_GENZ_GAUSS, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < 2 * D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { const double t = q[i] * (x[i] - q[D + i]); s = fma(t, t, s); }
        return exp_bf(-s);
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int k, double x, const double *par) { const double t = par[k] * (x - par[D + k]); return exp_bf(-(t * t)); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GENZ_GAUSS, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_GENZ_C0, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < 2 * D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(q[i], fabs(x[i] - q[D + i]), s);
        return exp_bf(-s);
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int k, double x, const double *par) { return exp_bf(-(par[k] * fabs(x - par[D + k]))); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GENZ_C0, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_GENZ_DISC, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < 2 * D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        bool zero = x[0] > q[D];
        if (D > 1) zero = zero || (x[D > 1 ? 1 : 0] > q[D + (D > 1 ? 1 : 0)]);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(q[i], x[i], s);
        return zero ? 0.0 : exp_bf(s);
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int k, double x, const double *par) {
        return (k < 2 && x > par[D + k]) ? 0.0 : exp_bf(par[k] * x);
    }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GENZ_DISC, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_COSPROD, D> {
    static constexpr int NQ = D; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double pr = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) pr *= cos(q[i] * x[i]);
        return pr;
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int k, double x, const double *par) { return cos(par[k] * x); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_COSPROD, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_GAUSSPOLY, D> {
    static constexpr int NQ = 3 * D + 1; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < 3 * D + 1; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 0.0, pr = q[3 * D];
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double t = q[i] * (x[i] - q[D + i]);
            s = fma(t, t, s);
            const int ki = (int)q[2 * D + i];
            for (int j = 0; j < ki; j++) pr *= x[i];
        }
        return pr * exp_bf(-s);
    }
    __device__ static double sep_init(const double *par) { return par[3 * D]; }
    __device__ static double sep_term(int k, double x, const double *par) {
        const double t = par[k] * (x - par[D + k]);
        double pr = 1.0;
        const int ki = (int)par[2 * D + k];
        for (int j = 0; j < ki; j++) pr *= x;
        return pr * exp_bf(-(t * t));
    }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        finishN_default<Family<IB200_FAM_GAUSSPOLY, D>, N>(acc, par, out);
    }
};

template <int D> struct Family<IB200_FAM_EXPLIN, D> {
    static constexpr int NQ = D; static constexpr bool kProd = false;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(q[i], x[i], s);
        return exp_bf(s);
    }
    __device__ static double sep_init(const double *) { return 0.0; }
    __device__ static double sep_term(int k, double x, const double *par) { return par[k] * x; }
    __device__ static double sep_finish(double acc, const double *) { return exp_bf(acc); }
    template <int N> __device__ static void sep_finishN(const double (&acc)[N], const double *par, double (&out)[N]) {
        exp_bfN<N>(acc, out);
    }
};

// ------------------------------------------------------------------------------------------------
// The separable terms of one axis at the 7 abscissae of the Genz-Malik rule, evaluated together.  Default: 7 calls of
// sep_term.  Product peak: the 7 reciprocals 1 / (t_j^2 + a^-2) share ONE division (prefix products, one reciprocal,
// back-substitution: 18 multiplies) -- ncu showed 33 divisions per box, each a MUFU.RCP64H + 8 FP64 instructions + a
// guarded slow path, 29 % of the kernel's instructions (profiles/r01b_pair_genz_product_peak_*).  Exponential families:
// the 7 exponentials run in lock step (independent polynomial chains side by side).
// ------------------------------------------------------------------------------------------------
template <int FAM, int D> struct Terms7 {
    __device__ __forceinline__ static void run(int k, const double (&u)[7], const double *par, double (&T)[7]) {
#pragma unroll
        for (int j = 0; j < 7; j++) T[j] = Family<FAM, D>::sep_term(k, u[j], par);
    }
};
template <int D> struct Terms7<IB200_FAM_GENZ_PROD, D> {
    __device__ __forceinline__ static void run(int k, const double (&u)[7], const double *par, double (&T)[7]) {
        const double ainv2 = 1.0 / (par[k] * par[k]);
        double d[7], pre[7];
        bool safe = ainv2 > 1e-40 && ainv2 < 1e40;           // products of 7 denominators stay far from over/underflow
#pragma unroll
        for (int j = 0; j < 7; j++) { const double t = u[j] - par[D + k]; d[j] = fma(t, t, ainv2); safe = safe && d[j] < 1e40; }
        if (safe) {
            pre[0] = d[0];
#pragma unroll
            for (int j = 1; j < 7; j++) pre[j] = pre[j - 1] * d[j];
            double inv = 1.0 / pre[6];                        // 1 / (d0 ... d6)
#pragma unroll
            for (int j = 6; j >= 1; j--) { T[j] = inv * pre[j - 1]; inv = inv * d[j]; }
            T[0] = inv;
        } else {
#pragma unroll
            for (int j = 0; j < 7; j++) T[j] = 1.0 / d[j];
        }
    }
};
template <int D> struct Terms7<IB200_FAM_GENZ_GAUSS, D> {
    __device__ __forceinline__ static void run(int k, const double (&u)[7], const double *par, double (&T)[7]) {
        double x[7];
#pragma unroll
        for (int j = 0; j < 7; j++) { const double t = par[k] * (u[j] - par[D + k]); x[j] = -(t * t); }
        exp_bfN<7>(x, T);
    }
};
template <int D> struct Terms7<IB200_FAM_GENZ_C0, D> {
    __device__ __forceinline__ static void run(int k, const double (&u)[7], const double *par, double (&T)[7]) {
        double x[7];
#pragma unroll
        for (int j = 0; j < 7; j++) x[j] = -(par[k] * fabs(u[j] - par[D + k]));
        exp_bfN<7>(x, T);
    }
};

// ------------------------------------------------------------------------------------------------
// Transformed integrand g(t) = f(u(t), p) * prod_k jac_k(t_k)   (infinity_handling.jl:11-25).
// ALLFIN: every axis finite -> u = lb + (1+t)*den, jac = prod den (precomputed constant jacc).
// ------------------------------------------------------------------------------------------------
template <int FAM, int D, bool ALLFIN>
struct Problem {
    using F = Family<FAM, D>;
    double q[F::NQ];
    double lb[D], den[D];       // ALLFIN path
    double jacc;
    AxisMap ax[ALLFIN ? 1 : D]; // general path
    int tr;

    __device__ void load(const double *par, const double *lbp, const double *ubp, int tr_) {
        F::prepare(par, q);
        tr = tr_;
        jacc = 1.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (ALLFIN) {
                lb[k] = lbp[k];
                den[k] = (ubp[k] - lbp[k]) * 0.5;
                jacc = (k == 0) ? den[k] : jacc * den[k];
            } else {
                make_axis(lbp[k], ubp[k], ax[k]);
            }
        }
    }
    // rule-domain affine map of quadrules.jl:2-3 on the transformed domain (tlo, thi)
    __device__ __forceinline__ double tscale(int k) const { return ALLFIN ? 1.0 : (ax[ALLFIN ? 0 : k].thi - ax[ALLFIN ? 0 : k].tlo) / 2; }
    __device__ __forceinline__ double tshift(int k) const { return ALLFIN ? 0.0 : (ax[ALLFIN ? 0 : k].tlo + ax[ALLFIN ? 0 : k].thi) / 2; }
    // g(t)
    __device__ __forceinline__ double operator()(const double (&t)[D]) const {
        double u[D];
        if (ALLFIN) {
#pragma unroll
            for (int k = 0; k < D; k++) u[k] = lb[k] + (1.0 + t[k]) * den[k];
            return F::eval(u, q) * jacc;
        } else {
            double jac = 1.0;
#pragma unroll
            for (int k = 0; k < D; k++) {
                double jk;
                t2ujac(tr, t[k], ax[k], u[k], jk);
                jac = (k == 0) ? jk : jac * jk;
            }
            return F::eval(u, q) * jac;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Parameter sensitivities: value and gradient of a family with respect to its RAW parameters at one point,
// out = [f, df/dpar[0], ..., df/dpar[nparam-1]].  The reference differentiates solve() through the integrand -- ForwardDiff duals
// flowing through the quadrature (ext/IntegralsForwardDiffExt.jl:55-134: "a vector-valued integral of the primal and dual
// simultaneously") or a second, vector-valued integral of df/dp (ext/IntegralsZygoteExt.jl:144-178); here the derivative of a registered
// family is written out and integrated as extra components by the vector-valued kernels (ib200_quadgk_sens_batch,
// ib200_hcubature_sens_batch).  Oracle: orc_family_grad.  The discontinuous family has no entry: the integral of its pointwise
// derivative is not the derivative of its integral (the support moves with u).
// ------------------------------------------------------------------------------------------------
struct SensIdx { int n; int idx[7]; };       // n > 0: sensitivity mode of the vector-valued kernels, components [f, df/dpar[idx[0]], ...]
template <int FAM, int D> struct Grad {
    static constexpr bool ok = false;
    __device__ static void run(const double (&)[D], const double *, double *) {}
};
template <int D> struct Grad<IB200_FAM_SINXP, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double sn, cs; sincos(x[0] * par[0], &sn, &cs);
        out[0] = sn; out[1] = x[0] * cs;
    }
};
template <int D> struct Grad<IB200_FAM_GLTEST, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        const double e = exp_bf(x[0]);
        out[0] = fma(5.0, x[0], sin(x[0])) - par[0] * e; out[1] = -e;
    }
};
template <int D> struct Grad<IB200_FAM_GENZ_OSC, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double s = IB_TWO_PI * par[D];
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(par[i], x[i], s);
        double sn, cs; sincos(s, &sn, &cs);
        out[0] = cs;
#pragma unroll
        for (int i = 0; i < D; i++) { out[1 + i] = -(x[i] * sn); out[1 + D + i] = 0.0; }
        out[1 + D] = -(IB_TWO_PI * sn);
    }
};
template <int D> struct Grad<IB200_FAM_GENZ_PROD, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double T[D], f = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) { const double t = x[i] - par[D + i]; T[i] = 1.0 / fma(t, t, 1.0 / (par[i] * par[i])); f *= T[i]; }
        out[0] = f;
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double t = x[i] - par[D + i], a = par[i];
            out[1 + i] = f * (2.0 * T[i]) / (a * a * a);
            out[1 + D + i] = f * (2.0 * t * T[i]);
        }
    }
};
template <int D> struct Grad<IB200_FAM_GENZ_CORN, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double s = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(par[i], x[i], s);
        const double r = 1.0 / s; double f = r;
#pragma unroll
        for (int i = 0; i < D; i++) f *= r;
        out[0] = f;
        const double g = -(double)(D + 1) * f * r;
#pragma unroll
        for (int i = 0; i < D; i++) { out[1 + i] = g * x[i]; out[1 + D + i] = 0.0; }
    }
};
template <int D> struct Grad<IB200_FAM_GENZ_GAUSS, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { const double t = par[i] * (x[i] - par[D + i]); s = fma(t, t, s); }
        const double f = exp_bf(-s);
        out[0] = f;
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double d = x[i] - par[D + i], a = par[i];
            out[1 + i] = -2.0 * a * d * d * f;
            out[1 + D + i] = 2.0 * a * a * d * f;
        }
    }
};
template <int D> struct Grad<IB200_FAM_GENZ_C0, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(par[i], fabs(x[i] - par[D + i]), s);
        const double f = exp_bf(-s);
        out[0] = f;
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double d = x[i] - par[D + i];
            out[1 + i] = -fabs(d) * f;
            out[1 + D + i] = (d > 0.0 ? par[i] : (d < 0.0 ? -par[i] : 0.0)) * f;
        }
    }
};
template <int D> struct Grad<IB200_FAM_COSPROD, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double sn[D], cs[D], f = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) { sincos(par[i] * x[i], &sn[i], &cs[i]); f *= cs[i]; }
        out[0] = f;
#pragma unroll
        for (int i = 0; i < D; i++) {
            double o = -(x[i] * sn[i]);
#pragma unroll
            for (int k = 0; k < D; k++) if (k != i) o *= cs[k];
            out[1 + i] = o;
        }
    }
};
template <int D> struct Grad<IB200_FAM_GAUSSPOLY, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double s = 0.0, pr = 1.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double t = par[i] * (x[i] - par[D + i]);
            s = fma(t, t, s);
            const int ki = (int)par[2 * D + i];
            for (int j = 0; j < ki; j++) pr *= x[i];
        }
        const double g = pr * exp_bf(-s), f = par[3 * D] * g;
        out[0] = f;
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double d = x[i] - par[D + i], a = par[i];
            out[1 + i] = -2.0 * a * d * d * f;
            out[1 + D + i] = 2.0 * a * a * d * f;
            out[1 + 2 * D + i] = 0.0;                 // the monomial exponents are integers: no derivative
        }
        out[1 + 3 * D] = g;
    }
};
template <int D> struct Grad<IB200_FAM_EXPLIN, D> {
    static constexpr bool ok = true;
    __device__ static void run(const double (&x)[D], const double *par, double *out) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(par[i], x[i], s);
        const double f = exp_bf(s);
        out[0] = f;
#pragma unroll
        for (int i = 0; i < D; i++) out[1 + i] = x[i] * f;
    }
};

}  // namespace ib
