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

template <int D> struct Family<IB200_FAM_SINXP, D> {
    static constexpr int NQ = 1; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) { q[0] = par[0]; }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) { return sin(x[0] * q[0]); }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int, double x, const double *par) { return sin(x * par[0]); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
};

template <int D> struct Family<IB200_FAM_GLTEST, D> {
    static constexpr int NQ = 1; static constexpr bool kProd = true;
    __device__ static void prepare(const double *par, double (&q)[NQ]) { q[0] = par[0]; }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        return fma(5.0, x[0], sin(x[0])) - q[0] * exp(x[0]);
    }
    __device__ static double sep_init(const double *) { return 1.0; }
    __device__ static double sep_term(int, double x, const double *par) { return fma(5.0, x, sin(x)) - par[0] * exp(x); }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
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
        return cos(s);
    }
    __device__ static double sep_init(const double *par) { return IB_TWO_PI * par[D]; }
    __device__ static double sep_term(int k, double x, const double *par) { return par[k] * x; }
    __device__ static double sep_finish(double acc, const double *) { return cos(acc); }
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
};

template <int D> struct Family<IB200_FAM_GENZ_GAUSS, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = false;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < 2 * D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { const double t = q[i] * (x[i] - q[D + i]); s = fma(t, t, s); }
        return exp(-s);
    }
    __device__ static double sep_init(const double *) { return 0.0; }
    __device__ static double sep_term(int k, double x, const double *par) { const double t = par[k] * (x - par[D + k]); return t * t; }
    __device__ static double sep_finish(double acc, const double *) { return exp(-acc); }
};

template <int D> struct Family<IB200_FAM_GENZ_C0, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = false;
    __device__ static void prepare(const double *par, double (&q)[NQ]) {
#pragma unroll
        for (int i = 0; i < 2 * D; i++) q[i] = par[i];
    }
    __device__ static double eval(const double (&x)[D], const double (&q)[NQ]) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) s = fma(q[i], fabs(x[i] - q[D + i]), s);
        return exp(-s);
    }
    __device__ static double sep_init(const double *) { return 0.0; }
    __device__ static double sep_term(int k, double x, const double *par) { return par[k] * fabs(x - par[D + k]); }
    __device__ static double sep_finish(double acc, const double *) { return exp(-acc); }
};

template <int D> struct Family<IB200_FAM_GENZ_DISC, D> {
    static constexpr int NQ = 2 * D; static constexpr bool kProd = false;
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
        return zero ? 0.0 : exp(s);
    }
    __device__ static double sep_init(const double *) { return 0.0; }
    __device__ static double sep_term(int k, double x, const double *par) {
        return (k < 2 && x > par[D + k]) ? -INFINITY : par[k] * x;   // exp(-inf) == 0
    }
    __device__ static double sep_finish(double acc, const double *) { return exp(acc); }
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
        return pr * exp(-s);
    }
    __device__ static double sep_init(const double *par) { return par[3 * D]; }
    __device__ static double sep_term(int k, double x, const double *par) {
        const double t = par[k] * (x - par[D + k]);
        double pr = 1.0;
        const int ki = (int)par[2 * D + k];
        for (int j = 0; j < ki; j++) pr *= x;
        return pr * exp(-(t * t));
    }
    __device__ static double sep_finish(double acc, const double *) { return acc; }
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
        return exp(s);
    }
    __device__ static double sep_init(const double *) { return 0.0; }
    __device__ static double sep_term(int k, double x, const double *par) { return par[k] * x; }
    __device__ static double sep_finish(double acc, const double *) { return exp(acc); }
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

}  // namespace ib
