// gm_rule_thread3.cuh -- one application of the Genz-Malik rule (degree 7, embedded degree 5) to a box by ONE thread in 6 .. 8
// dimensions (401 points in 8-D), finite AND infinite axes.  Used by the round-based batch driver (hcubature_rb.cuh) for the product
// families, where it replaces the warp-per-box rule (hcub_rounds.cuh WarpRule): that one spreads the points of a box over the lanes and
// pays for it in shared-memory table traffic, idle lanes in the table phase (7 d entries on 32 lanes) and a serial tail on lane 0
// (ncu, 8-D Gaussian: issue slots 72 %, FP64 pipe 38 %, LSU 46 %).  Here a lane owns a box, the whole rule runs in its registers and
// 32 boxes advance per warp instruction.
//
// Same decomposition as gm_rule_thread2.cuh (shared sub-combinations: centre terms of all axes but one / but two, half tables of the
// corner signs), organised so that few values are live at a time -- 8-D needs 7 x 8 separable terms per box, which do not fit the
// register file next to the box itself:
//   sweep 1   centre terms T0[k] of every axis (in lock step: one exponential chain per axis side by side)
//   sweep 2   per axis i: the six off-centre terms; the +-l2 / +-l3 axis points are formed and summed at once with O_i = init (+|*)
//             T0[k != i]; only T(+-l3), T(+-l5) are kept; the fourth difference and the half-width go to the lane's shared scratch
//   pairs     O_ij (+|*) T_i(+-l3) (+|*) T_j(+-l3), four points per pair
//   corners   the 2^(d/2) sign combinations of the low axes tabulated, those of the high axes formed on the way (nested)
// Infinite axes (ALLFIN = false; product families only): the Jacobian factor of an abscissa is folded into its term, T := T * jac
// (finish is the identity for every product family), so a point is still ONE product of d entries.
// The integrand is formed at every point of the reference's rule and enters the same sums (f1 .. f5, fourth differences) as in
// HCubature.jl's GenzMalik rule (call site src/Integrals.jl:380-393); a point's value differs from the oracle's expression by rounding.
#pragma once
#include "gm_rule_thread2.cuh"

namespace g3 {

using g2::c_rule;
using g2::rcp_nr;

// separable terms of axis k at N abscissae, evaluated together
template <int FAM, int D, int N> struct TermsN {
    __device__ __forceinline__ static void run(int k, const double (&u)[N], const double *par, const double *prep, double (&T)[N]) {
#pragma unroll
        for (int j = 0; j < N; j++) T[j] = ib::Family<FAM, D>::sep_term(k, u[j], par);
    }
};
template <int D, int N> struct TermsN<IB200_FAM_GENZ_GAUSS, D, N> {
    __device__ __forceinline__ static void run(int k, const double (&u)[N], const double *par, const double *prep, double (&T)[N]) {
        double x[N];
#pragma unroll
        for (int j = 0; j < N; j++) { const double t = par[k] * (u[j] - par[D + k]); x[j] = -(t * t); }
        ib::exp_bfN<N, true>(x, T);
    }
};
template <int D, int N> struct TermsN<IB200_FAM_GENZ_C0, D, N> {
    __device__ __forceinline__ static void run(int k, const double (&u)[N], const double *par, const double *prep, double (&T)[N]) {
        double x[N];
#pragma unroll
        for (int j = 0; j < N; j++) x[j] = -(par[k] * fabs(u[j] - par[D + k]));
        ib::exp_bfN<N, true>(x, T);
    }
};
// product peak: the N reciprocals 1 / (t_j^2 + a^-2) share one Newton reciprocal (prefix products, back-substitution)
template <int D, int N> struct TermsN<IB200_FAM_GENZ_PROD, D, N> {
    __device__ __forceinline__ static void run(int k, const double (&u)[N], const double *par, const double *prep, double (&T)[N]) {
        const double ainv2 = prep[k], u0 = par[D + k];
        double d[N], dmax = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) { const double t = u[j] - u0; d[j] = fma(t, t, ainv2); dmax = fmax(dmax, d[j]); }
        if (ainv2 > 1e-40 && dmax < 1e40) {
            double pre[N];
            pre[0] = d[0];
#pragma unroll
            for (int j = 1; j < N; j++) pre[j] = pre[j - 1] * d[j];
            double inv = rcp_nr(pre[N - 1]);
#pragma unroll
            for (int j = N - 1; j >= 1; j--) { T[j] = inv * pre[j - 1]; inv = inv * d[j]; }
            T[0] = inv;
        } else {
#pragma unroll
            for (int j = 0; j < N; j++) T[j] = 1.0 / d[j];
        }
    }
};

// the centre terms of all D axes in lock step
template <int FAM, int D> struct Centres {
    __device__ __forceinline__ static void run(const double (&u)[D], const double *par, const double *prep, double (&T)[D]) {
#pragma unroll
        for (int k = 0; k < D; k++) T[k] = ib::Family<FAM, D>::sep_term(k, u[k], par);
    }
};
template <int D> struct Centres<IB200_FAM_GENZ_GAUSS, D> {
    __device__ __forceinline__ static void run(const double (&u)[D], const double *par, const double *prep, double (&T)[D]) {
        double x[D];
#pragma unroll
        for (int k = 0; k < D; k++) { const double t = par[k] * (u[k] - par[D + k]); x[k] = -(t * t); }
        ib::exp_bfN<D, true>(x, T);
    }
};
template <int D> struct Centres<IB200_FAM_GENZ_C0, D> {
    __device__ __forceinline__ static void run(const double (&u)[D], const double *par, const double *prep, double (&T)[D]) {
        double x[D];
#pragma unroll
        for (int k = 0; k < D; k++) x[k] = -(par[k] * fabs(u[k] - par[D + k]));
        ib::exp_bfN<D, true>(x, T);
    }
};
template <int D> struct Centres<IB200_FAM_GENZ_PROD, D> {
    __device__ __forceinline__ static void run(const double (&u)[D], const double *par, const double *prep, double (&T)[D]) {
#pragma unroll
        for (int k = 0; k < D; k++) { const double t = u[k] - par[D + k]; T[k] = rcp_nr(fma(t, t, prep[k])); }
    }
};

// t -> (u, du/dt) of family.cuh t2ujac with the rarely used maps (tan / cot) out of line: the rule below maps 7 d abscissae per box, and
// an inlined tan() at every one of them is what made the first version of this rule 23,784 SASS instructions (ncu: no_instruction 26
// cycles per issue -- the instruction cache, not the FP64 pipe, set the time)
static __device__ __noinline__ double2 t2ujac_trig(int tr, double t, int kind, double lb, double ub) {
    ib::AxisMap m; m.lb = lb; m.ub = ub; m.kind = kind; m.den = 0.0; m.sg = 1.0; m.tlo = 0.0; m.thi = 0.0;
    double u, jac;
    ib::t2ujac(tr, t, m, u, jac);
    return make_double2(u, jac);
}
// transformation_if_inf (src/infinity_handling.jl:112-120) with ONE branch-free reciprocal whatever the kind of infinite axis: kind 3
// (both bounds infinite) u = t / (1 - t^2), jac = (1 + t^2) / (1 - t^2)^2; kinds 1, 2: u = bound + t / (1 - t sg), jac = 1 / (1 - t sg)^2.
// Abscissae of the rule lie strictly inside (-1, 1), so 1 - q is a positive normal number and the Newton reciprocal is exact to <= 1 ulp.
__device__ __forceinline__ void t2ujac_lean(int tr, double t, const ib::AxisMap &m, double &u, double &jac) {
    if (m.kind == 0) { u = m.lb + (1.0 + t) * m.den; jac = m.den; }
    else if (tr == IB200_TR_IF_INF) {
        const bool both = m.kind == 3;
        const double q = both ? t * t : t * m.sg;
        const double den = rcp_nr(1.0 - q);
        const double base = both ? 0.0 : (m.kind == 1 ? m.ub : m.lb);
        u = base + t * den;
        jac = both ? (1.0 + q) * (den * den) : den * den;
    } else { const double2 r = t2ujac_trig(tr, t, m.kind, m.lb, m.ub); u = r.x; jac = r.y; }
}

// pc = par[NPAR] | lbs[D] | aux[D] (ALLFIN: half-widths den, else ub) | finit | jacc | prep[D]   (hcubature_rb.cuh's per-warp row)
// scr: this lane's scratch in shared memory, D entries of stride 32 (the lane's column of the record staging buffer, free while the
// rule runs): (centre, half-width) of every axis, then (fourth difference, half-width).
// Code size matters as much as the instruction count here: the loop over the axes of sweep 2 and the loop over the sign combinations of
// the high axes are ROLLED (the register arrays they touch are read and written through selects on the loop index), which keeps the
// kernel near 3,000 SASS instructions.
template <int FAM, int D, bool ALLFIN>
__device__ __forceinline__ void gm_rule_thread3(const double *pc, int tr, const double (&a)[D], const double (&b)[D], double2 *scr,
                                                double &Iout, double &Eout, int &kdiv) {
    using F = ib::Family<FAM, D>;
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    constexpr bool P = F::kProd;
    static_assert(ALLFIN || P, "infinite axes: product families only (the Jacobian is folded into the terms)");
    const double *par = pc, *lbs = pc + NPAR, *aux = lbs + D, *prep = aux + D + 2;
    const double finit = aux[D], jacc = aux[D + 1];
    const double *cr = c_rule[D];
    auto op = [](double x, double y) { return P ? x * y : x + y; };
    auto finish4 = [&](const double (&acc)[4], double (&f)[4]) { F::template sep_finishN<4>(acc, par, f); };
    constexpr double kNeutral = P ? 1.0 : 0.0;

    // ---- sweep 1: centres
    double T0[D], V = 1.0;
    {
        double u0[D], j0[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const double c = cr[g2::HALF] * (a[k] + b[k]);
            const double h = cr[g2::HALF] * fabs(b[k] - a[k]);
            scr[k * 32] = make_double2(c, h);
            V = (k == 0) ? h : V * h;
            if (ALLFIN) { u0[k] = lbs[k] + (1.0 + c) * aux[k]; j0[k] = 1.0; }
            else { ib::AxisMap m; ib::make_axis(lbs[k], aux[k], m); t2ujac_lean(tr, c, m, u0[k], j0[k]); }
        }
        Centres<FAM, D>::run(u0, par, prep, T0);
        if (!ALLFIN) {
#pragma unroll
            for (int k = 0; k < D; k++) T0[k] = T0[k] * j0[k];
        }
    }
    double f1;
    {
        double acc = finit;
#pragma unroll
        for (int k = 0; k < D; k++) acc = op(acc, T0[k]);
        f1 = F::sep_finish(acc, par);
    }
    const double twelvef1 = cr[g2::TWELVE] * f1;

    // ---- sweep 2 (rolled): per axis, the six off-centre terms; axis points
    double T3p[D], T3m[D], T5p[D], T5m[D];
#pragma unroll
    for (int k = 0; k < D; k++) { T3p[k] = 0.0; T3m[k] = 0.0; T5p[k] = 0.0; T5m[k] = 0.0; }
    double f2 = 0.0, f3 = 0.0;
#pragma unroll 1
    for (int i = 0; i < D; i++) {
        const double2 ch = scr[i * 32];
        const double c = ch.x, h = ch.y;
        const double p2 = h * cr[g2::L2], p3 = h * cr[g2::L3], p5 = h * cr[g2::L5];
        const double t6[6] = {c + p2, c - p2, c + p3, c - p3, c + p5, c - p5};
        double u6[6], j6[6], T6[6];
        if (ALLFIN) {
            const double l = lbs[i], dn = aux[i];
#pragma unroll
            for (int j = 0; j < 6; j++) { u6[j] = l + (1.0 + t6[j]) * dn; j6[j] = 1.0; }
        } else {
            ib::AxisMap m; ib::make_axis(lbs[i], aux[i], m);
#pragma unroll
            for (int j = 0; j < 6; j++) t2ujac_lean(tr, t6[j], m, u6[j], j6[j]);
        }
        TermsN<FAM, D, 6>::run(i, u6, par, prep, T6);
        if (!ALLFIN) {
#pragma unroll
            for (int j = 0; j < 6; j++) T6[j] = T6[j] * j6[j];
        }
        double Oi = finit;                      // init (+|*) centre terms of every axis but i
#pragma unroll
        for (int k = 0; k < D; k++) Oi = op(Oi, k == i ? kNeutral : T0[k]);
        const double acc[4] = {op(Oi, T6[0]), op(Oi, T6[1]), op(Oi, T6[2]), op(Oi, T6[3])};
        double f[4];
        finish4(acc, f);
        const double f2i = f[0] + f[1];
        const double f3i = f[2] + f[3];
        f2 += f2i; f3 += f3i;
        scr[i * 32] = make_double2(fabs(f3i + twelvef1 - cr[g2::SEVEN] * f2i), h);
#pragma unroll
        for (int k = 0; k < D; k++) if (k == i) { T3p[k] = T6[2]; T3m[k] = T6[3]; T5p[k] = T6[4]; T5m[k] = T6[5]; }
    }

    // ---- pair points: +-l3 on two axes.  O_ij = init (+|*) centre terms of the axes other than i, j = pre_i (+|*) T0[i+1 .. j-1] (+|*) suf_j
    double f4 = 0.0;
    {
        double suf[D];                          // suf[j] = T0[j+1] (+|*) ... (+|*) T0[D-1]
        suf[D - 1] = kNeutral;
#pragma unroll
        for (int j = D - 2; j >= 0; j--) suf[j] = op(T0[j + 1], suf[j + 1]);
        double pre = finit;                     // init (+|*) T0[0 .. i-1]
#pragma unroll
        for (int i = 0; i < D - 1; i++) {
            double left = pre;                  // pre (+|*) T0[i+1 .. j-1]
#pragma unroll
            for (int j = i + 1; j < D; j++) {
                const double Oij = op(left, suf[j]);
                const double Ap = op(Oij, T3p[i]), Am = op(Oij, T3m[i]);
                const double acc[4] = {op(Ap, T3p[j]), op(Ap, T3m[j]), op(Am, T3p[j]), op(Am, T3m[j])};
                double f[4];
                finish4(acc, f);
                f4 += (f[0] + f[1]) + (f[2] + f[3]);
                left = op(left, T0[j]);
            }
            pre = op(pre, T0[i]);
        }
    }

    // ---- corner points: +-l5 on every axis
    double f5 = 0.0;
    {
        constexpr int DL = D / 2, NL = 1 << DL, DH = D - DL;
        double CL[NL];                          // sign combinations of the low axes (tree: 2 + 4 + ... + 2^DL products)
        CL[0] = T5p[0]; CL[1] = T5m[0];
#pragma unroll
        for (int k = 1; k < DL; k++) {
#pragma unroll
            for (int c = (1 << k) - 1; c >= 0; c--) { const double v = CL[c]; CL[c + (1 << k)] = op(v, T5m[k]); CL[c] = op(v, T5p[k]); }
        }
        double acc5[4] = {0.0, 0.0, 0.0, 0.0};  // four partial sums (shorter dependent add chains)
#pragma unroll 1
        for (int chs = 0; chs < (1 << DH); chs++) {                 // sign combination of the high axes (rolled)
            double v = finit;
#pragma unroll
            for (int k = 0; k < DH; k++) v = op(v, ((chs >> k) & 1) ? T5m[DL + k] : T5p[DL + k]);
#pragma unroll
            for (int c0 = 0; c0 < NL; c0 += 4) {
                const double acc[4] = {op(v, CL[c0]), op(v, CL[c0 + 1]), op(v, CL[c0 + 2]), op(v, CL[c0 + 3])};
                double f[4];
                finish4(acc, f);
#pragma unroll
                for (int s = 0; s < 4; s++) acc5[s] += f[s];
            }
        }
        f5 = (acc5[0] + acc5[1]) + (acc5[2] + acc5[3]);
    }

    const double Vj = ALLFIN ? V * jacc : V;    // the constant Jacobian of the finite map, factored out of the sums
    const double Iv = Vj * (cr[g2::W1] * f1 + cr[g2::W2] * f2 + cr[g2::W3] * f3 + cr[g2::W4] * f4 + cr[g2::W5] * f5);
    const double Ip = Vj * (cr[g2::W1P] * f1 + cr[g2::W2P] * f2 + cr[g2::W3P] * f3 + cr[g2::W4P] * f4);
    const double Er = fabs(Iv - Ip);
    const double deltaf = Er / (cr[g2::P10] * V);
    const double aj = ALLFIN ? fabs(jacc) : 1.0;
    int kdv = 0; double maxdd = 0.0, hkd = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double2 dh = scr[k * 32];
        const double dd = dh.x * aj;            // fourth differences of g = f * jac
        if (k == 0) hkd = dh.y;
        const double delta = dd - maxdd;
        if (delta > deltaf) { kdv = k; maxdd = dd; hkd = dh.y; }
        else if (fabs(delta) <= deltaf && dh.y > hkd) { kdv = k; hkd = dh.y; }
    }
    Iout = Iv; Eout = Er; kdiv = kdv;
}

}  // namespace g3
