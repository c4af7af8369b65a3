// gm_rule_thread2.cuh -- one application of the Genz-Malik rule (degree 7, embedded degree 5) to a box by ONE thread, finite domains,
// second version: the rule of hcubature_pair.cuh gm_rule_thread with the work per point cut down.  Used by the round-based batch
// driver (hcubature_rb.cuh); the reference-order drivers keep the first version.
//
// Every registered family is separable, f(x) = finish(init (+|*) term_1(x_1) ... (+|*) term_d(x_d)), and the rule only uses 7 abscissae
// per axis, so a point's value is a combination of d table entries (hcubature_pair.cuh).  Here the combinations shared between points
// are formed once:
//   * centre and axis points:  O_i = combination of the centre terms of all axes but i (prefix / suffix), a point = O_i (+|*) its own term
//   * pair points (i, j):      O_ij likewise; the four sign combinations of (T_i, T_j) are formed once and combined with O_ij
//   * corner points:           the 2^(d/2) sign combinations of the low and of the high axes are tabulated, a corner = one of each
//   * the constant Jacobian of the finite map is factored out of the rule sums (the fourth differences are scaled by |jac|)
//   * product peak: the 7 reciprocals of an axis share one Newton reciprocal (MUFU seed + 2 FMA steps, no slow-path branch) through a
//     product TREE (depth 3 up, 3 down instead of 6 + 6 dependent multiplies); 1 / a_k^2 is prepared once per problem
//   * rule constants come from constant memory (an FP64 instruction cannot carry a 64-bit immediate: every literal costs two UMOVs)
// The integrand is still formed at every point of the reference's rule and enters the same sums (f1 .. f5, fourth differences); a
// point's value differs from the oracle's expression by rounding (association order), like the first version's.
// 4-D product peak: 496 -> ~390 executed FP64 instructions per box.
#pragma once
#include "hcubature_impl.cuh"

namespace g2 {

// [D][..]: l2, l3, l5, w1..w5, w1p..w4p, 10^D
#define G2_ROW(D) { 0.35856858280031806, 0.9486832980505138, 0.6882472016116853, \
    (double)(1 << D) * ((12824 - 9120 * D + 400 * D * D) / 19683.0), (double)(1 << D) * (980 / 6561.0), \
    (double)(1 << D) * ((1820 - 400 * D) / 19683.0), (double)(1 << D) * (200 / 19683.0), 6859 / 19683.0, \
    (double)(1 << D) * ((729 - 950 * D + 50 * D * D) / 729.0), (double)(1 << D) * (245 / 486.0), \
    (double)(1 << D) * ((265 - 100 * D) / 1458.0), (double)(1 << D) * (25 / 729.0), \
    (D == 2 ? 1e2 : D == 3 ? 1e3 : D == 4 ? 1e4 : D == 5 ? 1e5 : D == 6 ? 1e6 : D == 7 ? 1e7 : 1e8), 0.5, 12.0, 7.0 }
__constant__ double c_rule[9][16] = { {0}, {0}, G2_ROW(2), G2_ROW(3), G2_ROW(4), G2_ROW(5), G2_ROW(6), G2_ROW(7), G2_ROW(8) };
#undef G2_ROW
enum { L2 = 0, L3, L5, W1, W2, W3, W4, W5, W1P, W2P, W3P, W4P, P10, HALF, TWELVE, SEVEN };

// 1 / x for a positive normal x well inside the exponent range: MUFU seed + 2 Newton steps (<= 1 ulp), branch-free
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// per-problem prepared values (the kernel calls prepare once per problem; prep has D doubles)
template <int FAM, int D> struct Prep {
    __device__ static void run(const double *par, double *prep) {
        for (int k = 0; k < D; k++) prep[k] = 0.0;
    }
};
template <int D> struct Prep<IB200_FAM_GENZ_PROD, D> {
    __device__ static void run(const double *par, double *prep) {
        for (int k = 0; k < D; k++) prep[k] = 1.0 / (par[k] * par[k]);       // a_k^-2 exactly as Family::sep_term forms it
    }
};

// the separable terms of axis k at the rule's 7 abscissae (centre, +-l2 h, +-l3 h, +-l5 h)
template <int FAM, int D> struct Terms7b {
    __device__ __forceinline__ static void run(int k, const double (&u)[7], const double *par, const double *prep, double (&T)[7]) {
        ib::Terms7<FAM, D>::run(k, u, par, T);
    }
};
template <int D> struct Terms7b<IB200_FAM_GENZ_PROD, D> {
    __device__ __forceinline__ static void run(int k, const double (&u)[7], const double *par, const double *prep, double (&T)[7]) {
        const double ainv2 = prep[k], u0 = par[D + k];
        double d[7];
#pragma unroll
        for (int j = 0; j < 7; j++) { const double t = u[j] - u0; d[j] = fma(t, t, ainv2); }
        // d_j >= a^-2 > 0; the largest are the +-l3 abscissae.  Products of 7 such numbers stay far from over/underflow when
        // a^-2 and the largest d are moderate (always, for a box of a unit-scale domain)
        const bool safe = ainv2 > 1e-40 && fmax(d[3], d[4]) < 1e40;
        if (safe) {
            const double p01 = d[0] * d[1], p23 = d[2] * d[3], p45 = d[4] * d[5];
            const double p0123 = p01 * p23, p456 = p45 * d[6];
            const double inv = rcp_nr(p0123 * p456);                         // 1 / (d0 ... d6)
            const double i0123 = inv * p456, i456 = inv * p0123;             // 1/(d0 d1 d2 d3), 1/(d4 d5 d6)
            const double i01 = i0123 * p23, i23 = i0123 * p01, i45 = i456 * d[6];
            T[6] = i456 * p45;
            T[0] = i01 * d[1]; T[1] = i01 * d[0];
            T[2] = i23 * d[3]; T[3] = i23 * d[2];
            T[4] = i45 * d[5]; T[5] = i45 * d[4];
        } else {
#pragma unroll
            for (int j = 0; j < 7; j++) T[j] = 1.0 / d[j];
        }
    }
};

// pc = par[NPAR] | lbs[D] | den[D] | finit | jacc | prep[D]   (hcubature_rb.cuh's per-warp row; ALLFIN only)
template <int FAM, int D>
__device__ __forceinline__ void gm_rule_thread2(const double *pc, const double (&a)[D], const double (&b)[D],
                                                double &Iout, double &Eout, int &kdiv) {
    using F = ib::Family<FAM, D>;
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    constexpr bool P = F::kProd;
    const double *par = pc, *lbs = pc + NPAR, *den = lbs + D, *prep = den + D + 2;
    const double finit = den[D], jacc = den[D + 1];
    const double *cr = c_rule[D];
    auto op = [](double x, double y) { return P ? x * y : x + y; };

    double h[D], V = 1.0;
    double T0[D], T2p[D], T2m[D], T3p[D], T3m[D], T5p[D], T5m[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double c = cr[HALF] * (a[k] + b[k]);
        h[k] = cr[HALF] * fabs(b[k] - a[k]);
        V = (k == 0) ? h[0] : V * h[k];
        const double p2 = h[k] * cr[L2], p3 = h[k] * cr[L3], p5 = h[k] * cr[L5];
        const double t7[7] = {c, c + p2, c - p2, c + p3, c - p3, c + p5, c - p5};
        double u7[7], T7[7];
#pragma unroll
        for (int j = 0; j < 7; j++) u7[j] = lbs[k] + (1.0 + t7[j]) * den[k];
        Terms7b<FAM, D>::run(k, u7, par, prep, T7);
        T0[k] = T7[0]; T2p[k] = T7[1]; T2m[k] = T7[2]; T3p[k] = T7[3]; T3m[k] = T7[4]; T5p[k] = T7[5]; T5m[k] = T7[6];
    }
    auto finish4 = [&](const double (&acc)[4], double (&f)[4]) { F::template sep_finishN<4>(acc, par, f); };

    // O[i] = init (+|*) centre terms of every axis but i; the centre itself
    double O[D];
    {
        double pre[D], suf[D];                  // pre[i] = init (+|*) T0[0..i-1], suf[i] = T0[i+1..D-1]
        pre[0] = finit;
#pragma unroll
        for (int i = 1; i < D; i++) pre[i] = op(pre[i - 1], T0[i - 1]);
        suf[D - 1] = 0.0;
        if (D >= 2) suf[D - 2] = T0[D - 1];
#pragma unroll
        for (int i = D - 3; i >= 0; i--) suf[i] = op(T0[i + 1], suf[i + 1]);
#pragma unroll
        for (int i = 0; i < D; i++) O[i] = (i == D - 1) ? pre[i] : op(pre[i], suf[i]);
    }
    const double f1 = F::sep_finish(op(O[0], T0[0]), par);
    const double twelvef1 = cr[TWELVE] * f1;

    double divdiff[D];
    double f2 = 0.0, f3 = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        const double acc[4] = {op(O[i], T2p[i]), op(O[i], T2m[i]), op(O[i], T3p[i]), op(O[i], T3m[i])};
        double f[4];
        finish4(acc, f);
        const double f2i = f[0] + f[1];
        const double f3i = f[2] + f[3];
        f2 += f2i; f3 += f3i;
        divdiff[i] = fabs(f3i + twelvef1 - cr[SEVEN] * f2i);
    }
    double f4 = 0.0;
#pragma unroll
    for (int i = 0; i < D - 1; i++) {
#pragma unroll
        for (int j = i + 1; j < D; j++) {
            double Oij = finit;                 // init (+|*) centre terms of the axes other than i, j
#pragma unroll
            for (int k = 0; k < D; k++) if (k != i && k != j) Oij = op(Oij, T0[k]);
            const double acc[4] = {op(Oij, op(T3p[i], T3p[j])), op(Oij, op(T3p[i], T3m[j])), op(Oij, op(T3m[i], T3p[j])), op(Oij, op(T3m[i], T3m[j]))};
            double f[4];
            finish4(acc, f);
#pragma unroll
            for (int s = 0; s < 4; s++) f4 += f[s];
        }
    }
    double f5 = 0.0;
    {
        constexpr int DL = D / 2, DH = D - DL, NL = 1 << DL, NH = 1 << DH;
        double CL[NL], CH[NH];                  // sign combinations of the low axes; init (+|*) those of the high axes
#pragma unroll
        for (int c = 0; c < NL; c++) {
            double v = (c & 1) ? T5m[0] : T5p[0];
#pragma unroll
            for (int k = 1; k < DL; k++) v = op(v, ((c >> k) & 1) ? T5m[k] : T5p[k]);
            CL[c] = v;
        }
#pragma unroll
        for (int c = 0; c < NH; c++) {
            double v = finit;
#pragma unroll
            for (int k = 0; k < DH; k++) v = op(v, ((c >> k) & 1) ? T5m[DL + k] : T5p[DL + k]);
            CH[c] = v;
        }
#pragma unroll
        for (int s0 = 0; s0 < (1 << D); s0 += 4) {
            double acc[4], f[4];
#pragma unroll
            for (int s1 = 0; s1 < 4; s1++) { const int sb = s0 + s1; acc[s1] = op(CH[sb >> DL], CL[sb & (NL - 1)]); }
            finish4(acc, f);
#pragma unroll
            for (int s1 = 0; s1 < 4; s1++) f5 += f[s1];
        }
    }
    const double Vj = V * jacc;                 // the constant Jacobian of the finite map, factored out of the sums
    const double Iv = Vj * (cr[W1] * f1 + cr[W2] * f2 + cr[W3] * f3 + cr[W4] * f4 + cr[W5] * f5);
    const double Ip = Vj * (cr[W1P] * f1 + cr[W2P] * f2 + cr[W3P] * f3 + cr[W4P] * f4);
    const double Er = fabs(Iv - Ip);
    const double deltaf = Er / (cr[P10] * V);
    const double aj = fabs(jacc);
    int kdv = 0; double maxdd = 0.0, hkd = h[0];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double dd = divdiff[k] * aj;      // fourth differences of g = f * jac
        const double delta = dd - maxdd;
        if (delta > deltaf) { kdv = k; maxdd = dd; hkd = h[k]; }
        else if (fabs(delta) <= deltaf && h[k] > hkd) { kdv = k; hkd = h[k]; }
    }
    Iout = Iv; Eout = Er; kdiv = kdv;
}

}  // namespace g2
