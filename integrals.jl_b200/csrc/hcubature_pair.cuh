// hcubature_pair.cuh -- batched adaptive Genz-Malik cubature, "pair mode": one THREAD per box,
// two lanes (the two children of a bisection) per problem, 16 problems per warp.
// With D == 1 the same machinery runs QuadGK's (7,15) Gauss-Kronrod bisection (ib200_quadgk_batch, large batches).
//
// Same algorithm and the same selection order as hcubature_impl.cuh (HCubature.jl as called at
// src/Integrals.jl:380-393; oracle: orc_hcubature / gm_rule): largest error first, ties -> lowest
// slot, I and E updated by difference, final re-sum.  What changes is the mapping onto the SM.
// ncu on the warp-per-problem kernel (profiles/r01_*) showed it instruction-issue bound: ~1,050 warp
// instructions per refinement step, 6x the FP64 work, spent on index arithmetic, shared-memory
// tables, shuffles and the selection scan.  Here
//   * a lane evaluates the whole 2^d + 2d^2 + 2d + 1 point rule of ITS box with a fully unrolled
//     point loop; the per-axis separable terms (family.cuh) of the 7 abscissae live in registers, so
//     one evaluation is d adds + the family's transcendental + one multiply + one add;
//   * the rule sums are formed in the oracle's order (f1; f2_i, f3_i per axis; pair points; corner
//     points; I = V*(w1 f1 + ... + w5 f5));
//   * the worst box is found by a three-level tournament with fan-out 32: error keys (global), group
//     maxima (global), maxima of 1024-box blocks (shared).  Every level also stores WHERE its maximum
//     sits, so the slot of the worst box falls out of the shared-memory scan alone and the only
//     global load in front of the rule evaluation is the 16(d+1)-byte box record.  The two 32-entry
//     groups on the selected path are staged into shared memory with cp.async while the rule runs and
//     scanned after it for their runner-ups; the maxima along the path are then repaired from registers;
//   * finished problems are re-summed by the whole warp (coalesced), then the pair pulls the next
//     problem from the global atomic counter -- in longest-expected-first order when a pilot pass ran
//     (lpt_bucket below, hcubature.cu);
//   * per-family rule specialisations: exponential families multiply per-axis factors, the 7 terms of
//     an axis are evaluated together (family.cuh Terms7), the oscillatory family uses angle addition
//     (gm_rule_thread_osc).
#pragma once
#include "hcubature_impl.cuh"

namespace hp {

constexpr int kBlock = 128;                 // 4 warps = 64 pairs
constexpr int kPairs = kBlock / 2;
constexpr int kFan = 32;

struct Params {
    const double *params; int nparam; int64_t nprob;
    const double *lb, *ub; int bpp, tr;
    double rtol, atol; int64_t maxevals; int initdiv;
    int cap;                    // boxes per pair, multiple of 1024
    double *keys;               // [pairs][cap]          error of every box
    double *l1;                 // [pairs][cap/32]       maxima of groups of 32 keys
    unsigned char *l1off;       // [pairs][cap/32]       position (0..31) of that maximum in its group
    double *rec;                // [pairs][cap][2D+2]    a[D], b[D], I, split axis
    unsigned long long *counter;
    int gk_order;               // D == 1 only: 0 = the built-in (7,15) pair, else the order n of the rule below
    const double *gk_x, *gk_w, *gk_wg;   // QuadGK.kronrod(n) layout: x[n+1], w[n+1], wg[(n+1)/2] (device)
    const unsigned *order;      // problem indices in processing order (NULL: 0, 1, 2, ...), see lpt_* below
    const unsigned long long *norder;   // device-side length of `order` (NULL: nprob)
    double *out_I, *out_E; int64_t *out_nevals; int32_t *out_status;
};

enum { ST_NEED = 0, ST_INIT, ST_REFINE, ST_FIN, ST_EXHAUSTED };

__device__ __forceinline__ unsigned long long dk(double e) { return (unsigned long long)__double_as_longlong(e); }
__device__ __forceinline__ double kd2d(unsigned long long k) { return __longlong_as_double((long long)k); }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

__device__ __forceinline__ unsigned long long shfl64(unsigned mask, unsigned long long v, int src) {
    return __shfl_sync(mask, v, src);
}

// per-pair shared-memory row: top keys [n2max] | top slots [n2max ints] | par[NPAR] | lbs[D] | aux[D] (ALLFIN: den, else ub) | finit | jacc
template <int FAM, int D> __host__ __device__ constexpr int pc_doubles() { return ib::fam_nparam(FAM, D) + 2 * D + 2; }
__host__ __device__ inline int row_doubles(int n2max, int pcd) { int r = n2max + (n2max + 1) / 2 + pcd; return (r & 1) ? r : r + 1; }

// ------------------------------------------------------------------------------------------------
// One application of the Genz-Malik rule (degree 7, embedded degree 5) to the box (a, b), by one thread.
// ------------------------------------------------------------------------------------------------
template <int FAM, int D, bool ALLFIN>
__device__ __forceinline__ void gm_rule_thread(const double *pc, int tr, const double (&a)[D], const double (&b)[D],
                                               double &Iout, double &Eout, int &kdiv) {
    using F = ib::Family<FAM, D>;
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    const double *par = pc, *lbs = pc + NPAR, *aux = lbs + D;
    const double finit = aux[D], jacc = aux[D + 1];
    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
    constexpr double twon = (double)(1 << D);
    constexpr double w1 = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0);
    constexpr double w2 = twon * (980 / 6561.0);
    constexpr double w3 = twon * ((1820 - 400 * D) / 19683.0);
    constexpr double w4 = twon * (200 / 19683.0);
    constexpr double w5 = 6859 / 19683.0;
    constexpr double w4p = twon * (25 / 729.0);
    constexpr double w3p = twon * ((265 - 100 * D) / 1458.0);
    constexpr double w2p = twon * (245 / 486.0);
    constexpr double w1p = twon * ((729 - 950 * D + 50 * D * D) / 729.0);

    double c[D], h[D], V = 1.0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        c[k] = 0.5 * (a[k] + b[k]);
        h[k] = 0.5 * fabs(b[k] - a[k]);
        V = (k == 0) ? h[0] : V * h[k];
    }
    // separable terms (and Jacobian factors) of the 7 abscissae of every axis: c, c +- l2 h, c +- l3 h, c +- l5 h
    // (family.cuh Terms7: one division per axis for the product peak, lock-step exponentials for the exp families)
    double T0[D], J0[D], T2p[D], T2m[D], J2p[D], J2m[D], T3p[D], T3m[D], J3p[D], J3m[D], T5p[D], T5m[D], J5p[D], J5m[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double p2 = h[k] * l2, p3 = h[k] * l3, p5 = h[k] * l5;
        const double t7[7] = {c[k], c[k] + p2, c[k] - p2, c[k] + p3, c[k] - p3, c[k] + p5, c[k] - p5};
        double u7[7], j7[7], T7[7];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            if (ALLFIN) { u7[j] = lbs[k] + (1.0 + t7[j]) * aux[k]; j7[j] = 1.0; }
            else { ib::AxisMap m; ib::make_axis(lbs[k], aux[k], m); ib::t2ujac(tr, t7[j], m, u7[j], j7[j]); }
        }
        ib::Terms7<FAM, D>::run(k, u7, par, T7);
        T0[k] = T7[0]; T2p[k] = T7[1]; T2m[k] = T7[2]; T3p[k] = T7[3]; T3m[k] = T7[4]; T5p[k] = T7[5]; T5m[k] = T7[6];
        J0[k] = j7[0]; J2p[k] = j7[1]; J2m[k] = j7[2]; J3p[k] = j7[3]; J3m[k] = j7[4]; J5p[k] = j7[5]; J5m[k] = j7[6];
    }
    // f = finish(init (+|*) T_{D-1} ... (+|*) T_0) * jac ; folding from the last axis down lets the
    // compiler share prefixes between consecutive points (the first axis toggles fastest).  The
    // transcendental is applied to four points at a time in lock step (family.cuh sep_finishN).
    auto fold = [&](const double (&T)[D], const double (&J)[D], double &acc, double &jac) {
        acc = finit; jac = 1.0;
#pragma unroll
        for (int k = D - 1; k >= 0; k--) {
            acc = F::kProd ? acc * T[k] : acc + T[k];
            if (!ALLFIN) jac = (k == D - 1) ? J[k] : jac * J[k];
        }
    };
    auto finish4 = [&](const double (&acc)[4], const double (&jac)[4], double (&f)[4]) {
        F::template sep_finishN<4>(acc, par, f);
#pragma unroll
        for (int j = 0; j < 4; j++) f[j] = f[j] * (ALLFIN ? jacc : jac[j]);
    };

    double f1;
    {
        double acc, jac;
        fold(T0, J0, acc, jac);
        f1 = F::sep_finish(acc, par) * (ALLFIN ? jacc : jac);
    }
    const double twelvef1 = 12 * f1;

    double divdiff[D];
    double f2 = 0.0, f3 = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double Tt[D], Jt[D], acc[4], jac[4], f[4];
#pragma unroll
        for (int k = 0; k < D; k++) { Tt[k] = T0[k]; Jt[k] = J0[k]; }
        Tt[i] = T2p[i]; Jt[i] = J2p[i];
        fold(Tt, Jt, acc[0], jac[0]);
        Tt[i] = T2m[i]; Jt[i] = J2m[i];
        fold(Tt, Jt, acc[1], jac[1]);
        Tt[i] = T3p[i]; Jt[i] = J3p[i];
        fold(Tt, Jt, acc[2], jac[2]);
        Tt[i] = T3m[i]; Jt[i] = J3m[i];
        fold(Tt, Jt, acc[3], jac[3]);
        finish4(acc, jac, f);
        const double f2i = f[0] + f[1];
        const double f3i = f[2] + f[3];
        f2 += f2i; f3 += f3i;
        divdiff[i] = fabs(f3i + twelvef1 - 7 * f2i);
    }
    double f4 = 0.0;
#pragma unroll
    for (int i = 0; i < D - 1; i++) {
#pragma unroll
        for (int j = i + 1; j < D; j++) {
            double acc[4], jac[4], f[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                double Tt[D], Jt[D];
#pragma unroll
                for (int k = 0; k < D; k++) { Tt[k] = T0[k]; Jt[k] = J0[k]; }
                Tt[i] = (s & 2) ? T3m[i] : T3p[i]; Jt[i] = (s & 2) ? J3m[i] : J3p[i];
                Tt[j] = (s & 1) ? T3m[j] : T3p[j]; Jt[j] = (s & 1) ? J3m[j] : J3p[j];
                fold(Tt, Jt, acc[s], jac[s]);
            }
            finish4(acc, jac, f);
#pragma unroll
            for (int s = 0; s < 4; s++) f4 += f[s];
        }
    }
    double f5 = 0.0;
#pragma unroll
    for (int s0 = 0; s0 < (1 << D); s0 += 4) {
        double acc[4], jac[4], f[4];
#pragma unroll
        for (int s1 = 0; s1 < 4; s1++) {
            const int sb = s0 + s1;
            double Tt[D], Jt[D];
#pragma unroll
            for (int k = 0; k < D; k++) { Tt[k] = ((sb >> k) & 1) ? T5m[k] : T5p[k]; Jt[k] = ((sb >> k) & 1) ? J5m[k] : J5p[k]; }
            fold(Tt, Jt, acc[s1], jac[s1]);
        }
        finish4(acc, jac, f);
#pragma unroll
        for (int s1 = 0; s1 < 4; s1++) f5 += f[s1];
    }
    const double Iv = V * (w1 * f1 + w2 * f2 + w3 * f3 + w4 * f4 + w5 * f5);
    const double Ip = V * (w1p * f1 + w2p * f2 + w3p * f3 + w4p * f4);
    const double Er = fabs(Iv - Ip);
    double p10 = 1.0;
#pragma unroll
    for (int k = 0; k < D; k++) p10 *= 10.0;
    const double deltaf = Er / (p10 * V);
    int kdv = 0; double maxdd = 0.0, hkd = h[0];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double delta = divdiff[k] - maxdd;
        if (delta > deltaf) { kdv = k; maxdd = divdiff[k]; hkd = h[k]; }
        else if (fabs(delta) <= deltaf && h[k] > hkd) { kdv = k; hkd = h[k]; }
    }
    Iout = Iv; Eout = Er; kdiv = kdv;
}

// ------------------------------------------------------------------------------------------------
// Genz-Malik rule for the oscillatory family cos(2 pi u_1 + sum a_k x_k) on a finite domain, by angle addition.
// Every point of the rule is the box centre moved by +-delta on at most two axes, or by +-delta_5 on all of them, so
// its phase is S_c + sum(+-theta_k) with theta_k = a_k * (lambda h_k) * den_k.  One sincos of S_c and of the 3 D
// angles theta_k(lambda_2, lambda_3, lambda_5) (13 in 4-D, evaluated four at a time in lock step) replaces the 2^D + 2 D^2
// + 2 D + 1 cosines (57): a point then costs 2-7 multiply-adds.  The integrand is still formed at every point and
// enters the sums of the reference's rule (f1 .. f5, fourth differences); values differ from cos-of-the-sum by ~1 ulp of 1.
// SURVEY 8(d) strength-reduction caveat; measured on config 4: see DESIGN.md section 4.
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void gm_rule_thread_osc(const double *pc, const double (&a)[D], const double (&b)[D],
                                                   double &Iout, double &Eout, int &kdiv) {
    constexpr int NPAR = ib::fam_nparam(IB200_FAM_GENZ_OSC, D);
    const double *par = pc, *lbs = pc + NPAR, *den = lbs + D;
    const double finit = den[D], jacc = den[D + 1];
    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
    constexpr double twon = (double)(1 << D);
    constexpr double w1 = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0);
    constexpr double w2 = twon * (980 / 6561.0);
    constexpr double w3 = twon * ((1820 - 400 * D) / 19683.0);
    constexpr double w4 = twon * (200 / 19683.0);
    constexpr double w5 = 6859 / 19683.0;
    constexpr double w4p = twon * (25 / 729.0);
    constexpr double w3p = twon * ((265 - 100 * D) / 1458.0);
    constexpr double w2p = twon * (245 / 486.0);
    constexpr double w1p = twon * ((729 - 950 * D + 50 * D * D) / 729.0);

    double h[D], V = 1.0, S = finit;
    double ang[3 * D + 1];                       // [theta_k(l2) | theta_k(l3) | theta_k(l5) | S_c]
    {
        double T0[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const double c = 0.5 * (a[k] + b[k]);
            h[k] = 0.5 * fabs(b[k] - a[k]);
            V = (k == 0) ? h[0] : V * h[k];
            T0[k] = par[k] * (lbs[k] + (1.0 + c) * den[k]);
            const double g = par[k] * den[k];
            ang[k] = g * (h[k] * l2); ang[D + k] = g * (h[k] * l3); ang[2 * D + k] = g * (h[k] * l5);
        }
#pragma unroll
        for (int k = D - 1; k >= 0; k--) S = S + T0[k];
        ang[3 * D] = S;
    }
    double sn[3 * D + 1], cs[3 * D + 1];
    {
        constexpr int NA = 3 * D + 1;
        bool big = false;
#pragma unroll
        for (int g0 = 0; g0 < NA; g0 += 4) {
            double x4[4], s4[4], c4[4];
#pragma unroll
            for (int j = 0; j < 4; j++) x4[j] = ang[g0 + j < NA ? g0 + j : NA - 1];
            big = ib::sincos_bfN_fast<4>(x4, s4, c4) || big;
#pragma unroll
            for (int j = 0; j < 4; j++) if (g0 + j < NA) { sn[g0 + j] = s4[j]; cs[g0 + j] = c4[j]; }
        }
        if (big) {
            // an angle of 2^30 or more (or NaN): all of them through libm, in ONE out-of-line call.  Inlined per group, libm's slow
            // paths sat between the groups and stretched the rule's straight-line code to 37 KB -- more than the 32 KB instruction
            // cache level holds (ncu: no_instruction was 24 % of the stall samples of the 4-D oscillatory kernel)
            double tx[NA], ts[NA], tc[NA];
#pragma unroll
            for (int j = 0; j < NA; j++) tx[j] = ang[j];
            ib::sincos_libm_n(tx, ts, tc, NA);
#pragma unroll
            for (int j = 0; j < NA; j++) { sn[j] = ts[j]; cs[j] = tc[j]; }
        }
    }
    const double cS = cs[3 * D], sS = sn[3 * D];
    const double f1 = cS;
    const double twelvef1 = 12 * f1;
    // axis points: cos(S +- theta) = cS cos(theta) -+ sS sin(theta)
    double divdiff[D], f2 = 0.0, f3 = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        const double c2 = cS * cs[i], s2 = sS * sn[i], c3 = cS * cs[D + i], s3 = sS * sn[D + i];
        const double f2i = (c2 - s2) + (c2 + s2);
        const double f3i = (c3 - s3) + (c3 + s3);
        f2 += f2i; f3 += f3i;
        divdiff[i] = fabs(f3i + twelvef1 - 7 * f2i);
    }
    // pair points: S +- theta_i +- theta_j at lambda_3
    double f4 = 0.0;
#pragma unroll
    for (int i = 0; i < D - 1; i++) {
#pragma unroll
        for (int j = i + 1; j < D; j++) {
            const double cc = cs[D + i] * cs[D + j], ss = sn[D + i] * sn[D + j], sc = sn[D + i] * cs[D + j], csn = cs[D + i] * sn[D + j];
            const double cp = cc - ss, cm = cc + ss, sp = sc + csn, sm = sc - csn;   // cos / sin of theta_i +- theta_j
            const double A = cS * cp, B = sS * sp, C = cS * cm, Dd = sS * sm;
            f4 += A - B;        // (+, +)
            f4 += C - Dd;       // (+, -)
            f4 += C + Dd;       // (-, +)
            f4 += A + B;        // (-, -)
        }
    }
    // corner points: S + sum_k +-theta_k(l5).  Unit vectors e^{i(+theta_{D-1} +- ... +- theta_0)} are built axis by axis (the
    // combinations with -theta_{D-1} are their conjugates): 2^(D-1) complex numbers, two points each.
    double f5 = 0.0;
    {
        double re[1 << (D - 1)], im[1 << (D - 1)];
        re[0] = cs[2 * D + D - 1]; im[0] = sn[2 * D + D - 1];
#pragma unroll
        for (int k = D - 2; k >= 0; k--) {
            const int m = 1 << (D - 2 - k);                     // combinations so far
            const double ck = cs[2 * D + k], sk = sn[2 * D + k];
#pragma unroll
            for (int t = m - 1; t >= 0; t--) {
                const double r0 = re[t], i0 = im[t];
                const double rc = r0 * ck, is = i0 * sk, rs = r0 * sk, ic = i0 * ck;
                re[2 * t] = rc - is; im[2 * t] = ic + rs;       // * e^{+i theta_k}
                re[2 * t + 1] = rc + is; im[2 * t + 1] = ic - rs;   // * e^{-i theta_k}
            }
        }
#pragma unroll
        for (int t = 0; t < (1 << (D - 1)); t++) {
            const double A = cS * re[t], B = sS * im[t];
            f5 += A - B;                                        // cos(S + phi_t)
            f5 += A + B;                                        // cos(S - phi_t)
        }
    }
    const double Vj = V * jacc;                                 // the constant Jacobian of the finite map, factored out of the sums
    const double Iv = Vj * (w1 * f1 + w2 * f2 + w3 * f3 + w4 * f4 + w5 * f5);
    const double Ip = Vj * (w1p * f1 + w2p * f2 + w3p * f3 + w4p * f4);
    const double Er = fabs(Iv - Ip);
    double p10 = 1.0;
#pragma unroll
    for (int k = 0; k < D; k++) p10 *= 10.0;
    const double deltaf = Er / (p10 * V);
    int kdv = 0; double maxdd = 0.0, hkd = h[0];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double dd = divdiff[k] * fabs(jacc);              // fourth differences of g = f * jacc
        const double delta = dd - maxdd;
        if (delta > deltaf) { kdv = k; maxdd = dd; hkd = h[k]; }
        else if (fabs(delta) <= deltaf && h[k] > hkd) { kdv = k; hkd = h[k]; }
    }
    Iout = Iv; Eout = Er; kdiv = kdv;
}

// ------------------------------------------------------------------------------------------------
// D == 1: the pair machinery runs QuadGK (QuadGK.quadgk as called at src/Integrals.jl:327-335; oracle: orc_quadgk).
// One application of the (7,15) Gauss-Kronrod rule to the segment (a, b) by one thread, summed in QuadGK's order
// (symmetric pairs first; oracle gk_evalrule).
// ------------------------------------------------------------------------------------------------
template <int FAM, bool ALLFIN>
__device__ __forceinline__ void gk_rule_thread(const double *pc, int tr, double a, double b, double &Iout, double &Eout) {
    using F = ib::Family<FAM, 1>;
    constexpr int NPAR = ib::fam_nparam(FAM, 1);
    const double *par = pc, *lbs = pc + NPAR, *aux = lbs + 1;
    const double jacc = aux[2];
    double q[F::NQ];
    F::prepare(par, q);
    constexpr double X[7] = {-0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
                             -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
                             -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
                             -0.207784955007898467600689403773245};
    constexpr double W[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                             0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                             0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                             0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
    constexpr double WG[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                              0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
    const double s = 0.5 * (b - a);
    auto g = [&](double t) {
        double u[1], J;
        if (ALLFIN) { u[0] = lbs[0] + (1.0 + t) * aux[0]; J = jacc; }
        else { ib::AxisMap m; ib::make_axis(lbs[0], aux[0], m); ib::t2ujac(tr, t, m, u[0], J); }
        return F::eval(u, q) * J;
    };
    double fp[7], fm[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { fp[j] = g(a + (1 + X[j]) * s); fm[j] = g(a + (1 - X[j]) * s); }
    const double f0 = g(a + s);
    double fg = fp[1] + fm[1], fk = fp[0] + fm[0];
    double Ig = fg * WG[0];
    double Ik = fg * W[1] + fk * W[0];
#pragma unroll
    for (int i = 2; i <= 3; i++) {
        fg = fp[2 * i - 1] + fm[2 * i - 1];
        fk = fp[2 * i - 2] + fm[2 * i - 2];
        Ig += fg * WG[i - 1];
        Ik += fg * W[2 * i - 1] + fk * W[2 * i - 2];
    }
    Ig += f0 * WG[3];
    Ik += f0 * W[7] + (fp[6] + fm[6]) * W[6];
    Iout = Ik * s;
    Eout = fabs(Ik * s - Ig * s);
}
// the same for a Gauss-Kronrod rule of any order n >= 2 (QuadGKJL(order = n)): tables in shared memory, QuadGK's evalrule
// for any order (oracle gk_evalrule_n): even orders have no Gauss node at the centre
constexpr int kGkMaxOrder = 30;
template <int FAM, bool ALLFIN>
__device__ __forceinline__ void gk_rule_thread_n(const double *pc, int tr, double a, double b, int n, const double *x, const double *w,
                                                 const double *wg, double &Iout, double &Eout) {
    using F = ib::Family<FAM, 1>;
    constexpr int NPAR = ib::fam_nparam(FAM, 1);
    const double *par = pc, *lbs = pc + NPAR, *aux = lbs + 1;
    const double jacc = aux[2];
    double q[F::NQ];
    F::prepare(par, q);
    const double s = 0.5 * (b - a);
    auto g = [&](double t) {
        double u[1], J;
        if (ALLFIN) { u[0] = lbs[0] + (1.0 + t) * aux[0]; J = jacc; }
        else { ib::AxisMap m; ib::make_axis(lbs[0], aux[0], m); ib::t2ujac(tr, t, m, u[0], J); }
        return F::eval(u, q) * J;
    };
    auto pair = [&](int i) { return g(a + (1 + x[i - 1]) * s) + g(a + (1 - x[i - 1]) * s); };     // 1-based like the Julia source
    const int lenx = n + 1, lenwg = (n + 1) / 2, n1 = 1 - (lenx & 1);
    double fg = pair(2), fk = pair(1);
    double Ig = fg * wg[0];
    double Ik = fg * w[1] + fk * w[0];
    for (int i = 2; i <= lenwg - n1; i++) {
        fg = pair(2 * i); fk = pair(2 * i - 1);
        Ig += fg * wg[i - 1];
        Ik += fg * w[2 * i - 1] + fk * w[2 * i - 2];
    }
    if (n1 == 0) Ik += g(a + s) * w[lenx - 1];
    else {
        const double f0 = g(a + s);
        Ig += f0 * wg[lenwg - 1];
        Ik += f0 * w[lenx - 1] + pair(lenx - 1) * w[lenx - 2];
    }
    Iout = Ik * s;
    Eout = fabs(Ik * s - Ig * s);
}
// QuadGK's check_endpoint_roundoff: does the rule evaluate at an endpoint of (a, b)?
__device__ __forceinline__ bool gk_endpoint_roundoff(double a, double b, double x1 = -0.991455371120812639206854697526329) {
    const double c = 0.5 * (b - a);
    return (a == a + (1 + x1) * c) || (b == a + (1 - x1) * c);
}

// ------------------------------------------------------------------------------------------------
// Tournament helpers.  A candidate is (key, index); NONE marks "no candidate".  Larger key wins,
// ties go to the lower index (the reference order: largest error first; ties -> lowest slot).
// ------------------------------------------------------------------------------------------------
constexpr int NONE = 0x7fffffff;
struct Cand { unsigned long long k; int idx; };
__device__ __forceinline__ bool beats(const Cand &x, const Cand &y) {     // x strictly better than y
    return x.idx != NONE && (y.idx == NONE || x.k > y.k || (x.k == y.k && x.idx < y.idx));
}

// The two 32-entry groups on the selected path are needed only AFTER the rule evaluation (runner-up scans).  They are
// staged into this lane's shared-memory slot with cp.async while the rule runs, so the scans read shared memory instead
// of waiting on L2/DRAM (ncu, profiles/r01_hcub_pair_prod_c4_*: 31 % of the stall samples sat on these loads).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// stage this lane's half (16 doubles) of the group at arr[base ..] (n valid entries overall) into stage[0..16)
__device__ __forceinline__ void stage_half(double *stage, const double *arr, int base, int n, int half) {
    const int e0 = base + half * 16;
#pragma unroll
    for (int t = 0; t < 8; t++) if (e0 + 2 * t < n) cp_async16(stage + 2 * t, arr + e0 + 2 * t);
}
// best entry of this lane's half (16 entries) of a staged 32-entry group (n valid entries overall), skipping group
// position `excl`; idx is the group position 0..31
__device__ __forceinline__ Cand scan_staged_excl(const double *stage, int base, int n, int half, int excl) {
    Cand r; r.k = 0ULL; r.idx = NONE;
    const int e0 = base + half * 16;
    const double2 *p = reinterpret_cast<const double2 *>(stage);
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int ia = half * 16 + 2 * t, ib_ = ia + 1;
        const double2 v = p[t];
        if (e0 + 2 * t < n && ia != excl) { const unsigned long long k = dk(v.x); if (r.idx == NONE || k > r.k) { r.k = k; r.idx = ia; } }
        if (e0 + 2 * t + 1 < n && ib_ != excl) { const unsigned long long k = dk(v.y); if (r.idx == NONE || k > r.k) { r.k = k; r.idx = ib_; } }
    }
    return r;
}
// combine the two halves of a pair; identical result in both lanes
__device__ __forceinline__ Cand combine_pair(unsigned pm, const Cand &mine) {
    Cand o;
    o.k = __shfl_xor_sync(pm, mine.k, 1);
    o.idx = __shfl_xor_sync(pm, mine.idx, 1);
    return beats(o, mine) ? o : mine;
}

#ifndef HP_MINB
#define HP_MINB 1
#endif
template <int FAM, int D, bool ALLFIN>
__global__ void __launch_bounds__(kBlock, HP_MINB)
hcub_pair_kernel(const Params p) {
    using F = ib::Family<FAM, D>;
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    constexpr int PCD = pc_doubles<FAM, D>();
    constexpr int RECW = 2 * D + 2;
    extern __shared__ double sm[];
    const int NP = (D == 1 && p.gk_order) ? 2 * p.gk_order + 1 : hc::gm_npoints(D);      // evaluations per rule application
    const int lane = threadIdx.x & 31, half = lane & 1;
    const unsigned full = 0xffffffffu;
    const unsigned pm = 3u << (lane & ~1);               // this pair's lanes
    const int pib = threadIdx.x >> 1;
    const int64_t gpair = (int64_t)blockIdx.x * kPairs + pib;
    const int n2max = p.cap >> 10;
    double *topk = sm + (size_t)pib * row_doubles(n2max, PCD);
    int *tops = reinterpret_cast<int *>(topk + n2max);
    double *pc = topk + n2max + (n2max + 1) / 2;
    double *par = pc, *lbs = pc + NPAR, *aux = lbs + D;
    double *stage = sm + (size_t)kPairs * row_doubles(n2max, PCD) + (size_t)threadIdx.x * 32;          // [keys half | l1 half]
    unsigned char *stage_off = reinterpret_cast<unsigned char *>(sm + (size_t)kPairs * row_doubles(n2max, PCD) + (size_t)kBlock * 32) + pib * 32;
    double *gkt = sm + (size_t)kPairs * row_doubles(n2max, PCD) + (size_t)kBlock * 32 + (size_t)kPairs * 4;    // D == 1: x | w | wg of the rule
    if (D == 1 && p.gk_order) {
        for (int t = threadIdx.x; t <= p.gk_order; t += kBlock) { gkt[t] = p.gk_x[t]; gkt[kGkMaxOrder + 1 + t] = p.gk_w[t]; }
        for (int t = threadIdx.x; t < (p.gk_order + 1) / 2; t += kBlock) gkt[2 * (kGkMaxOrder + 1) + t] = p.gk_wg[t];
        __syncthreads();
    }

    double *keys = p.keys + (size_t)gpair * p.cap;
    double *l1 = p.l1 + (size_t)gpair * (p.cap / kFan);
    unsigned char *l1off = p.l1off + (size_t)gpair * (p.cap / kFan);
    double *rec = p.rec + (size_t)gpair * p.cap * RECW;

    double rtol = p.rtol;
    const double atol = p.atol;
    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps), HCubature default

    const int64_t nwork = p.norder ? (int64_t)*p.norder : p.nprob;
    // pair state (identical in both lanes)
    int state = ST_NEED, nbox = 0, status = 0;
    int64_t prob = -1, nev = 0, q = 0, qtotal = 0;
    double I = 0.0, E = 0.0;
    bool degenerate = false;

    // register box s (record and key already written by its owner) in the tournament: the new box has the highest
    // slot so far, so it only displaces strictly smaller maxima.  Both lanes compute, lane 0 stores.
    auto tour_append = [&](int s, double Ebox) {
        const int i = s >> 5, j = s >> 10;
        Cand g; g.k = dk(Ebox); g.idx = s;
        if (s & 31) { const unsigned long long ok = dk(__ldcg(l1 + i)); if (!(g.k > ok)) { g.k = ok; g.idx = (i << 5) + __ldcg(l1off + i); } }
        Cand t = g;
        if (s & 1023) { const unsigned long long ok = dk(topk[j]); if (!(t.k > ok)) { t.k = ok; t.idx = tops[j]; } }
        __syncwarp(pm);
        if (half == 0) { l1[i] = kd2d(g.k); l1off[i] = (unsigned char)(g.idx & 31); topk[j] = kd2d(t.k); tops[j] = t.idx; }
        __syncwarp(pm);
    };

    for (;;) {
        __syncwarp(full);
        // ---- 1. finished problems: re-sum over all boxes (HCubature's final sum), whole warp, coalesced keys
        unsigned fin = __ballot_sync(full, state == ST_FIN) & 0x55555555u;
        while (fin) {
            const int src = __ffs(fin) - 1;
            fin &= fin - 1;
            const int nb = __shfl_sync(full, nbox, src);
            const double *kk = reinterpret_cast<const double *>((uintptr_t)__shfl_sync(full, (unsigned long long)(uintptr_t)keys, src));
            const double *rr = reinterpret_cast<const double *>((uintptr_t)__shfl_sync(full, (unsigned long long)(uintptr_t)rec, src));
            double pI = 0.0, pE = 0.0;
            int s = lane;
            for (; s + 96 < nb; s += 128) {
                const double i0 = __ldcg(rr + (size_t)s * RECW + 2 * D), i1 = __ldcg(rr + (size_t)(s + 32) * RECW + 2 * D);
                const double i2 = __ldcg(rr + (size_t)(s + 64) * RECW + 2 * D), i3 = __ldcg(rr + (size_t)(s + 96) * RECW + 2 * D);
                const double e0 = __ldcg(kk + s), e1 = __ldcg(kk + s + 32), e2 = __ldcg(kk + s + 64), e3 = __ldcg(kk + s + 96);
                pI += i0; pE += e0; pI += i1; pE += e1; pI += i2; pE += e2; pI += i3; pE += e3;
            }
            for (; s < nb; s += 32) { pI += __ldcg(rr + (size_t)s * RECW + 2 * D); pE += __ldcg(kk + s); }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pI += __shfl_xor_sync(full, pI, off);
                pE += __shfl_xor_sync(full, pE, off);
            }
            if (lane == src) {
                p.out_I[prob] = pI;
                if (p.out_E) p.out_E[prob] = pE;
                if (p.out_nevals) p.out_nevals[prob] = nev;
                if (p.out_status) p.out_status[prob] = status;
            }
        }
        if (state == ST_FIN) state = ST_NEED;

        // ---- 2. refill
        if (state == ST_NEED) {
            unsigned long long pidx = 0;
            if (half == 0) pidx = atomicAdd(p.counter, 1ULL);
            pidx = __shfl_sync(pm, pidx, lane & ~1);
            if ((int64_t)pidx >= nwork) state = ST_EXHAUSTED;
            else {
                prob = p.order ? (int64_t)p.order[pidx] : (int64_t)pidx;
                __syncwarp(pm);
                for (int t = half; t < NPAR; t += 2) par[t] = p.params[prob * p.nparam + t];
                for (int k = half; k < D; k += 2) {
                    const double l = p.bpp ? p.lb[prob * D + k] : p.lb[k];
                    const double u = p.bpp ? p.ub[prob * D + k] : p.ub[k];
                    lbs[k] = l; aux[k] = ALLFIN ? (u - l) * 0.5 : u;
                }
                __syncwarp(pm);
                double jacc = 1.0;
                if (ALLFIN) {
#pragma unroll
                    for (int k = 0; k < D; k++) jacc = (k == 0) ? aux[0] : jacc * aux[k];
                }
                if (half == 0) { aux[D] = F::sep_init(par); aux[D + 1] = jacc; }
                __syncwarp(pm);
                // initial grid (HCubature initdiv; CartesianIndices order, first index fastest)
                qtotal = 1;
                double prodD = 1.0;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    qtotal *= p.initdiv;
                    double tl = -1.0, th = 1.0;
                    if (!ALLFIN) { ib::AxisMap m; ib::make_axis(lbs[k], aux[k], m); tl = m.tlo; th = m.thi; }
                    prodD *= (th - tl) / p.initdiv;
                }
                degenerate = (prodD == 0.0);
                if (degenerate) qtotal = 1;                 // HCubature returns after the first box
                q = 0; nbox = 0; nev = 0; I = 0.0; E = 0.0; status = IB200_ST_CONVERGED;
                state = ST_INIT;
            }
        }
        if (__all_sync(full, state == ST_EXHAUSTED)) break;

        // ---- 3. choose this lane's box
        double a[D], b[D];
        bool have = false;
        int slot = 0;
        double Ebox = 0.0, Ibox = 0.0;
        unsigned long long app_k = 0ULL; int app_off = 0;   // tournament entry of the group the appended box joins
        if (state == ST_REFINE) {
            if (degenerate || E <= fmax(rtol * fabs(I), atol)) state = ST_FIN;
            else if (nev >= p.maxevals) { status = IB200_ST_MAXEVALS; state = ST_FIN; }
            else if (!isfinite(E)) { status = IB200_ST_NONFINITE; state = ST_FIN; }
            else if (nbox + 1 > p.cap) { status = IB200_ST_MAXEVALS; state = ST_FIN; }
        }
        if (state == ST_REFINE) {
            // top level (shared): the pair splits the n2 block maxima, lower indices to lane 0; the winner's stored
            // slot IS the worst box
            const int n2 = (nbox + 1023) >> 10;
            {
                const int m = (n2 + 1) >> 1;
                const int lo = half ? m : 0, hi = half ? n2 : m;
                Cand bst; bst.k = 0ULL; bst.idx = NONE;
#pragma unroll 4
                for (int t = lo; t < hi; t++) { const unsigned long long k = dk(topk[t]); if (bst.idx == NONE || k > bst.k) { bst.k = k; bst.idx = t; } }
                bst = combine_pair(pm, bst);
                slot = tops[bst.idx];
                Ebox = kd2d(bst.k);
            }
            // the only global load in front of the evaluation: the parent's record
            const double2 *src = reinterpret_cast<const double2 *>(rec + (size_t)slot * RECW);
            double r[RECW];
#pragma unroll
            for (int t = 0; t < D + 1; t++) { const double2 v = __ldcg(src + t); r[2 * t] = v.x; r[2 * t + 1] = v.y; }
            // start staging the two groups on the selected path (+ the block's 32 position bytes) for the runner-up scans
            {
                const int i = slot >> 5, j = slot >> 10;
                stage_half(stage, keys, i << 5, nbox, half);
                stage_half(stage + 16, l1, j << 5, (nbox + 31) >> 5, half);
                cp_async16(stage_off + half * 16, l1off + (j << 5) + half * 16);
                cp_async_commit();
            }
            if (nbox & 31) { app_k = dk(__ldcg(l1 + (nbox >> 5))); app_off = __ldcg(l1off + (nbox >> 5)); }
            Ibox = r[2 * D];
            const int kdp = (int)r[2 * D + 1];
            double akd = 0.0, bkd = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) { a[k] = r[k]; b[k] = r[D + k]; if (k == kdp) { akd = r[k]; bkd = r[D + k]; } }
            have = true;
            if (D == 1) {
                // QuadGK: mid = (a + b) / 2; the pair evaluates (a, mid) and (mid, b); a segment that can no longer be
                // bisected ends the refinement (check_endpoint_roundoff)
                const double mid = (akd + bkd) / 2;
                const double x1 = p.gk_order ? gkt[0] : -0.991455371120812639206854697526329;
                if (gk_endpoint_roundoff(akd, mid, x1) || gk_endpoint_roundoff(mid, bkd, x1)) { status = IB200_ST_ROUNDOFF; state = ST_FIN; have = false; }
                if (half == 0) b[0] = mid; else a[0] = mid;
            } else {
                const double w = (bkd - akd) * 0.5;
                // child 0 = upper half (a[kd] += w), child 1 = lower half (b[kd] -= w)
#pragma unroll
                for (int k = 0; k < D; k++) {
                    if (k == kdp) { if (half == 0) a[k] = akd + w; else b[k] = bkd - w; }
                }
            }
        } else if (state == ST_INIT) {
            const int64_t myq = q + half;
            have = myq < qtotal;
            int64_t r = myq;
#pragma unroll
            for (int k = 0; k < D; k++) {
                double tl = -1.0, th = 1.0;
                if (!ALLFIN) { ib::AxisMap m; ib::make_axis(lbs[k], aux[k], m); tl = m.tlo; th = m.thi; }
                const int cidx = (int)(r % p.initdiv); r /= p.initdiv;
                const double Dl = (th - tl) / p.initdiv;
                a[k] = tl + cidx * Dl;
                b[k] = (cidx == p.initdiv - 1) ? th : a[k] + Dl;
            }
        }

        // ---- 4. the rule on this lane's box
        double Ic = 0.0, Ec = 0.0; int kc = 0;
        if (have) {
            if constexpr (D == 1) {
                if (p.gk_order) gk_rule_thread_n<FAM, ALLFIN>(pc, p.tr, a[0], b[0], p.gk_order, gkt, gkt + kGkMaxOrder + 1, gkt + 2 * (kGkMaxOrder + 1), Ic, Ec);
                else gk_rule_thread<FAM, ALLFIN>(pc, p.tr, a[0], b[0], Ic, Ec);
            }
            else if constexpr (FAM == IB200_FAM_GENZ_OSC && ALLFIN) gm_rule_thread_osc<D>(pc, a, b, Ic, Ec, kc);
            else gm_rule_thread<FAM, D, ALLFIN>(pc, p.tr, a, b, Ic, Ec, kc);
        }

        // ---- 5. bookkeeping (per pair)
        if (state == ST_REFINE || state == ST_INIT) {
            const double Io = __shfl_xor_sync(pm, Ic, 1), Eo = __shfl_xor_sync(pm, Ec, 1);
            const double I1 = half ? Io : Ic, I2 = half ? Ic : Io, E1 = half ? Eo : Ec, E2 = half ? Ec : Eo;
            auto write_box = [&](int s) {
                double2 *dst = reinterpret_cast<double2 *>(rec + (size_t)s * RECW);
                double r[RECW];
#pragma unroll
                for (int k = 0; k < D; k++) { r[k] = a[k]; r[D + k] = b[k]; }
                r[2 * D] = Ic; r[2 * D + 1] = (double)kc;
#pragma unroll
                for (int t = 0; t < D + 1; t++) dst[t] = make_double2(r[2 * t], r[2 * t + 1]);
                keys[s] = Ec;
            };
            if (state == ST_REFINE) {
                nev += 2 * NP;
                if (D == 1) { I = (I - Ibox) + I1 + I2; E = (E - Ebox) + E1 + E2; }     // QuadGK's order
                else { I += I1 + I2 - Ibox; E += E1 + E2 - Ebox; }
                const int s2 = nbox;
                const int i = slot >> 5, j = slot >> 10, i2 = s2 >> 5, j2 = s2 >> 10;
                // runner-ups of the two groups on the selected path (data prefetched into L2 before the evaluation)
                cp_async_wait_all();
                __syncwarp(pm);                                                           // the position bytes are read across the pair
                Cand ru0 = combine_pair(pm, scan_staged_excl(stage, i << 5, nbox, half, slot & 31));
                Cand ru1 = combine_pair(pm, scan_staged_excl(stage + 16, j << 5, (nbox + 31) >> 5, half, i & 31));
                if (ru0.idx != NONE) ru0.idx += i << 5;                                   // -> slot
                if (ru1.idx != NONE) ru1.idx = (((j << 5) + ru1.idx) << 5) + stage_off[ru1.idx];   // -> slot of that group's maximum
                write_box(half == 0 ? slot : s2);
                // group i: runner-up, child 0 at `slot`, child 1 if it lands in the same group
                Cand g0; g0.k = dk(E1); g0.idx = slot;
                if (beats(ru0, g0)) g0 = ru0;
                if (i2 == i && dk(E2) > g0.k) { g0.k = dk(E2); g0.idx = s2; }
                // block j: runner-up group, group i, and the appended box's group if it is another group of this block
                Cand g1 = g0;
                if (beats(ru1, g1)) g1 = ru1;
                Cand a0; a0.k = dk(E2); a0.idx = s2;
                Cand t2 = a0;
                if (i2 != i) {
                    if ((s2 & 31) && !(a0.k > app_k)) { a0.k = app_k; a0.idx = (i2 << 5) + app_off; }
                    if (j2 == j) { if (beats(a0, g1)) g1 = a0; }
                    else { t2 = a0; if (s2 & 1023) { const unsigned long long ok = dk(topk[j2]); if (!(t2.k > ok)) { t2.k = ok; t2.idx = tops[j2]; } } }
                }
                __syncwarp(pm);
                if (half == 0) {
                    l1[i] = kd2d(g0.k); l1off[i] = (unsigned char)(g0.idx & 31);
                    if (i2 != i) {
                        l1[i2] = kd2d(a0.k); l1off[i2] = (unsigned char)(a0.idx & 31);
                        if (j2 != j) { topk[j2] = kd2d(t2.k); tops[j2] = t2.idx; }
                    }
                    topk[j] = kd2d(g1.k); tops[j] = g1.idx;
                }
                nbox++;
                __syncwarp(pm);
            } else {
                const bool second = q + 1 < qtotal;
                if (have) write_box(nbox + half);
                __syncwarp(pm);
                if (q == 0) { I = I1; E = E1; } else { I += I1; E += E1; }
                nev += NP;
                tour_append(nbox, E1);
                nbox++;
                if (second) { I += I2; E += E2; nev += NP; tour_append(nbox, E2); nbox++; }
                q += 2;
                if (q >= qtotal) state = ST_REFINE;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Longest-expected-first ordering.  Problem costs in one batch differ by 10^3 (BASELINE config 4: 10^4 .. 4*10^6
// evaluations) and a problem is a strictly sequential chain of refinement steps, so a late-starting expensive problem
// leaves the rest of the GPU idle (measured: ~0.4 s of a 1.4 s launch).  A pilot pass gives every problem a few
// refinement steps; the ones that do not converge are re-run from scratch in order of decreasing E / tol after the
// pilot (rank correlation with the final cost 0.6-0.9 on the Genz families), found by a counting sort on the top
// 16 bits of that ratio.  Results are unaffected: every problem is refined exactly as before, only its start time moves.
// ------------------------------------------------------------------------------------------------
constexpr int kLptBuckets = 4096;
__device__ __forceinline__ int lpt_bucket(double I, double E, double rtol, double atol) {
    const double tol = fmax(rtol * fabs(I), atol);
    const double ratio = E / tol;                                     // > 1 for an unconverged problem; NaN/inf -> first bucket
    if (!(ratio < 1.0e300)) return 0;
    long long b = (__double_as_longlong(fmax(ratio, 1.0)) >> 47) - (1023LL << 5);   // 32 buckets per octave
    b = b < 0 ? 0 : (b > kLptBuckets - 1 ? kLptBuckets - 1 : b);
    return kLptBuckets - 1 - (int)b;                                  // descending ratio
}
// (lpt_hist_kernel / lpt_scan_kernel / lpt_scatter_kernel: hcubature.cu)

// bytes of box pool per pair
inline size_t pool_bytes_per_pair(int D, int cap) {
    return (size_t)cap * (8 + 8 * (2 * (size_t)D + 2)) + (size_t)(cap / kFan) * 9;
}

template <int D>
int32_t launch(ib200_ctx *ctx, int family, bool allfin, Params &p);

template <int D>
int32_t launch_impl(ib200_ctx *ctx, int family, bool allfin, Params &p) {
    int32_t rc = IB200_ERR_UNSUPPORTED;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, D)) {
            auto go = [&](auto kern) -> int32_t {
                const size_t smem = ((size_t)kPairs * row_doubles(p.cap >> 10, pc_doubles<FAM, D>()) + (size_t)kBlock * 32 + (size_t)kPairs * 4 +
                                     (D == 1 ? 3 * (kGkMaxOrder + 1) : 0)) * sizeof(double);
                if (smem > ctx->smem_optin) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_hcubature_batch: box capacity too large for shared memory");
                IB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int occ = 0;
                IB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem));
                if (occ < 1) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_hcubature_batch: kernel does not fit on an SM");
                if (ctx->opt_adaptive_warps_per_sm > 0) {
                    const int want = (int)((ctx->opt_adaptive_warps_per_sm + 3) / 4);
                    if (want < occ) occ = want;
                }
                int64_t blocks = (p.nprob + kPairs - 1) / kPairs;
                if (blocks > (int64_t)ctx->sm_count * occ) blocks = (int64_t)ctx->sm_count * occ;
                // the pools are sized by residency: shrink the grid until they fit in free HBM
                const size_t per_pair = pool_bytes_per_pair(D, p.cap);
                const auto fits_held = [&](int64_t nb) {      // the pools this grid needs are already allocated
                    const size_t pr = (size_t)nb * kPairs;
                    return ctx->scratch[SL_WORK0].cap >= pr * p.cap * sizeof(double) &&
                           ctx->scratch[SL_WORK2].cap >= pr * p.cap * (2 * D + 2) * sizeof(double) &&
                           ctx->scratch[SL_WORK4].cap >= pr * (p.cap / kFan) * sizeof(double) && ctx->scratch[SL_WORK5].cap >= pr * (p.cap / kFan);
                };
                if (!fits_held(blocks)) {                      // (cudaMemGetInfo costs milliseconds: only when the pools have to grow)
                    size_t fr = 0, tot = 0;
                    IB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
                    size_t held = 0;
                    for (int s : {SL_WORK0, SL_WORK2, SL_WORK3, SL_WORK4, SL_WORK5}) held += ctx->scratch[s].cap;
                    const size_t budget = (size_t)((double)(fr + held) * 0.85);
                    while (blocks > 1 && (size_t)blocks * kPairs * per_pair > budget) blocks = blocks > ctx->sm_count ? blocks - ctx->sm_count : blocks - 1;
                }
                const size_t pairs = (size_t)blocks * kPairs;
                p.keys = (double *)ctx->get(SL_WORK0, pairs * p.cap * sizeof(double));
                p.rec = (double *)ctx->get(SL_WORK2, pairs * p.cap * (2 * D + 2) * sizeof(double));
                p.l1 = (double *)ctx->get(SL_WORK4, pairs * (p.cap / kFan) * sizeof(double));
                p.l1off = (unsigned char *)ctx->get(SL_WORK5, pairs * (p.cap / kFan));
                if (!p.keys || !p.rec || !p.l1 || !p.l1off)
                    return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: box pool allocation failed (lower maxiters or the hcub_capacity option)");
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(p);
                IB_LAUNCH_CHECK(ctx);
                return IB200_OK;
            };
            rc = allfin ? go(hcub_pair_kernel<FAM, D, true>) : go(hcub_pair_kernel<FAM, D, false>);
        }
    });
    if (rc == IB200_ERR_UNSUPPORTED && ctx->err.empty())
        return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_batch: family not compiled in for this dimension (pair mode)");
    return rc;
}

}  // namespace hp
