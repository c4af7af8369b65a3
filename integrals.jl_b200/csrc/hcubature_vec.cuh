// hcubature_vec.cuh -- batched adaptive cubature of VECTOR-valued integrands (Genz-Malik degree 7/5, or GK(7,15) in
// 1-D), one warp per problem.  Included by hcubature_vec_d<D>.cu, one translation unit per dimension.
//
// SURVEY 8(f) rank 1.  Replaces HCubature.hcubature / hquadrature(f, lb, ub; norm, rtol, atol, maxevals, initdiv) for
// array-valued integrands as called from __solvebp_call(::HCubatureJL) (src/Integrals.jl:347-397, `norm = alg.norm` at
// :380-393; HCubatureJL.norm, src/algorithms.jl:66-67,108-115), behind the automatic ChangeOfVariables
// (src/common.jl:97-103).  Component c of problem k is the registered family evaluated with parameter column
// params[:, c, k]; the ncomp components share the points and the subdivision; the error of a box is norm(I - I') over
// the components, the fourth differences that choose the split axis are norms over the components, and the refinement
// stops when E <= max(rtol * norm(I), atol).  Oracle: orc_hcubature_vec / gm_rule_vec.
//
// Mapping: the driver of hcubature_impl.cuh (persistent grid, two boxes = the two children of a bisection per rule
// pass, box pool {key = error, I[ncomp], a, b, split axis} in global memory with the two-level tournament).  The rule
// tabulates the transformed coordinate and Jacobian of the 7 abscissae of every axis once per box (shared by all
// components), spreads the 2 * NP points over the lanes, and a lane evaluates every component at its point with
// Family::eval -- the oracle's operation order.
#pragma once
#include "hcubature_impl.cuh"

namespace hv {

constexpr int kWarpsPerBlock = hc::kWarpsPerBlock;
constexpr int kBlock = hc::kBlock;
constexpr int kGroup = hc::kGroup;
constexpr int NC = 8;           // maximum number of components

struct Params {
    const double *params; int nparam; int ncomp; int64_t nprob;
    const double *lb, *ub; int bpp, tr;
    double rtol, atol; int64_t maxevals; int initdiv; int normkind;
    ib::SensIdx sens;           // sens.n > 0: sensitivities -- params is nparam x nprob, components [f, df/dpar[idx[0]], ...] (ib::Grad)
    int cap;                    // boxes per warp (multiple of kGroup)
    double *keys;               // [warps][cap]          error of every box
    double *ivals;              // [warps][cap][ncomp]   integral of every box
    double *recs;               // [warps][cap][RECW]    a[D], b[D], kdiv
    unsigned long long *counter;
    double *out_I, *out_E; int64_t *out_nevals; int32_t *out_status;
};

__device__ __forceinline__ double vnorm(const double (&v)[NC], int nc, int kind) {
    double r = 0.0;
    if (kind == 2) return fabs(v[0]);          // the first component alone (sensitivities: a ForwardDiff.Dual's norm is its value's)
    if (kind == 1) {
#pragma unroll
        for (int c = 0; c < NC; c++) if (c < nc) { const double t = fabs(v[c]); if (t > r || isnan(t)) r = t; }
        return r;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) if (c < nc) r = r + v[c] * v[c];
    return sqrt(r);
}

// per-warp shared-memory footprint in doubles (besides the tournament level)
template <int D> __host__ __device__ constexpr int warp_stage_doubles() {
    // tabU[2][D][7] + tabJ[2][D][7] + fst[2][1+4D][NC] + hs[2][D] + box[2][2D] + qs[NC][3D+1] + lb[D] + ub[D]
    return 28 * D + 2 * (1 + 4 * D) * NC + 2 * D + 4 * D + NC * (3 * D + 1) + 2 * D;
}

template <int FAM, int D>
__global__ void __launch_bounds__(kBlock)
hcub_vec_kernel(const Params p) {
    using F = ib::Family<FAM, D>;
    constexpr int NQ = F::NQ;
    constexpr int NP = hc::gm_npoints(D);
    constexpr int RECW = hc::rec_width(D);
    constexpr int NAX = 1 + 4 * D;
    static_assert(NQ <= 3 * D + 1, "prepared parameters do not fit the staging area");
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
    const int ngroups = p.cap / kGroup;
    const int ncomp = p.ncomp, normkind = p.normkind;
    unsigned long long *L1 = reinterpret_cast<unsigned long long *>(sm) + (size_t)wib * ngroups;
    double *stage = sm + (size_t)kWarpsPerBlock * ngroups + (size_t)wib * warp_stage_doubles<D>();
    double *tabU = stage;                      // [2][D][7]
    double *tabJ = tabU + 14 * D;              // [2][D][7]
    double *fst = tabJ + 14 * D;               // [2][NAX][NC]
    double *hs = fst + 2 * NAX * NC;           // [2][D]
    double *box = hs + 2 * D;                  // [2][2D]
    double *qs = box + 4 * D;                  // [NC][NQ]
    double *lbs = qs + NC * (3 * D + 1);       // [D]
    double *ubs = lbs + D;                     // [D]
    double *wtab = sm + (size_t)kWarpsPerBlock * ngroups + (size_t)kWarpsPerBlock * warp_stage_doubles<D>();  // [10]
    unsigned *pt = reinterpret_cast<unsigned *>(wtab + 10);
    double *keys = p.keys + (size_t)gwarp * p.cap;
    double *ivals = p.ivals + (size_t)gwarp * p.cap * ncomp;
    double *recs = p.recs + (size_t)gwarp * p.cap * RECW;

    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
    if (D >= 2) {
        // Genz-Malik weights and point table, as hcubature_impl.cuh
        if (threadIdx.x == 0) {
            const double twon = (double)(1 << D);
            wtab[0] = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0);
            wtab[1] = twon * (980 / 6561.0);
            wtab[2] = twon * ((1820 - 400 * D) / 19683.0);
            wtab[3] = twon * (200 / 19683.0);
            wtab[4] = 6859 / 19683.0;
            wtab[5] = twon * ((729 - 950 * D + 50 * D * D) / 729.0);
            wtab[6] = twon * (245 / 486.0);
            wtab[7] = twon * ((265 - 100 * D) / 1458.0);
            wtab[8] = twon * (25 / 729.0);
            wtab[9] = 0.0;
        }
        for (int q = threadIdx.x; q < NP; q += kBlock) {
            unsigned e = 0, cls = 0;
            if (q == 0) { cls = 0; }
            else if (q < 1 + 2 * D) { const int r = q - 1; cls = 1; e = (1u + (r & 1)) << (3 * (r >> 1)); }
            else if (q < 1 + 4 * D) { const int r = q - 1 - 2 * D; cls = 2; e = (3u + (r & 1)) << (3 * (r >> 1)); }
            else if (q < 1 + 4 * D + 2 * D * (D - 1)) {
                int r = q - 1 - 4 * D; const int s = r & 3; r >>= 2;
                int i = 0, j = 1;
                for (int cnt = 0; cnt < r; cnt++) { if (++j == D) { ++i; j = i + 1; } }
                cls = 3;
                e = ((3u + ((s >> 1) & 1)) << (3 * i)) | ((3u + (s & 1)) << (3 * j));
            } else {
                const unsigned r = q - (1 + 4 * D + 2 * D * (D - 1)); cls = 4;
                for (int k = 0; k < D; k++) e |= (5u + ((r >> k) & 1u)) << (3 * k);
            }
            pt[q] = e | (cls << 24);
        }
    }
    __syncthreads();

    double rtol = p.rtol;
    const double atol = p.atol;
    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps), HCubature default

    for (;;) {
        unsigned long long pidx = 0;
        if (lane == 0) pidx = atomicAdd(p.counter, 1ULL);
        pidx = __shfl_sync(0xffffffffu, pidx, 0);
        if ((int64_t)pidx >= p.nprob) break;
        const int64_t prob = (int64_t)pidx;

        // ---- problem data -> shared: prepared parameters of every component, bounds
        __syncwarp();
        if (p.sens.n > 0) {        // sensitivities: one RAW parameter column per problem
            for (int t = lane; t < p.nparam; t += 32) qs[t] = p.params[(size_t)prob * p.nparam + t];
        } else {
            for (int c = lane; c < ncomp; c += 32) {
                double q[NQ];
                F::prepare(p.params + ((size_t)prob * ncomp + c) * p.nparam, q);
#pragma unroll
                for (int t = 0; t < NQ; t++) qs[c * NQ + t] = q[t];
            }
        }
        if (lane < D) {
            lbs[lane] = p.bpp ? p.lb[prob * D + lane] : p.lb[lane];
            ubs[lane] = p.bpp ? p.ub[prob * D + lane] : p.ub[lane];
        }
        __syncwarp();

        // t -> (u, jac) on axis k (infinity_handling.jl maps; a finite axis is the affine map)
        auto map_axis = [&](int k, double t, double &u, double &jac) {
            ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m); ib::t2ujac(p.tr, t, m, u, jac);
        };
        // every component at the point x, times the Jacobian
        auto eval_all = [&](const double (&x)[D], double jac, double (&f)[NC]) {
            if constexpr (ib::Grad<FAM, D>::ok) {
                if (p.sens.n > 0) {
                    double go[1 + ib::fam_nparam(FAM, D)];
                    ib::Grad<FAM, D>::run(x, qs, go);
                    f[0] = go[0] * jac;
#pragma unroll
                    for (int c = 1; c < NC; c++) f[c] = c < ncomp ? go[1 + p.sens.idx[c - 1]] * jac : 0.0;
                    return;
                }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                f[c] = 0.0;
                if (c < ncomp) {
                    double q[NQ];
#pragma unroll
                    for (int t = 0; t < NQ; t++) q[t] = qs[c * NQ + t];
                    f[c] = F::eval(x, q) * jac;
                }
            }
        };

        // rule on the two boxes staged in box[0], box[1]; results uniform in all lanes
        auto rule2 = [&](bool second, double (&I0)[NC], double &E0, int &k0, double (&I1)[NC], double &E1, int &k1) {
            if constexpr (D == 1) {
                // HCubature's 1-D rule: GK(7,15) about the centre with signed half-width (oracle gk_rule_hc_vec)
                const int ch = lane >> 4, j = lane & 15;
                const double a = box[2 * ch], b = box[2 * ch + 1];
                const double c = (a + b) * 0.5, Dl = (b - a) * 0.5;
                double g[NC];
#pragma unroll
                for (int cc = 0; cc < NC; cc++) g[cc] = 0.0;
                if (j < 15 && (second || ch == 0)) {
                    const double dx = j < 14 ? Dl * hc::c_x[j < 7 ? j : j - 7] : 0.0;
                    const double t = j < 7 ? c + dx : (j < 14 ? c - dx : c);
                    double x[D], jac;
                    map_axis(0, t, x[0], jac);
                    eval_all(x, jac, g);
                }
                const double wk = j < 7 ? hc::c_w[j] : (j == 14 ? hc::c_w[7] : 0.0);
                const double wg = (j < 7 && (j & 1)) ? hc::c_wg[j >> 1] : (j == 14 ? hc::c_wg[3] : 0.0);
                double Is[NC], dd[NC];
#pragma unroll
                for (int cc = 0; cc < NC; cc++) {
                    Is[cc] = 0.0; dd[cc] = 0.0;
                    if (cc < ncomp) {
                        const double partner = __shfl_down_sync(0xffffffffu, g[cc], 7);
                        const double v = j < 7 ? g[cc] + partner : g[cc];
                        double ck = v * wk, cg = v * wg;
#pragma unroll
                        for (int off = 8; off > 0; off >>= 1) {
                            ck += __shfl_xor_sync(0xffffffffu, ck, off);
                            cg += __shfl_xor_sync(0xffffffffu, cg, off);
                        }
                        Is[cc] = ck * Dl; dd[cc] = ck * Dl - cg * Dl;
                    }
                }
                const double Er = vnorm(dd, ncomp, normkind);
#pragma unroll
                for (int cc = 0; cc < NC; cc++) {
                    I0[cc] = __shfl_sync(0xffffffffu, Is[cc], 0);
                    I1[cc] = __shfl_sync(0xffffffffu, Is[cc], 16);
                }
                E0 = __shfl_sync(0xffffffffu, Er, 0); E1 = __shfl_sync(0xffffffffu, Er, 16);
                k0 = 0; k1 = 0;
            } else {
                // phase 1: transformed coordinate and Jacobian of the 7 abscissae of every axis  idx = (ch*D + k)*7 + o
                const int ntask = second ? 14 * D : 7 * D;
                for (int idx = lane; idx < ntask; idx += 32) {
                    const int ch = idx >= 7 * D;
                    const int r = idx - ch * 7 * D;
                    const int k = r / 7, o = r - 7 * k;
                    const double a = box[ch * 2 * D + k], b = box[ch * 2 * D + D + k];
                    const double c = 0.5 * (a + b), h = 0.5 * fabs(b - a);
                    const double lam = o == 0 ? 0.0 : (o <= 2 ? l2 : (o <= 4 ? l3 : l5));
                    const double dl = h * lam;
                    const double t = o == 0 ? c : ((o & 1) ? c + dl : c - dl);
                    double u, jac;
                    map_axis(k, t, u, jac);
                    tabU[idx] = u; tabJ[idx] = jac;
                    if (o == 0) hs[ch * D + k] = h;
                }
                __syncwarp();
                // phase 2: the points; every lane evaluates all components at its points
                double sI0[NC], sP0[NC], sI1[NC], sP1[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) { sI0[c] = 0.0; sP0[c] = 0.0; sI1[c] = 0.0; sP1[c] = 0.0; }
                const int total = second ? 2 * NP : NP;
                for (int w = lane; w < total; w += 32) {
                    const int ch = w >= NP;
                    const int q = w - ch * NP;
                    const unsigned e = pt[q];
                    const double *tU = tabU + ch * 7 * D;
                    const double *tJ = tabJ + ch * 7 * D;
                    double x[D], jac = 1.0;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const int sel = (e >> (3 * k)) & 7u;
                        x[k] = tU[7 * k + sel];
                        const double J = tJ[7 * k + sel];
                        jac = (k == 0) ? J : jac * J;
                    }
                    double f[NC];
                    eval_all(x, jac, f);
                    const unsigned cls = e >> 24;
                    const double wI = wtab[cls], wP = wtab[5 + cls];
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        if (c < ncomp) {
                            if (q < NAX) fst[(ch * NAX + q) * NC + c] = f[c];
                            if (ch) { sI1[c] = fma(wI, f[c], sI1[c]); sP1[c] = fma(wP, f[c], sP1[c]); }
                            else    { sI0[c] = fma(wI, f[c], sI0[c]); sP0[c] = fma(wP, f[c], sP0[c]); }
                        }
                    }
                }
                // warp sums of both rules for both boxes (uniform in all lanes afterwards)
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    if (c < ncomp) {
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            sI0[c] += __shfl_xor_sync(0xffffffffu, sI0[c], off);
                            sP0[c] += __shfl_xor_sync(0xffffffffu, sP0[c], off);
                            sI1[c] += __shfl_xor_sync(0xffffffffu, sI1[c], off);
                            sP1[c] += __shfl_xor_sync(0xffffffffu, sP1[c], off);
                        }
                    }
                }
                __syncwarp();
                // volume, I, E and split axis of a box (largest fourth-difference norm, HCubature's tie fuzz); every lane
                // computes both boxes from uniform data
                auto finish_box = [&](int ch, const double (&kI)[NC], const double (&kP)[NC], double (&Iv)[NC], double &Er, int &kd) {
                    const double *hh = hs + ch * D;
                    const double *ff = fst + ch * NAX * NC;
                    double V = hh[0];
#pragma unroll
                    for (int k = 1; k < D; k++) V *= hh[k];
                    double dd[NC];
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        Iv[c] = 0.0; dd[c] = 0.0;
                        if (c < ncomp) { Iv[c] = V * kI[c]; dd[c] = Iv[c] - V * kP[c]; }
                    }
                    Er = vnorm(dd, ncomp, normkind);
                    double p10 = 1.0;
#pragma unroll
                    for (int k = 0; k < D; k++) p10 *= 10.0;
                    const double deltaf = Er / (p10 * V);
                    kd = 0; double maxdd = 0.0, hkd = hh[0];
#pragma unroll
                    for (int k = 0; k < D; k++) {
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            dd[c] = 0.0;
                            if (c < ncomp) {
                                const double f2k = ff[(1 + 2 * k) * NC + c] + ff[(2 + 2 * k) * NC + c];
                                const double f3k = ff[(1 + 2 * D + 2 * k) * NC + c] + ff[(2 + 2 * D + 2 * k) * NC + c];
                                dd[c] = f3k + 12 * ff[c] - 7 * f2k;
                            }
                        }
                        const double dn = vnorm(dd, ncomp, normkind);
                        const double delta = dn - maxdd;
                        if (delta > deltaf) { kd = k; maxdd = dn; hkd = hh[k]; }
                        else if (fabs(delta) <= deltaf && hh[k] > hkd) { kd = k; hkd = hh[k]; }
                    }
                };
                finish_box(0, sI0, sP0, I0, E0, k0);
                if (second) finish_box(1, sI1, sP1, I1, E1, k1);
                else {
#pragma unroll
                    for (int c = 0; c < NC; c++) I1[c] = 0.0;
                    E1 = 0.0; k1 = 0;
                }
                __syncwarp();
            }
        };

        // ---- pool helpers (uniform control flow)
        int nbox = 0;
        auto write_box = [&](int slot, int ch, const double (&Ib)[NC], double E, int kd) {
            if (lane < RECW) recs[(size_t)slot * RECW + lane] = lane < 2 * D ? box[ch * 2 * D + lane] : (double)kd;
            if (lane == 0) {
                keys[slot] = E;
#pragma unroll
                for (int c = 0; c < NC; c++) if (c < ncomp) ivals[(size_t)slot * ncomp + c] = Ib[c];
            }
        };
        auto append = [&](int ch, const double (&Ib)[NC], double E, int kd) {
            const int slot = nbox;
            write_box(slot, ch, Ib, E, kd);
            if (lane == 0) {
                const int g = slot / kGroup;
                const unsigned long long k = hc::dkey(E);
                L1[g] = (slot % kGroup == 0) ? k : (L1[g] > k ? L1[g] : k);
            }
            nbox++;
            __syncwarp();
        };

        // ---- initial box(es): HCubature's initdiv grid, CartesianIndices order (first index fastest)
        double I[NC], E = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c++) I[c] = 0.0;
        int64_t nev = 0;
        int status = IB200_ST_CONVERGED;
        bool degenerate = false;
        {
            const int initdiv = p.initdiv;
            int64_t total = 1;
            double prodD = 1.0;
#pragma unroll
            for (int k = 0; k < D; k++) {
                total *= initdiv;
                ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m);
                prodD *= (m.thi - m.tlo) / initdiv;
            }
            degenerate = (prodD == 0.0);
            if (degenerate) total = 1;              // HCubature returns after the first box
            for (int64_t q = 0; q < total; q += 2) {
                const bool second = q + 1 < total;
                if (lane < 2 * D) {                 // lane -> (box, axis)
                    const int ch = lane >= D, k = lane - ch * D;
                    int64_t r = q + ch;
                    for (int kk = 0; kk < k; kk++) r /= initdiv;
                    const int c = (int)(r % initdiv);
                    ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m);
                    const double Dl = (m.thi - m.tlo) / initdiv;
                    const double a = m.tlo + c * Dl;
                    const double b = (c == initdiv - 1) ? m.thi : a + Dl;
                    box[ch * 2 * D + k] = a; box[ch * 2 * D + D + k] = b;
                }
                __syncwarp();
                double I0[NC], I1[NC], E0, E1; int k0, k1;
                rule2(second, I0, E0, k0, I1, E1, k1);
                if (q == 0) {
#pragma unroll
                    for (int c = 0; c < NC; c++) I[c] = I0[c];
                    E = E0;
                } else {
#pragma unroll
                    for (int c = 0; c < NC; c++) I[c] += I0[c];
                    E += E0;
                }
                nev += NP;
                append(0, I0, E0, k0);
                if (second) {
#pragma unroll
                    for (int c = 0; c < NC; c++) I[c] += I1[c];
                    E += E1; nev += NP;
                    append(1, I1, E1, k1);
                }
            }
        }

        // ---- adaptive loop
        if (!degenerate) {
            for (;;) {
                if (E <= fmax(rtol * vnorm(I, ncomp, normkind), atol)) break;
                if (nev >= p.maxevals) { status = IB200_ST_MAXEVALS; break; }
                if (!isfinite(E)) { status = IB200_ST_NONFINITE; break; }
                if (nbox + 1 > p.cap) { status = IB200_ST_MAXEVALS; break; }
                // select: group maxima (shared), then the group's 64 keys (global)
                const int n1 = (nbox + kGroup - 1) / kGroup;
                unsigned long long bk = 0ULL; int bi = 0x7fffffff;
                for (int g = lane; g < n1; g += 32) {
                    const unsigned long long k = L1[g];
                    if (bi == 0x7fffffff || k > bk) { bk = k; bi = g; }
                }
                const int g = hc::warp_argmax(bk, bi);
                const int s_lo = g * kGroup + lane, s_hi = s_lo + 32;
                const unsigned long long klo = s_lo < nbox ? hc::dkey(keys[s_lo]) : 0ULL;
                const unsigned long long khi = s_hi < nbox ? hc::dkey(keys[s_hi]) : 0ULL;
                unsigned long long kk = klo; int ki = s_lo < nbox ? s_lo : 0x7fffffff;
                if (s_hi < nbox && khi > kk) { kk = khi; ki = s_hi; }
                const int slot = hc::warp_argmax(kk, ki);
                const double Ebox = __longlong_as_double((long long)hc::warp_max_key(ki == slot ? kk : 0ULL));
                double Ibox[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) Ibox[c] = c < ncomp ? ivals[(size_t)slot * ncomp + c] : 0.0;
                double rv = 0.0;
                if (lane < RECW) rv = recs[(size_t)slot * RECW + lane];
                const int kd = (int)__shfl_sync(0xffffffffu, rv, 2 * D);
                // children: child 0 = upper half (a[kd] += w), child 1 = lower half (b[kd] -= w)
                {
                    const double akd = __shfl_sync(0xffffffffu, rv, kd), bkd = __shfl_sync(0xffffffffu, rv, D + kd);
                    const double w = (bkd - akd) * 0.5;
                    if (lane < 2 * D) {
                        box[lane] = (lane == kd) ? akd + w : rv;
                        box[2 * D + lane] = (lane == D + kd) ? bkd - w : rv;
                    }
                }
                __syncwarp();
                double I1[NC], I2[NC], E1, E2; int k1, k2;
                rule2(true, I1, E1, k1, I2, E2, k2);
                nev += 2 * NP;
#pragma unroll
                for (int c = 0; c < NC; c++) I[c] += I1[c] + I2[c] - Ibox[c];
                E += E1 + E2 - Ebox;
                // child 0 replaces the parent's slot: recompute the group maximum from the keys in registers
                write_box(slot, 0, I1, E1, k1);
                {
                    const unsigned long long e1 = hc::dkey(E1);
                    const unsigned long long nlo = (s_lo == slot) ? e1 : klo, nhi = (s_hi == slot) ? e1 : khi;
                    const unsigned long long gm = hc::warp_max_key(nlo > nhi ? nlo : nhi);
                    if (lane == 0) L1[g] = gm;
                }
                __syncwarp();
                append(1, I2, E2, k2);
            }
        }

        // ---- re-sum over all boxes (HCubature's "roundoff paranoia"): lane-strided, then xor tree
        {
            double pI[NC], pE = 0.0;
#pragma unroll
            for (int c = 0; c < NC; c++) pI[c] = 0.0;
            for (int s = lane; s < nbox; s += 32) {
                pE += keys[s];
#pragma unroll
                for (int c = 0; c < NC; c++) if (c < ncomp) pI[c] += ivals[(size_t)s * ncomp + c];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pE += __shfl_xor_sync(0xffffffffu, pE, off);
#pragma unroll
                for (int c = 0; c < NC; c++) pI[c] += __shfl_xor_sync(0xffffffffu, pI[c], off);
            }
#pragma unroll
            for (int c = 0; c < NC; c++) I[c] = pI[c];
            E = pE;
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < NC; c++) if (c < ncomp) p.out_I[prob * ncomp + c] = I[c];
            if (p.out_E) p.out_E[prob] = E;
            if (p.out_nevals) p.out_nevals[prob] = nev;
            if (p.out_status) p.out_status[prob] = status;
        }
        __syncwarp();
    }
}

template <int D>
inline size_t smem_bytes(int cap) {
    return ((size_t)kWarpsPerBlock * (cap / kGroup) + (size_t)kWarpsPerBlock * warp_stage_doubles<D>() + 10) * sizeof(double) +
           (size_t)hc::gm_npoints(D) * sizeof(unsigned) + 16;
}

// launcher for one dimension; instantiated in hcubature_vec_d<D>.cu
template <int D>
int32_t launch(ib200_ctx *ctx, int family, Params &p);

template <int D>
int32_t launch_impl(ib200_ctx *ctx, int family, Params &p) {
    constexpr int RECW = hc::rec_width(D);
    const size_t smem = smem_bytes<D>(p.cap);
    if (smem > ctx->smem_optin) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_hcubature_vec_batch: box capacity too large for shared memory");
    int32_t rc = IB200_ERR_UNSUPPORTED;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, D)) {
            auto go = [&](auto kern) -> int32_t {
                IB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int occ = 0;
                IB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem));
                if (occ < 1) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_hcubature_vec_batch: kernel does not fit on an SM");
                int64_t blocks = (p.nprob + kWarpsPerBlock - 1) / kWarpsPerBlock;
                if (blocks > (int64_t)ctx->sm_count * occ) blocks = (int64_t)ctx->sm_count * occ;
                const size_t warps = (size_t)blocks * kWarpsPerBlock;
                p.keys = (double *)ctx->get(SL_WORK0, warps * p.cap * sizeof(double));
                p.ivals = (double *)ctx->get(SL_WORK3, warps * p.cap * p.ncomp * sizeof(double));
                p.recs = (double *)ctx->get(SL_WORK2, warps * p.cap * RECW * sizeof(double));
                if (!p.keys || !p.ivals || !p.recs)
                    return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_vec_batch: box pool allocation failed (lower maxiters or the hcub_capacity option)");
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(p);
                IB_LAUNCH_CHECK(ctx);
                return IB200_OK;
            };
            rc = go(hcub_vec_kernel<FAM, D>);
        }
    });
    if (rc == IB200_ERR_UNSUPPORTED && ctx->err.empty())
        return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_vec_batch: family not compiled in for this dimension");
    return rc;
}

}  // namespace hv
