// hcubature_impl.cuh -- batched adaptive cubature (Genz-Malik degree 7/5, or GK(7,15) in 1-D),
// one warp per problem.  Included by hcubature_d<D>.cu, one translation unit per dimension.
//
// Replaces HCubature.hcubature / hquadrature(f, lb, ub; rtol, atol, maxevals, initdiv) as called from
// __solvebp_call(::HCubatureJL) (src/Integrals.jl:347-397, call sites :380-393) including the
// automatic ChangeOfVariables in front of it (src/common.jl:97-103).  Algorithm (HCubature.jl,
// restated in oracle/oracle.c orc_hcubature / gm_rule): rule on the initial box(es); until
// E <= max(rtol*|I|, atol) or evals >= maxevals or E non-finite: take the box with the largest
// error, halve it along the axis of largest fourth difference, apply the rule to both halves,
// update I and E by difference; finally re-sum I and E over all boxes.
//
// B200 mapping
//   * persistent warps pull problems from a global atomic counter;
//   * the 2*NP points of the two child boxes are spread over the 32 lanes (4-D: 114 points in 4
//     passes); every lane keeps weighted partial sums of both rules, the warp combines them with
//     xor-shuffles; the 1+4D axis values needed for the fourth differences go through a small
//     shared-memory stage;
//   * the reference's binary heap is replaced by a two-level tournament per warp: the error keys of
//     all boxes live in global memory in groups of 64, the maximum of every group lives in shared
//     memory.  Selecting the worst box = strided scan of the shared level + REDUX arg-max, one
//     coalesced 512-byte key-group load + REDUX arg-max, one broadcast record load.  Child 1
//     overwrites the parent's slot, child 2 is appended.  Same selection order as the reference
//     (largest error first; ties -> lowest slot).
#pragma once
#include "common.cuh"
#include "family.cuh"

namespace hc {

constexpr int kWarpsPerBlock = 4;
constexpr int kBlock = 32 * kWarpsPerBlock;
constexpr int kGroup = 64;   // boxes per tournament group

struct Params {
    const double *params; int nparam; int64_t nprob;
    const double *lb, *ub; int bpp, tr;
    double rtol, atol; int64_t maxevals; int initdiv;
    int cap;                    // boxes per warp (multiple of kGroup)
    double *keys;               // [warps][cap]
    double *recs;               // [warps][cap][RECW]
    unsigned long long *counter;
    double *out_I, *out_E; int64_t *out_nevals; int32_t *out_status;
};

__host__ __device__ constexpr int gm_npoints(int d) { return d == 1 ? 15 : 1 + 4 * d + 2 * d * (d - 1) + (1 << d); }

__constant__ double c_x[8] = {
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.0};
__constant__ double c_w[8] = {
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_wg[4] = {
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

// warp arg-max of non-negative double keys (as ordered 64-bit patterns); ties -> smallest idx.
// Lanes without a candidate pass idx = 0x7fffffff.  Returns the winning idx (uniform).
__device__ __forceinline__ int warp_argmax(unsigned long long key, int idx) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const bool valid = idx != 0x7fffffff;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, valid ? hi : 0u);
    const bool c1 = valid && hi == mhi;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
    const bool win = c1 && lo == mlo;
    return (int)__reduce_min_sync(0xffffffffu, win ? (unsigned)idx : 0x7fffffffu);
}
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ unsigned long long dkey(double e) { return (unsigned long long)__double_as_longlong(e); }

template <int FAM, int D, bool ALLFIN>
__global__ void __launch_bounds__(kBlock)
hcub_kernel(const Params p) {
    constexpr int NP = gm_npoints(D);
    constexpr int RECW = 2 * D + 3;            // a[D], b[D], I, E, kdiv
    constexpr int NAX = 1 + 4 * D;             // centre + lambda2/lambda3 axis points (D >= 2)
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
    const int ngroups = p.cap / kGroup;
    // shared layout: [warps][ngroups] group maxima ; [warps][2][NAX] axis-value stage ; point table
    unsigned long long *L1 = reinterpret_cast<unsigned long long *>(sm) + (size_t)wib * ngroups;
    double *fst = sm + (size_t)kWarpsPerBlock * ngroups + (size_t)wib * 2 * NAX;
    unsigned *pt = reinterpret_cast<unsigned *>(sm + (size_t)kWarpsPerBlock * ngroups + (size_t)kWarpsPerBlock * 2 * NAX);
    double *keys = p.keys + (size_t)gwarp * p.cap;
    double *recs = p.recs + (size_t)gwarp * p.cap * RECW;

    // ---- Genz-Malik constants (HCubature.jl genz-malik.jl; oracle gm_rule)
    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
    const double twon = (double)(1 << D);
    double wcls[5], wpcls[5];
    wcls[0] = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0);
    wcls[1] = twon * (980 / 6561.0);
    wcls[2] = twon * ((1820 - 400 * D) / 19683.0);
    wcls[3] = twon * (200 / 19683.0);
    wcls[4] = 6859 / 19683.0;
    wpcls[0] = twon * ((729 - 950 * D + 50 * D * D) / 729.0);
    wpcls[1] = twon * (245 / 486.0);
    wpcls[2] = twon * ((265 - 100 * D) / 1458.0);
    wpcls[3] = twon * (25 / 729.0);
    wpcls[4] = 0.0;

    // ---- point table: bits 0-7 axes with an offset, 8-15 negative-sign axes, 16-18 class
    if (D >= 2) {
        for (int q = threadIdx.x; q < NP; q += kBlock) {
            unsigned m = 0, sg = 0, cls = 0;
            if (q == 0) { cls = 0; }
            else if (q < 1 + 2 * D) { const int r = q - 1; cls = 1; m = 1u << (r >> 1); sg = (r & 1) ? m : 0; }
            else if (q < 1 + 4 * D) { const int r = q - 1 - 2 * D; cls = 2; m = 1u << (r >> 1); sg = (r & 1) ? m : 0; }
            else if (q < 1 + 4 * D + 2 * D * (D - 1)) {
                int r = q - 1 - 4 * D; const int s = r & 3; r >>= 2;
                int i = 0, j = 1;
                for (int cnt = 0; cnt < r; cnt++) { if (++j == D) { ++i; j = i + 1; } }
                cls = 3; m = (1u << i) | (1u << j);
                sg = ((s & 2) ? (1u << i) : 0) | ((s & 1) ? (1u << j) : 0);
            } else { const unsigned r = q - (1 + 4 * D + 2 * D * (D - 1)); cls = 4; m = (1u << D) - 1; sg = r; }
            pt[q] = m | (sg << 8) | (cls << 16);
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

        ib::Problem<FAM, D, ALLFIN> P;
        const double *lbp = p.bpp ? p.lb + prob * D : p.lb;
        const double *ubp = p.bpp ? p.ub + prob * D : p.ub;
        P.load(p.params + prob * p.nparam, lbp, ubp, p.tr);

        // rule on two boxes: (A0,B0) and (A1,B1); results uniform in all lanes
        auto rule2 = [&](const double (&A0)[D], const double (&B0)[D], const double (&A1)[D], const double (&B1)[D],
                         bool second, double &I0, double &E0, int &k0, double &I1, double &E1, int &k1) {
            if constexpr (D == 1) {
                // HCubature's 1-D rule: GK(7,15) about the centre with signed half-width
                const bool hi = lane >= 16;
                const int j = lane & 15;
                const double a = hi ? A1[0] : A0[0], b = hi ? B1[0] : B0[0];
                const double c = (a + b) * 0.5, Dl = (b - a) * 0.5;
                double g = 0.0;
                if (j < 15 && (second || !hi)) {
                    const double dx = j < 14 ? Dl * c_x[j < 7 ? j : j - 7] : 0.0;
                    double t[1] = {j < 7 ? c + dx : (j < 14 ? c - dx : c)};
                    g = P(t);
                }
                const double partner = __shfl_down_sync(0xffffffffu, g, 7);
                const double v = j < 7 ? g + partner : g;
                double ck = v * (j < 7 ? c_w[j] : (j == 14 ? c_w[7] : 0.0));
                double cg = v * ((j < 7 && (j & 1)) ? c_wg[j >> 1] : (j == 14 ? c_wg[3] : 0.0));
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    ck += __shfl_xor_sync(0xffffffffu, ck, off);
                    cg += __shfl_xor_sync(0xffffffffu, cg, off);
                }
                const double Ik = ck * Dl, Ig = cg * Dl, Er = fabs(Ik - Ig);
                I0 = __shfl_sync(0xffffffffu, Ik, 0);  E0 = __shfl_sync(0xffffffffu, Er, 0);
                I1 = __shfl_sync(0xffffffffu, Ik, 16); E1 = __shfl_sync(0xffffffffu, Er, 16);
                k0 = 0; k1 = 0;
            } else {
                double sI[2] = {0.0, 0.0}, sP[2] = {0.0, 0.0};
                const int total = second ? 2 * NP : NP;
                for (int w = lane; w < total; w += 32) {
                    const int ch = w >= NP;
                    const int q = w - ch * NP;
                    const unsigned e = pt[q];
                    const unsigned cls = (e >> 16) & 7u;
                    const double lam = cls == 1 ? l2 : (cls == 4 ? l5 : l3);
                    double t[D];
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const double a = ch ? A1[k] : A0[k], b = ch ? B1[k] : B0[k];
                        const double c = 0.5 * (a + b);
                        double x = c;
                        if ((e >> k) & 1u) {
                            const double dl = (0.5 * fabs(b - a)) * lam;
                            x = ((e >> (8 + k)) & 1u) ? c - dl : c + dl;
                        }
                        t[k] = x;
                    }
                    const double f = P(t);
                    if (q < NAX) fst[ch * NAX + q] = f;
                    sI[ch] = fma(wcls[cls], f, sI[ch]);
                    sP[ch] = fma(wpcls[cls], f, sP[ch]);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    sI[0] += __shfl_xor_sync(0xffffffffu, sI[0], off);
                    sP[0] += __shfl_xor_sync(0xffffffffu, sP[0], off);
                    sI[1] += __shfl_xor_sync(0xffffffffu, sI[1], off);
                    sP[1] += __shfl_xor_sync(0xffffffffu, sP[1], off);
                }
                __syncwarp();
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    double V = 1.0;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const double dk = 0.5 * fabs((ch ? B1[k] : B0[k]) - (ch ? A1[k] : A0[k]));
                        V = (k == 0) ? dk : V * dk;
                    }
                    const double I = V * sI[ch], Ip = V * sP[ch];
                    const double Er = fabs(I - Ip);
                    // split axis: largest fourth difference with HCubature's tie fuzz
                    const double f1 = fst[ch * NAX];
                    const double twelvef1 = 12 * f1;
                    double p10 = 1.0;
#pragma unroll
                    for (int k = 0; k < D; k++) p10 *= 10.0;
                    const double deltaf = Er / (p10 * V);
                    int kd = 0; double maxdd = 0.0, dkd = 0.5 * fabs((ch ? B1[0] : B0[0]) - (ch ? A1[0] : A0[0]));
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const double f2k = fst[ch * NAX + 1 + 2 * k] + fst[ch * NAX + 2 + 2 * k];
                        const double f3k = fst[ch * NAX + 1 + 2 * D + 2 * k] + fst[ch * NAX + 2 + 2 * D + 2 * k];
                        const double dd = fabs(f3k + twelvef1 - 7 * f2k);
                        const double dk = 0.5 * fabs((ch ? B1[k] : B0[k]) - (ch ? A1[k] : A0[k]));
                        const double delta = dd - maxdd;
                        if (delta > deltaf) { kd = k; maxdd = dd; dkd = dk; }
                        else if (fabs(delta) <= deltaf && dk > dkd) { kd = k; dkd = dk; }
                    }
                    if (ch == 0) { I0 = I; E0 = Er; k0 = kd; } else { I1 = I; E1 = Er; k1 = kd; }
                }
                __syncwarp();
            }
        };

        // ---- pool helpers (uniform control flow; lane 0 writes)
        int nbox = 0;
        auto write_rec = [&](int slot, const double (&A)[D], const double (&B)[D], double I, double E, int kd) {
            if (lane == 0) {
                double *r = recs + (size_t)slot * RECW;
#pragma unroll
                for (int k = 0; k < D; k++) { r[k] = A[k]; r[D + k] = B[k]; }
                r[2 * D] = I; r[2 * D + 1] = E; r[2 * D + 2] = (double)kd;
                keys[slot] = E;
            }
        };
        auto append = [&](const double (&A)[D], const double (&B)[D], double I, double E, int kd) {
            const int slot = nbox;
            write_rec(slot, A, B, I, E, kd);
            if (lane == 0) {
                const int g = slot / kGroup;
                const unsigned long long k = dkey(E);
                L1[g] = (slot % kGroup == 0) ? k : (L1[g] > k ? L1[g] : k);
            }
            nbox++;
            __syncwarp();
        };

        // ---- initial box(es)
        double a0[D], b0[D];
        double tl[D], th[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            tl[k] = ALLFIN ? -1.0 : P.ax[ALLFIN ? 0 : k].tlo;
            th[k] = ALLFIN ? 1.0 : P.ax[ALLFIN ? 0 : k].thi;
        }
        double I = 0.0, E = 0.0;
        int64_t nev = 0;
        int status = IB200_ST_CONVERGED;
        bool degenerate = false;
        {
            const int initdiv = p.initdiv;
            double Dl[D]; double prodD = 1.0;
#pragma unroll
            for (int k = 0; k < D; k++) { Dl[k] = (th[k] - tl[k]) / initdiv; prodD *= Dl[k]; }
            int64_t total = 1;
#pragma unroll
            for (int k = 0; k < D; k++) total *= initdiv;
            if (prodD == 0.0) total = 1;           // HCubature returns after the first box
            for (int64_t q = 0; q < total; q += 2) {   // CartesianIndices order, first index fastest
                double A1[D], B1[D];
                const bool second = q + 1 < total;
                int64_t r0 = q, r1 = q + 1;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const int c0 = (int)(r0 % initdiv), c1 = (int)(r1 % initdiv);
                    r0 /= initdiv; r1 /= initdiv;
                    a0[k] = tl[k] + c0 * Dl[k];
                    b0[k] = (c0 == initdiv - 1) ? th[k] : a0[k] + Dl[k];
                    A1[k] = tl[k] + c1 * Dl[k];
                    B1[k] = (c1 == initdiv - 1) ? th[k] : A1[k] + Dl[k];
                }
                double I0, E0, I1, E1; int k0, k1;
                rule2(a0, b0, A1, B1, second, I0, E0, k0, I1, E1, k1);
                if (q == 0) { I = I0; E = E0; } else { I += I0; E += E0; }
                nev += NP;
                append(a0, b0, I0, E0, k0);
                if (second) { I += I1; E += E1; nev += NP; append(A1, B1, I1, E1, k1); }
            }
            degenerate = (prodD == 0.0);
        }

        // ---- adaptive loop
        if (!degenerate && status == IB200_ST_CONVERGED) {
            for (;;) {
                if (E <= fmax(rtol * fabs(I), atol)) break;
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
                const int g = warp_argmax(bk, bi);
                const int s_lo = g * kGroup + lane, s_hi = s_lo + 32;
                const unsigned long long klo = s_lo < nbox ? dkey(keys[s_lo]) : 0ULL;
                const unsigned long long khi = s_hi < nbox ? dkey(keys[s_hi]) : 0ULL;
                unsigned long long kk = klo; int ki = s_lo < nbox ? s_lo : 0x7fffffff;
                if (s_hi < nbox && khi > kk) { kk = khi; ki = s_hi; }
                const int slot = warp_argmax(kk, ki);
                // parent record (broadcast loads)
                const double *r = recs + (size_t)slot * RECW;
                double A[D], B[D];
#pragma unroll
                for (int k = 0; k < D; k++) { A[k] = r[k]; B[k] = r[D + k]; }
                const double Ibox = r[2 * D], Ebox = r[2 * D + 1];
                const int kd = (int)r[2 * D + 2];
                // children: upper half first (ma = a with a[k] += w), then lower half (mb = b with b[k] -= w)
                double A1[D], B2[D];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const double w = (B[k] - A[k]) * 0.5;
                    A1[k] = (k == kd) ? A[k] + w : A[k];
                    B2[k] = (k == kd) ? B[k] - w : B[k];
                }
                double I1, E1, I2, E2; int k1, k2;
                rule2(A1, B, A, B2, true, I1, E1, k1, I2, E2, k2);
                nev += 2 * NP;
                I += I1 + I2 - Ibox;
                E += E1 + E2 - Ebox;
                // child 1 replaces the parent's slot: recompute the group maximum from the keys in registers
                write_rec(slot, A1, B, I1, E1, k1);
                {
                    const unsigned long long e1 = dkey(E1);
                    const unsigned long long nlo = (s_lo == slot) ? e1 : klo, nhi = (s_hi == slot) ? e1 : khi;
                    const unsigned long long gm = warp_max_key(nlo > nhi ? nlo : nhi);
                    if (lane == 0) L1[g] = gm;
                }
                __syncwarp();
                append(A, B2, I2, E2, k2);
            }
        }

        // ---- re-sum over all boxes (HCubature's "roundoff paranoia"): lane-strided, then xor tree
        {
            double pI = 0.0, pE = 0.0;
            for (int s = lane; s < nbox; s += 32) {
                const double *r = recs + (size_t)s * RECW;
                pI += r[2 * D]; pE += r[2 * D + 1];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pI += __shfl_xor_sync(0xffffffffu, pI, off);
                pE += __shfl_xor_sync(0xffffffffu, pE, off);
            }
            I = pI; E = pE;
        }
        if (lane == 0) {
            p.out_I[prob] = I;
            if (p.out_E) p.out_E[prob] = E;
            if (p.out_nevals) p.out_nevals[prob] = nev;
            if (p.out_status) p.out_status[prob] = status;
        }
        __syncwarp();
    }
}

template <int D>
inline size_t smem_bytes(int cap) {
    constexpr int NAX = 1 + 4 * D;
    return ((size_t)kWarpsPerBlock * (cap / kGroup) + (size_t)kWarpsPerBlock * 2 * NAX) * sizeof(double) +
           (size_t)gm_npoints(D) * sizeof(unsigned) + 16;
}

// launcher for one dimension; instantiated in hcubature_d<D>.cu.  Sizes the persistent grid from the
// kernel's occupancy, carves the per-warp box pools out of the context's scratch and launches.
template <int D>
int32_t launch(ib200_ctx *ctx, int family, bool allfin, Params &p);

template <int D>
int32_t launch_impl(ib200_ctx *ctx, int family, bool allfin, Params &p) {
    constexpr int RECW = 2 * D + 3;
    const size_t smem = smem_bytes<D>(p.cap);
    if (smem > ctx->smem_optin) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_hcubature_batch: box capacity too large for shared memory");
    int32_t rc = IB200_ERR_UNSUPPORTED;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, D)) {
            auto go = [&](auto kern) -> int32_t {
                IB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int occ = 0;
                IB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem));
                if (occ < 1) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_hcubature_batch: kernel does not fit on an SM");
                if (ctx->opt_adaptive_warps_per_sm > 0) {
                    const int want = (int)((ctx->opt_adaptive_warps_per_sm + kWarpsPerBlock - 1) / kWarpsPerBlock);
                    if (want < occ) occ = want;
                }
                int64_t blocks = (p.nprob + kWarpsPerBlock - 1) / kWarpsPerBlock;
                if (blocks > (int64_t)ctx->sm_count * occ) blocks = (int64_t)ctx->sm_count * occ;
                const size_t warps = (size_t)blocks * kWarpsPerBlock;
                p.keys = (double *)ctx->get(SL_WORK0, warps * p.cap * sizeof(double));
                p.recs = (double *)ctx->get(SL_WORK2, warps * p.cap * RECW * sizeof(double));
                if (!p.keys || !p.recs) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: box pool allocation failed (lower maxiters or the hcub_capacity option)");
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(p);
                IB_LAUNCH_CHECK(ctx);
                return IB200_OK;
            };
            rc = allfin ? go(hcub_kernel<FAM, D, true>) : go(hcub_kernel<FAM, D, false>);
        }
    });
    if (rc == IB200_ERR_UNSUPPORTED && ctx->err.empty())
        return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_batch: family not compiled in for this dimension");
    return rc;
}

}  // namespace hc
