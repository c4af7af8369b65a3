// hcubature_rb.cuh -- batched adaptive Genz-Malik cubature, ROUND-BASED driver (ib200_hcubature_batch, mode 1).
//
// Same call as the other two drivers: HCubature.hcubature(f, lb, ub; rtol, atol, maxevals, initdiv) as called at
// src/Integrals.jl:380-393 under the automatic ChangeOfVariables (src/common.jl:97-103), for a batch of problems.
// The reference refines ONE box per iteration in strict largest-error order (a chain of 10^2 .. 4*10^4 dependent
// steps per 4-D problem, which is what bounds the reference-order drivers hcubature_pair.cuh / hcubature_impl.cuh).
// Here a problem is refined in ROUNDS, the scheme of ib200_hcubature_single (csrc/hcub_single.cu) applied per problem:
//
//   round:  evaluate   the rule on every box created by the previous round -- one THREAD per box, 32 boxes per pass
//           test       E <= max(rtol*|I|, atol) (HCubature's test) / budget of rule applications / non-finite
//           threshold  theta = min(half the largest error seen, theta_h): theta_h comes from the histogram of the
//                      error estimates by binary exponent (2048 integer bins in shared memory, kept up to date
//                      incrementally) -- every box outside the longest ascending-error prefix whose errors sum to <= tol
//                      is bisected by the reference's worst-first loop too; the budget and the per-round share of the
//                      store (hcub_single.cu: rules 2 and 3) cap the set from the other side
//           select     one coalesced pass over the error keys of the problem's store: warp ballot + prefix popcount
//                      compacts the slots with error >= theta into the split list (no atomics, deterministic order)
//           split      child 0 (upper half along the parent's fourth-difference axis) takes the parent's slot, child 1
//                      is appended at nbox + rank -- fused into the next round's evaluation pass
//
// One WARP owns one problem at a time (persistent grid, problems pulled from a global atomic counter); its box store
// (bounds, integral, error, split axis: SoA) lives in a per-warp arena in HBM that is reused from problem to problem,
// so device memory is sized by residency, not by the batch.  Everything a round needs is in registers / shared
// memory of that warp; there is no host round trip, no grid-wide synchronisation and no floating-point atomic:
// a problem's result is bit-identical whatever batch it is solved in.
//
// Selection differs from the reference (every box above a threshold instead of the single worst), so evaluation
// counts differ (measured on BASELINE config 4: +5 .. +19 % evaluations); parity is on (I, E): oracle
// orc_hcubature_rounds_batch (oracle/oracle.c) restates this scheme and refines identically, and both agree with the
// reference-order orc_hcubature within the requested tolerance (tests/test_gpu_rounds_batch.py).
#pragma once
#include "hcubature_pair.cuh"

namespace hb {

constexpr int kWarps = 4;
constexpr int kBlock = 32 * kWarps;
constexpr int kBins = 2048;                   // one bin per binary exponent of a double

struct Params {
    const double *params; int nparam; int64_t nprob;
    const double *lb, *ub; int bpp, tr;
    double rtol, atol;
    long long max_apps;         // budget of rule applications: ceil(maxevals / points per rule)
    int initdiv;
    double frac;                // largest share of a problem's stored boxes bisected in one round
    int cap;                    // boxes per warp arena
    double *bounds;             // [warps][cap][2D]   a[D], b[D]
    double *ival;               // [warps][cap]
    double *err;                // [warps][cap]
    unsigned char *kdiv;        // [warps][cap]       split axis
    int *list;                  // [warps][cap]       slots selected for bisection in the current round
    unsigned long long *counter;
    double *out_I, *out_E; int64_t *out_nevals; int32_t *out_status;
};

__device__ __forceinline__ int ebin(double e) { return (int)(((unsigned long long)__double_as_longlong(e) >> 52) & 0x7ffull); }
// lower edge of exponent bin b: 0 for b = 0, 2^(b-1023) otherwise (== scalbn(1.0, b - 1023) of the oracle)
__device__ __forceinline__ double bin_edge(int b) { return b <= 0 ? 0.0 : __longlong_as_double((long long)b << 52); }

// fixed-shape warp sums (same association whatever the data): results uniform in all lanes
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// per-warp shared-memory row: hist[kBins] (unsigned) | pc = par[NPAR] lbs[D] aux[D] finit jacc  (the layout gm_rule_thread reads)
template <int FAM, int D> __host__ __device__ constexpr int row_doubles() { return kBins / 2 + hp::pc_doubles<FAM, D>() + 1; }

#ifndef HB_MINB
#define HB_MINB 2
#endif
template <int FAM, int D, bool ALLFIN>
__global__ void __launch_bounds__(kBlock, HB_MINB)
hcub_rb_kernel(const Params p) {
    using F = ib::Family<FAM, D>;
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    constexpr int NP = hc::gm_npoints(D);
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + wib;
    unsigned *hist = reinterpret_cast<unsigned *>(sm + (size_t)wib * row_doubles<FAM, D>());
    double *pc = sm + (size_t)wib * row_doubles<FAM, D>() + kBins / 2;
    double *par = pc, *lbs = pc + NPAR, *aux = lbs + D;

    double *bounds = p.bounds + (size_t)gwarp * p.cap * (2 * D);
    double *ival = p.ival + (size_t)gwarp * p.cap;
    double *err = p.err + (size_t)gwarp * p.cap;
    unsigned char *kdiv = p.kdiv + (size_t)gwarp * p.cap;
    int *list = p.list + (size_t)gwarp * p.cap;

    double rtol = p.rtol;
    const double atol = p.atol;
    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps), HCubature default

    for (;;) {
        unsigned long long pidx = 0;
        if (lane == 0) pidx = atomicAdd(p.counter, 1ULL);
        pidx = __shfl_sync(full, pidx, 0);
        if ((int64_t)pidx >= p.nprob) break;
        const int64_t prob = (int64_t)pidx;

        // ---- problem data -> this warp's shared row; empty histogram
        __syncwarp();
        for (int t = lane; t < NPAR; t += 32) par[t] = p.params[prob * p.nparam + t];
        for (int k = lane; k < D; k += 32) {
            const double l = p.bpp ? p.lb[prob * D + k] : p.lb[k];
            const double u = p.bpp ? p.ub[prob * D + k] : p.ub[k];
            lbs[k] = l; aux[k] = ALLFIN ? (u - l) * 0.5 : u;
        }
        for (int b = lane; b < kBins; b += 32) hist[b] = 0u;
        __syncwarp();
        if (lane == 0) {
            double jacc = 1.0;
            if (ALLFIN) {
#pragma unroll
                for (int k = 0; k < D; k++) jacc = (k == 0) ? aux[0] : jacc * aux[k];
            }
            aux[D] = F::sep_init(par); aux[D + 1] = jacc;
        }
        __syncwarp();

        // ---- initial grid (HCubature initdiv; CartesianIndices order, first index fastest) -> slots 0 .. qtotal-1
        int qtotal = 1;
        double prodD = 1.0;
        double tl[D], th[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            tl[k] = -1.0; th[k] = 1.0;
            if (!ALLFIN) { ib::AxisMap m; ib::make_axis(lbs[k], aux[k], m); tl[k] = m.tlo; th[k] = m.thi; }
            qtotal *= p.initdiv;
            prodD *= (th[k] - tl[k]) / p.initdiv;
        }
        const bool degenerate = (prodD == 0.0);
        if (degenerate) qtotal = 1;                 // HCubature returns after the first box

        int nbox = qtotal, ntodo = qtotal, nsplit = 0, status = IB200_ST_CONVERGED;
        int blo = kBins, bhi = -1;                  // occupied range of the histogram (a superset: bins may empty again)
        long long applied = 0;
        double I = 0.0, E = 0.0, theta = 0.0;
        bool first = true;

        for (;;) {
            // ---- evaluate: first round = the initial grid, later rounds = both halves of every selected box
            double sI = 0.0, sE = 0.0, mx = 0.0;
            int mylo = kBins, myhi = -1;
            for (int t0 = 0; t0 < ntodo; t0 += 32) {
                const int t = t0 + lane;
                const bool have = t < ntodo;
                double a[D], b[D];
                int slot = 0;
                if (have) {
                    if (first) {
                        slot = t;
                        int r = t;
#pragma unroll
                        for (int k = 0; k < D; k++) {
                            const int cidx = r % p.initdiv; r /= p.initdiv;
                            const double Dl = (th[k] - tl[k]) / p.initdiv;
                            a[k] = tl[k] + cidx * Dl;
                            b[k] = (cidx == p.initdiv - 1) ? th[k] : a[k] + Dl;
                        }
                    } else {
                        const int parent = list[t >> 1];
                        const double2 *src = reinterpret_cast<const double2 *>(bounds + (size_t)parent * (2 * D));
                        double r[2 * D];
#pragma unroll
                        for (int q = 0; q < D; q++) { const double2 v = __ldcg(src + q); r[2 * q] = v.x; r[2 * q + 1] = v.y; }
                        const int kd = kdiv[parent];
                        double akd = 0.0, bkd = 0.0;
#pragma unroll
                        for (int k = 0; k < D; k++) { a[k] = r[k]; b[k] = r[D + k]; if (k == kd) { akd = r[k]; bkd = r[D + k]; } }
                        const double w = (bkd - akd) * 0.5;
                        // child 0 = upper half (a[kd] += w) keeps the parent's slot, child 1 = lower half (b[kd] -= w) is appended
#pragma unroll
                        for (int k = 0; k < D; k++) if (k == kd) { if ((t & 1) == 0) a[k] = akd + w; else b[k] = bkd - w; }
                        slot = (t & 1) ? nbox - nsplit + (t >> 1) : parent;
                    }
                }
                __syncwarp();            // both children have read the parent's record before child 0 overwrites it
                if (have) {
                    double Ic, Ec; int kc = 0;
                    if constexpr (FAM == IB200_FAM_GENZ_OSC && ALLFIN) hp::gm_rule_thread_osc<D>(pc, a, b, Ic, Ec, kc);
                    else hp::gm_rule_thread<FAM, D, ALLFIN>(pc, p.tr, a, b, Ic, Ec, kc);
                    double2 *dst = reinterpret_cast<double2 *>(bounds + (size_t)slot * (2 * D));
                    double r[2 * D];
#pragma unroll
                    for (int k = 0; k < D; k++) { r[k] = a[k]; r[D + k] = b[k]; }
#pragma unroll
                    for (int q = 0; q < D; q++) dst[q] = make_double2(r[2 * q], r[2 * q + 1]);
                    ival[slot] = Ic; err[slot] = Ec; kdiv[slot] = (unsigned char)kc;
                    const int bn = ebin(Ec);
                    atomicAdd(&hist[bn], 1u);
                    mylo = min(mylo, bn); myhi = max(myhi, bn);
                    sI += Ic; sE += Ec; mx = fmax(mx, Ec);       // fmax drops NaN: a non-finite E is caught through sE
                }
            }
            sI = wsum(sI); sE = wsum(sE); mx = wmax(mx);
            blo = min(blo, (int)__reduce_min_sync(full, (unsigned)mylo));
            bhi = max(bhi, (int)__reduce_max_sync(full, (unsigned)(myhi + 1)) - 1);
            I += sI; E += sE;
            applied += ntodo;
            first = false;
            __syncwarp();

            // ---- HCubature's stopping test
            const double tol = fmax(rtol * fabs(I), atol);
            if (degenerate || E <= tol) break;
            if (!isfinite(E)) { status = IB200_ST_NONFINITE; break; }
            if (applied >= p.max_apps) { status = IB200_ST_MAXEVALS; break; }

            // ---- threshold (csrc/hcub_rounds.cuh decide_kernel, per problem; every lane computes the same value)
            theta = 0.5 * fmax(mx, theta);
            {
                const int b0 = blo, b1 = bhi < kBins - 2 ? bhi : kBins - 2;
                int bh = b0 - 1; double low = 0.0; bool broke = false;
                for (int b = b0; b <= b1; b++) {
                    const unsigned c = hist[b];
                    if (c != 0u) { low += (double)c * bin_edge(b); if (!(low <= tol)) { broke = true; break; } }
                    bh = b;
                }
                if (!broke) bh = kBins - 2;
                const double remaining = (double)p.max_apps - (double)applied;
                int bb = b1 + 1; double above = 0.0; broke = false;
                for (int b = b1; b >= b0; b--) {
                    above += (double)hist[b];
                    if (2.0 * above > remaining) { broke = true; break; }
                    bb = b;
                }
                if (!broke) bb = 0;
                int bf = b1 + 1; double top = 0.0; broke = false;
                for (int b = b1; b >= b0; b--) {
                    top += (double)hist[b];
                    if (top > p.frac * (double)nbox) { broke = true; break; }
                    bf = b;
                }
                if (!broke) bf = 0;
                if (bf > bb) bb = bf;
                const int bs = (bh + 1 > bb) ? bh + 1 : bb;
                const double theta_h = bs <= 0 ? 0.0 : (bs >= kBins - 1 ? 1.7976931348623157e308 : bin_edge(bs));
                theta = fmin(theta, theta_h);
            }

            // ---- select: slots with error >= theta, in store order
            nsplit = 0;
            double dI = 0.0, dE = 0.0;
            for (int i0 = 0; i0 < nbox; i0 += 128) {
                double e[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const int i = i0 + 32 * u + lane; e[u] = i < nbox ? __ldcg(err + i) : -1.0; }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + 32 * u + lane;
                    const bool flag = e[u] >= theta;
                    const unsigned bal = __ballot_sync(full, flag);
                    if (flag) {
                        list[nsplit + __popc(bal & lt)] = i;
                        dI += __ldcg(ival + i); dE += e[u];
                        atomicSub(&hist[ebin(e[u])], 1u);
                    }
                    nsplit += __popc(bal);
                }
            }
            if (nbox + nsplit > p.cap) { status = IB200_ST_MAXEVALS; break; }     // arena full: leave the boxes alone
            dI = wsum(dI); dE = wsum(dE);
            I -= dI; E -= dE;
            nbox += nsplit; ntodo = 2 * nsplit;
            __syncwarp();                 // the split list is read across lanes
        }

        // ---- HCubature's closing sum over the store
        double pI = 0.0, pE = 0.0;
        for (int s = lane; s < nbox; s += 32) { pI += __ldcg(ival + s); pE += __ldcg(err + s); }
        pI = wsum(pI); pE = wsum(pE);
        if (lane == 0) {
            p.out_I[prob] = pI;
            if (p.out_E) p.out_E[prob] = pE;
            if (p.out_nevals) p.out_nevals[prob] = (int64_t)applied * NP;
            if (p.out_status) p.out_status[prob] = status;
        }
    }
}

inline size_t arena_bytes_per_box(int D) { return (size_t)16 * D + 8 + 8 + 1 + 4; }

template <int D>
int32_t launch(ib200_ctx *ctx, int family, bool allfin, Params &p);

template <int D>
int32_t launch_impl(ib200_ctx *ctx, int family, bool allfin, Params &p) {
    int32_t rc = IB200_ERR_UNSUPPORTED;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, D)) {
            auto go = [&](auto kern) -> int32_t {
                const size_t smem = (size_t)kWarps * row_doubles<FAM, D>() * sizeof(double);
                IB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int occ = 0;
                IB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem));
                if (occ < 1) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_hcubature_batch: kernel does not fit on an SM");
                if (ctx->opt_adaptive_warps_per_sm > 0) {
                    const int want = (int)((ctx->opt_adaptive_warps_per_sm + kWarps - 1) / kWarps);
                    if (want < occ) occ = want;
                }
                int64_t blocks = (p.nprob + kWarps - 1) / kWarps;
                if (blocks > (int64_t)ctx->sm_count * occ) blocks = (int64_t)ctx->sm_count * occ;
                const size_t warps = (size_t)blocks * kWarps, cap = (size_t)p.cap;
                p.bounds = (double *)ctx->get(SL_WORK2, warps * cap * 2 * D * sizeof(double));
                p.ival = (double *)ctx->get(SL_WORK0, warps * cap * sizeof(double));
                p.err = (double *)ctx->get(SL_WORK4, warps * cap * sizeof(double));
                p.list = (int *)ctx->get(SL_WORK3, warps * cap * sizeof(int));
                p.kdiv = (unsigned char *)ctx->get(SL_WORK5, warps * cap);
                if (!p.bounds || !p.ival || !p.err || !p.list || !p.kdiv)
                    return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: box arena allocation failed (lower maxiters or the hcub_capacity option)");
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(p);
                IB_LAUNCH_CHECK(ctx);
                return IB200_OK;
            };
            rc = allfin ? go(hcub_rb_kernel<FAM, D, true>) : go(hcub_rb_kernel<FAM, D, false>);
        }
    });
    if (rc == IB200_ERR_UNSUPPORTED && ctx->err.empty())
        return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_batch: family not compiled in for this dimension (round-based mode)");
    return rc;
}

}  // namespace hb
