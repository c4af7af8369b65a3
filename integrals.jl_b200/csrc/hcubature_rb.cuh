// hcubature_rb.cuh -- batched adaptive Genz-Malik cubature, ROUND-BASED driver (ib200_hcubature_batch, mode 1; and the second
// level of ib200_hcubature_single).
//
// Same call as the other two drivers: HCubature.hcubature(f, lb, ub; rtol, atol, maxevals, initdiv) as called at
// src/Integrals.jl:380-393 under the automatic ChangeOfVariables (src/common.jl:97-103), for a batch of problems.
// The reference refines ONE box per iteration in strict largest-error order (a chain of 10^2 .. 4*10^4 dependent
// steps per 4-D problem, which is what bounds the reference-order drivers hcubature_pair.cuh / hcubature_impl.cuh).
// Here a problem is refined in ROUNDS, the scheme of ib200_hcubature_single (csrc/hcub_single.cu) applied per problem:
//
//   round:  evaluate   the rule on every box created by the previous round -- 2 <= dim <= 5: one THREAD per box, 32 boxes
//                      per pass (hcubature_pair.cuh gm_rule_thread); dim >= 6: one box at a time by the whole WARP, the
//                      2^d + 2 d^2 + 2 d + 1 points over the lanes (hcub_rounds.cuh WarpRule)
//           test       E <= max(rtol*|I|, atol) (HCubature's test) / budget of rule applications / non-finite
//           threshold  theta = min(half the largest error seen, theta_h): theta_h comes from the histogram of the
//                      error estimates by binary exponent (2048 integer bins in shared memory, kept up to date
//                      incrementally, with a bitmap of the occupied bins) -- every box outside the longest ascending-error
//                      prefix whose errors sum to <= tol is bisected by the reference's worst-first loop too; the budget
//                      and the per-round share of the store (hcub_single.cu: rules 2 and 3) cap the set from the other side
//           select     one coalesced pass over the error keys of the problem's store: warp ballot + prefix popcount
//                      compacts the slots with error >= theta into the split list (no atomics, deterministic order)
//           split      child 0 (upper half along the parent's fourth-difference axis) takes the parent's slot, child 1
//                      is appended at nbox + rank -- fused into the next round's evaluation pass
//
// One WARP owns one problem at a time (persistent grid, problems pulled from a global atomic counter); its box store
// (records a, b, I, split axis + error keys) lives in a per-warp arena in HBM that is reused from problem to problem,
// so device memory is sized by residency, not by the batch.  Everything a round needs is in registers / shared
// memory of that warp; there is no host round trip, no grid-wide synchronisation and no floating-point atomic:
// a problem's result is bit-identical whatever batch it is solved in.
//
// SUB-PROBLEM MODE (Params::sub_rec != nullptr): the "problems" are the boxes of the level-1 store of ONE integral
// (ib200_hcubature_single, two-level solve).  Box i is refined by the same rounds until ITS error is below ITS share of the
// tolerance, eps_i = tol * E_i^alpha / sum_j E_j^alpha; its boxes are summed and dropped when it is done.  The store is dealt to
// the ranks in chunks of 64 consecutive boxes, round-robin; no communication.
//
// Selection differs from the reference (every box above a threshold instead of the single worst), so evaluation
// counts differ (measured on BASELINE config 4: +5 .. +19 % evaluations); parity is on (I, E): oracle
// orc_hcubature_rounds_batch / orc_hcubature_two_level (oracle/oracle.c) restate this scheme and refine identically, and both
// agree with the reference-order orc_hcubature within the requested tolerance (tests/test_gpu_rounds_batch.py).
#pragma once
#include "hcubature_pair.cuh"
#include "hcub_rounds.cuh"
#include "gm_rule_thread2.cuh"
#include "gm_rule_thread3.cuh"

namespace hb {

constexpr int kWarps = 4;
constexpr int kBlock = 32 * kWarps;
constexpr int kBins = 2048;                   // one bin per binary exponent of a double
constexpr int kMaxThreadDim = 5;              // thread-per-box rule up to this dimension; above it for the product families only
// warp per box: 6 .. 8 dimensions, families whose value is finish(sum of terms) (oscillatory, corner peak, exp(a.x)); the product
// families have a thread-per-box rule up to 8-D (gm_rule_thread3.cuh)
template <int FAM, int D> __host__ __device__ constexpr bool warp_box() { return D > kMaxThreadDim && !ib::Family<FAM, D>::kProd; }

struct Params {
    const double *params; int nparam; int64_t nprob;
    const double *lb, *ub; int bpp, tr;
    double rtol, atol;
    long long max_apps;         // budget of rule applications per problem: ceil(maxevals / points per rule)
    int initdiv;
    double frac;                // largest share of a problem's stored boxes bisected in one round
    int cap;                    // boxes per warp arena
    double *rec;                // [warps][cap][2D+2]   a[D], b[D], I, split axis  (16-byte aligned records)
    double *err;                // [warps][cap]         the selection keys, scanned once per round
    int *list;                  // [warps][cap]         slots selected for bisection in the current round
    unsigned long long *counter;    // [1] problem counter (+ [5] diagnostics in sub-problem mode: ended by budget / full arena / non-finite, sub-problems
                                    //     that retired boxes, boxes retired)
    double *out_I, *out_E; int64_t *out_nevals; int32_t *out_status;
    // sub-problem mode (level 2 of ib200_hcubature_single)
    const double *sub_rec;      // [n][2D+2] records of the level-1 store (every box evaluated), or nullptr: batch mode
    const double *sub_err;      // [n]
    double sub_tol, sub_wsum, sub_alpha;
    int sub_rank, sub_world;
    long long sub_n;            // boxes in the level-1 store.  Chunks of hs::kSubChunk consecutive boxes are dealt to the ranks round-robin
                                // (chunk c -> rank c mod world); nprob = this rank's chunks x kSubChunk
    double *sub_out;            // [nprob][4]: integral, error, rule applications, status of this rank's sub-problems
    // compact-record launches (CPT, below): problems whose boxes outgrow the compact form are handed back through ovf_list / ovf_count and
    // solved by a second launch of the full-record kernel, which takes its problems from order[0 .. *norder)
    const unsigned *order = nullptr; const unsigned long long *norder = nullptr;
    unsigned *ovf_list = nullptr; unsigned long long *ovf_count = nullptr;
    int cpt_max_level = 31;     // halvings per axis a compact record may hold (<= kCptMaxLevel; an option so that tests reach the hand-back)
};

// ---- COMPACT RECORDS (batch mode, finite domains, initdiv = 1, thread per box in 2..5 dimensions).  Every box of such a solve comes
//      from halving [-1, 1]^D, so its corners are dyadic: a_k = -1 + i_k 2^(1-l_k), b_k = a_k + 2^(1-l_k) with the level l_k = number of
//      halvings along axis k and i_k < 2^l_k.  Up to l_k = 31 these are EXACTLY the doubles the successive a + (b - a)/2, b - (b - a)/2 of
//      the full-record kernel (and of the reference) produce -- every intermediate is representable -- so a box can be stored as
//      (i_k, l_k) and decoded with one FMA per axis: 32 bytes = ONE sector instead of 80 bytes across 3.5 sectors in 4-D.  The 4-D
//      product-peak kernel moves with its DRAM bytes (4 TB/s of scattered 16..80-byte pieces; padding its records to 96 B cost 9 %), and
//      the record traffic was 60 % of them.
//      Layout: uint4 {i_0, i_1, i_2, i_3} | uint2 {meta, i_4} | double I;   meta = l_0 | l_1 << 5 | ... | l_4 << 20 | split axis << 28.
//      A problem that would halve a box a 32nd time along one axis is abandoned and listed in ovf_list (see Params).
constexpr int kCptBytes = 32;
constexpr int kCptMaxLevel = 31;
#ifndef HB_COMPACT
#define HB_COMPACT 1
#endif

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

// ---- the split threshold of a round (csrc/hcub_rounds.cuh decide_kernel, per problem): half the largest error seen, lowered to the bound
//      theta_h that the exponent histogram gives.  The three scans walk only the OCCUPIED bins (bitmap), in the same order and with the same
//      arithmetic as the scans over all bins (oracle: orc_hcubature_rounds); every lane computes the same value.  Bin kBins-1 (inf / NaN
//      keys) never takes part: a non-finite E has ended the solve before.  tol_live = the tolerance minus the error retired so far.
__device__ __forceinline__ double round_threshold(const unsigned *hist, const unsigned *occ, int lane, double tol_live, double mx, double theta_prev,
                                                  double remaining, double share) {
    const unsigned full = 0xffffffffu;
    double theta = 0.5 * fmax(mx, theta_prev);
    const unsigned long long words = (unsigned long long)__ballot_sync(full, occ[lane] != 0u) |
                                     ((unsigned long long)__ballot_sync(full, occ[32 + lane] != 0u) << 32);
    int bh = kBins - 2;
    {   // (1) ascending: longest prefix of bins whose lower-edge mass stays <= tol
        double low = 0.0; bool broke = false;
        unsigned long long ws = words;
        while (ws && !broke) {
            const int w = __ffsll((long long)ws) - 1; ws &= ws - 1;
            unsigned m = occ[w];
            while (m) {
                const int b = (w << 5) + __ffs((int)m) - 1; m &= m - 1;
                if (b > kBins - 2) break;
                low += (double)hist[b] * bin_edge(b);
                if (!(low <= tol_live)) { bh = b - 1; broke = true; break; }
            }
        }
    }
    int bb = 0, bf = 0;
    {   // (2) budget and (3) selectivity, descending: the lowest bin down to which bisecting everything still fits
        double above = 0.0; bool broke_b = false, broke_f = false;
        unsigned long long ws = words;
        while (ws && !(broke_b && broke_f)) {
            const int w = 63 - __clzll((long long)ws); ws &= ~(1ull << w);
            unsigned m = occ[w];
            while (m && !(broke_b && broke_f)) {
                const int b = (w << 5) + 31 - __clz((int)m); m &= ~(1u << (b & 31));
                if (b > kBins - 2) continue;
                above += (double)hist[b];
                if (!broke_b && 2.0 * above > remaining) { bb = b + 1; broke_b = true; }
                if (!broke_f && above > share) { bf = b + 1; broke_f = true; }
            }
        }
    }
    if (bf > bb) bb = bf;
    const int bs = (bh + 1 > bb) ? bh + 1 : bb;
    const double theta_h = bs <= 0 ? 0.0 : (bs >= kBins - 1 ? 1.7976931348623157e308 : bin_edge(bs));
    return fmin(theta, theta_h);
}

// ---- a sub-problem's arena is full (level 2 of ib200_hcubature_single): RETIRE the boxes that have converged (only their sums
//      are kept, in ret[0..2] = integral, error, count) and decide the round again, the selection capped by the room that is left.
//      Out of line on purpose: the hot loop of the round kernel keeps the registers it had before this existed (inlined, the batch kernels
//      went from 154 to 168 registers and BASELINE config 4 lost 7 %).  Oracle: the `arena` branch of hcub_rounds_core2 (oracle/oracle.c).
//      Retired error cannot be refined away, so a retirement may take at most a quarter of the tolerance still unspent, measured by the UPPER
//      edges of the histogram bins (a bound on the retired sum): every box below theta_r = that bin edge goes, never a box selected for
//      bisection (theta_r <= theta).  The histogram is rebuilt from the boxes that stay.  Second selection: bins are taken whole from the top
//      while they fit the room; of the bin that does not fit any more, the first `quota` boxes in store order.  false: the arena stays full
//      (less than 1/64 of it freed, or nothing to select).
template <int D>
__device__ __noinline__ bool retire_reselect(double *rec, double *err, int *list, unsigned *hist, unsigned *occ, double *ret, int cap,
                                             double tol, double mx, double theta_prev, double theta_cur, double remaining, double frac,
                                             int *nbox_io, int *nsplit_out, double *dE_out, double *theta_out) {
    constexpr int RECW = 2 * D + 2, NCH = D + 1;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int nbox = *nbox_io;
    __syncwarp();
    {
        const double R = 0.25 * (tol - ret[1]);
        const unsigned long long words = (unsigned long long)__ballot_sync(full, occ[lane] != 0u) |
                                         ((unsigned long long)__ballot_sync(full, occ[32 + lane] != 0u) << 32);
        int br = -1;
        {
            double acc = 0.0; bool broke = false;
            unsigned long long ws = words;
            while (ws && !broke) {
                const int w = __ffsll((long long)ws) - 1; ws &= ws - 1;
                unsigned m = occ[w];
                while (m) {
                    const int b = (w << 5) + __ffs((int)m) - 1; m &= m - 1;
                    if (b > kBins - 3) { broke = true; break; }
                    acc += (double)hist[b] * bin_edge(b + 1);
                    if (!(acc <= R)) { broke = true; break; }
                    br = b;
                }
            }
        }
        const double theta_r = fmin(br < 0 ? 0.0 : bin_edge(br + 1), theta_cur);
        __syncwarp();
        for (int b = lane; b < kBins; b += 32) hist[b] = 0u;
        for (int b = lane; b < kBins / 32; b += 32) occ[b] = 0u;
        __syncwarp();
        int kept = 0;
        double rI = 0.0, rE = 0.0;
        for (int i0 = 0; i0 < nbox; i0 += 32) {
            const int i = i0 + lane;
            const bool have = i < nbox;
            const double e = have ? __ldcg(err + i) : 0.0;
            const bool keep = have && e >= theta_r;
            double r[RECW];
            if (keep) {
                const double2 *src = reinterpret_cast<const double2 *>(rec + (size_t)i * RECW);
#pragma unroll
                for (int q = 0; q < NCH; q++) { const double2 v = __ldcg(src + q); r[2 * q] = v.x; r[2 * q + 1] = v.y; }
            } else if (have) { rI += __ldcg(rec + (size_t)i * RECW + 2 * D); rE += e; }
            const unsigned bal = __ballot_sync(full, keep);
            __syncwarp();                                       // every lane has read its box before a slot of this batch is overwritten
            if (keep) {
                const int j = kept + __popc(bal & lt);          // j <= i: a slot this pass has already read
                if (j != i) {
                    double2 *dst = reinterpret_cast<double2 *>(rec + (size_t)j * RECW);
#pragma unroll
                    for (int q = 0; q < NCH; q++) dst[q] = make_double2(r[2 * q], r[2 * q + 1]);
                    err[j] = e;
                }
                const int bn = ebin(e);
                if (atomicAdd(&hist[bn], 1u) == 0u) atomicOr(&occ[bn >> 5], 1u << (bn & 31));
            }
            kept += __popc(bal);
            __syncwarp();
        }
        rI = wsum(rI); rE = wsum(rE);
        if (lane == 0) { ret[0] += rI; ret[1] += rE; ret[2] += (double)(nbox - kept); }
        nbox = kept;
        __syncwarp();
    }
    *nbox_io = nbox;
    *nsplit_out = 0; *dE_out = 0.0;
    // ---- the round again, on what is left
    double theta = round_threshold(hist, occ, lane, tol - ret[1], mx, theta_prev, remaining, frac * (double)nbox);
    *theta_out = theta;
    const int room = cap - nbox;
    if (room < cap / 64) return false;                          // not worth a round: the arena stays full
    double th_hi = theta, th_lo = theta;
    int quota = 0;
    {
        const unsigned long long words = (unsigned long long)__ballot_sync(full, occ[lane] != 0u) |
                                         ((unsigned long long)__ballot_sync(full, occ[32 + lane] != 0u) << 32);
        double up = 0.0; bool crossed = false;
        unsigned long long ws = words;
        while (ws && !crossed) {
            const int w = 63 - __clzll((long long)ws); ws &= ~(1ull << w);
            unsigned m = occ[w];
            while (m && !crossed) {
                const int b = (w << 5) + 31 - __clz((int)m); m &= ~(1u << (b & 31));
                if (b > kBins - 2) continue;
                const double c = (double)hist[b];
                if (up + c > (double)room) {
                    th_hi = fmax(theta, b + 1 >= kBins - 1 ? 1.7976931348623157e308 : bin_edge(b + 1));
                    th_lo = fmax(theta, bin_edge(b));
                    quota = room - (int)up;
                    crossed = true;
                } else up += c;
            }
        }
    }
    int nsplit = 0, nlo = 0;
    double dE = 0.0;
    for (int i0 = 0; i0 < nbox; i0 += 32) {
        const int i = i0 + lane;
        const double e = i < nbox ? __ldcg(err + i) : -1.0;
        const bool hi = e >= th_hi;
        const bool lo = !hi && e >= th_lo;
        const unsigned bl = __ballot_sync(full, lo);
        const bool flag = hi || (lo && nlo + __popc(bl & lt) < quota);
        nlo += __popc(bl);
        const unsigned bal = __ballot_sync(full, flag);
        if (flag) {
            list[nsplit + __popc(bal & lt)] = i;
            dE += e;
            const int bn = ebin(e);
            if (atomicSub(&hist[bn], 1u) == 1u) atomicAnd(&occ[bn >> 5], ~(1u << (bn & 31)));
        }
        nsplit += __popc(bal);
    }
    *nsplit_out = nsplit; *dE_out = dE;
    return nsplit > 0 && nbox + nsplit <= cap;
}

// shared memory.  Per warp: hist[kBins] (unsigned) | occ[kBins/32] (bitmap of the occupied bins) | ret[4] | problem row | staging.
//   thread-per-box: row = par[NPAR] lbs[D] aux[D] finit jacc prep[D] (the layout gm_rule_thread / gm_rule_thread2 read); staging = [2][D+1][32] double2, the
//                   parent records of the current and the next pass (cp.async)
//   warp-per-box:   row = par[NPAR] lb[D] ub[D] den[D] (WarpRule::kProb); staging = WarpRule::kStage doubles + the parent record
// Per block (warp-per-box only): the rule's weights and point table.
template <int FAM, int D> __host__ __device__ constexpr int pc_padded() { return (hp::pc_doubles<FAM, D>() + D + 1) & ~1; }   // + prep[D] (gm_rule_thread2.cuh)
template <int FAM, int D, bool ALLFIN> __host__ __device__ constexpr int row_doubles() {
    if (!warp_box<FAM, D>()) return kBins / 2 + kBins / 64 + 4 + pc_padded<FAM, D>() + 2 * (D + 1) * 32 * 2;
    using W = hs::WarpRule<FAM, D, ALLFIN>;
    return kBins / 2 + kBins / 64 + 4 + ((W::kProb + 1) & ~1) + ((W::kStage + 1) & ~1) + 2 * D + 2;
}
template <int FAM, int D> __host__ __device__ constexpr int block_doubles() {
    return !warp_box<FAM, D>() ? 0 : 10 + (hc::gm_npoints(D) + 1) / 2 + 1;
}

#ifndef HB_MINB
#define HB_MINB 3
#endif
#ifndef HB_MINB_HD
#define HB_MINB_HD 3            // thread per box in 6 .. 8 dimensions
#endif
// SUB: sub-problem mode (level 2 of ib200_hcubature_single) is a separate instantiation -- its arena-retirement call (an ABI call in the
// round loop) costs the batch kernels 13 registers and 11 % on the exponential families when it is merely present (measured, same box)
template <int FAM, int D, bool ALLFIN, bool SUB, bool CPT = false>
__global__ void __launch_bounds__(kBlock, (D <= kMaxThreadDim ? HB_MINB : (warp_box<FAM, D>() ? 4 : HB_MINB_HD)))
hcub_rb_kernel(const Params p) {
    static_assert(!CPT || (ALLFIN && !SUB && D <= kMaxThreadDim), "compact records: batch mode, finite domains, thread per box");
    using F = ib::Family<FAM, D>;
    using W = hs::WarpRule<FAM, D, ALLFIN>;
    constexpr bool WB = warp_box<FAM, D>();               // warp per box
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    constexpr int NP = hc::gm_npoints(D);
    constexpr int RECW = 2 * D + 2, NCH = D + 1;          // a record = NCH 16-byte chunks
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + wib;
    double *s_w = sm;                                                          // WB: [10]
    unsigned *s_pt = reinterpret_cast<unsigned *>(sm + 10);                    // WB: [NP]
    double *row = sm + block_doubles<FAM, D>() + (size_t)wib * row_doubles<FAM, D, ALLFIN>();
    unsigned *hist = reinterpret_cast<unsigned *>(row);
    unsigned *occ = reinterpret_cast<unsigned *>(row + kBins / 2);
    double *ret = row + kBins / 2 + kBins / 64;           // [4] sums over the boxes retired from a full arena: integral, error, count
    double *pc = ret + 4;
    // thread per box
    double2 *stage = reinterpret_cast<double2 *>(pc + pc_padded<FAM, D>());     // [2][NCH][32]
    double *par = pc, *lbs = pc + NPAR, *aux = lbs + D;
    // warp per box
    double *wstage = pc + ((W::kProb + 1) & ~1);
    double *prec = wstage + ((W::kStage + 1) & ~1);                             // [RECW] the parent's record
    if (WB) {
        W::setup_tables(s_w, s_pt, threadIdx.x, kBlock);
        __syncthreads();
    }

    double *rec = p.rec + (size_t)gwarp * p.cap * RECW;
    unsigned char *recb = reinterpret_cast<unsigned char *>(p.rec) + (size_t)gwarp * p.cap * kCptBytes;     // CPT: this warp's compact records
    double *err = p.err + (size_t)gwarp * p.cap;
    int *list = p.list + (size_t)gwarp * p.cap;
    constexpr bool sub = SUB;

    for (;;) {
        unsigned long long pidx = 0;
        if (lane == 0) pidx = atomicAdd(p.counter, 1ULL);
        pidx = __shfl_sync(full, pidx, 0);
        if (p.norder ? pidx >= *p.norder : (int64_t)pidx >= p.nprob) break;
        const int64_t prob = sub ? 0 : (p.order ? (int64_t)p.order[pidx] : (int64_t)pidx);   // parameter column
        // sub mode: box of the level-1 store -- entry (pidx % chunk) of this rank's chunk number pidx / chunk
        const int64_t sbox = (((int64_t)pidx / hs::kSubChunk) * p.sub_world + p.sub_rank) * hs::kSubChunk + (int64_t)pidx % hs::kSubChunk;
        if (sub && sbox >= p.sub_n) {                                            // tail of the last chunk
            if (lane == 0) { double *o = p.sub_out + (size_t)pidx * 4; o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; }
            continue;
        }

        // ---- problem data -> this warp's shared row; empty histogram
        __syncwarp();
        for (int t = lane; t < NPAR; t += 32) par[t] = p.params[prob * p.nparam + t];
        for (int k = lane; k < D; k += 32) {
            const double l = p.bpp ? p.lb[prob * D + k] : p.lb[k];
            const double u = p.bpp ? p.ub[prob * D + k] : p.ub[k];
            if (WB) { lbs[k] = l; lbs[D + k] = u; lbs[2 * D + k] = (u - l) * 0.5; }
            else { lbs[k] = l; aux[k] = ALLFIN ? (u - l) * 0.5 : u; }
        }
        for (int b = lane; b < kBins; b += 32) hist[b] = 0u;
        for (int b = lane; b < kBins / 32; b += 32) occ[b] = 0u;
        if (lane < 4) ret[lane] = 0.0;
        __syncwarp();
        double jacc = 1.0;
        if (ALLFIN) {
#pragma unroll
            for (int k = 0; k < D; k++) { const double den = WB ? lbs[2 * D + k] : aux[k]; jacc = (k == 0) ? den : jacc * den; }
        }
        const double finit = F::sep_init(par);
        if (!WB && lane == 0) { aux[D] = finit; aux[D + 1] = jacc; g2::Prep<FAM, D>::run(par, aux + D + 2); }
        __syncwarp();

        double rtol = p.rtol, atol = p.atol;
        int qtotal = 1;
        bool degenerate = false;
        if (sub) {
            // ---- level 2: this box's share of the tolerance; already below it -> the box stands as it is
            const double e_i = p.sub_err[sbox];
            const double eps = p.sub_tol * (pow(e_i, p.sub_alpha) / p.sub_wsum);
            const double *src = p.sub_rec + (size_t)sbox * RECW;
            if (!(e_i > eps)) {
                if (lane == 0) { double *o = p.sub_out + (size_t)pidx * 4; o[0] = src[2 * D]; o[1] = e_i; o[2] = 0.0; o[3] = 0.0; }
                continue;
            }
            rtol = 0.0; atol = eps;
            // the box is evaluated already: its first round is its own bisection -> the store starts with the two halves
            if (lane < 2) {
                double r[RECW];
#pragma unroll
                for (int k = 0; k < RECW; k++) r[k] = src[k];
                const int kd = (int)r[2 * D + 1];
#pragma unroll
                for (int k = 0; k < D; k++) if (k == kd) { const double w = (r[D + k] - r[k]) * 0.5; if (lane == 0) r[k] = r[k] + w; else r[D + k] = r[D + k] - w; }
                r[2 * D] = 0.0; r[2 * D + 1] = 0.0;
                double2 *dst = reinterpret_cast<double2 *>(rec + (size_t)lane * RECW);
#pragma unroll
                for (int q = 0; q < NCH; q++) dst[q] = make_double2(r[2 * q], r[2 * q + 1]);
            }
            qtotal = 2;
        } else if constexpr (CPT) {
            if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps), HCubature default
            if (lane == 0) {                                                // the root box [-1, 1]^D: all indices and levels zero
                uint4 *dst = reinterpret_cast<uint4 *>(recb);
                dst[0] = make_uint4(0u, 0u, 0u, 0u); dst[1] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps), HCubature default
            // ---- initial grid (HCubature initdiv; CartesianIndices order, first index fastest) -> slots 0 .. qtotal-1, unevaluated:
            //      the evaluation loop below takes them like the children of a split (the rule is instantiated once)
            double prodD = 1.0;
            double tl[D], th[D];
#pragma unroll
            for (int k = 0; k < D; k++) {
                tl[k] = -1.0; th[k] = 1.0;
                if (!ALLFIN) { ib::AxisMap m; ib::make_axis(lbs[k], WB ? lbs[D + k] : aux[k], m); tl[k] = m.tlo; th[k] = m.thi; }
                qtotal *= p.initdiv;
                prodD *= (th[k] - tl[k]) / p.initdiv;
            }
            degenerate = (prodD == 0.0);
            if (degenerate) qtotal = 1;                 // HCubature returns after the first box
            for (int t = lane; t < qtotal; t += 32) {
                double r[RECW];
                int q = t;
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const int cidx = q % p.initdiv; q /= p.initdiv;
                    const double Dl = (th[k] - tl[k]) / p.initdiv;
                    r[k] = tl[k] + cidx * Dl;
                    r[D + k] = (cidx == p.initdiv - 1) ? th[k] : r[k] + Dl;
                }
                r[2 * D] = 0.0; r[2 * D + 1] = 0.0;
                double2 *dst = reinterpret_cast<double2 *>(rec + (size_t)t * RECW);
#pragma unroll
                for (int q2 = 0; q2 < NCH; q2++) dst[q2] = make_double2(r[2 * q2], r[2 * q2 + 1]);
            }
        }
        __syncwarp();

        int nbox = qtotal, ntodo = qtotal, nsplit = 0, nbox0 = 0, status = IB200_ST_CONVERGED;
        long long applied = 0;
        double I = 0.0, E = 0.0, theta = 0.0, dE = 0.0;
        bool first = true;
        bool ovf = false;                                  // CPT: a box of this lane cannot be halved again in compact form

        for (;;) {
            // ---- evaluate.  Round 0: the boxes of the initial store (entry t = slot t).  Later rounds: both halves of every selected box.
            double sI = 0.0, sE = -dE, mx = 0.0;       // this lane's share of the round's new boxes (minus the parents' errors)
            if constexpr (!WB) {
                // thread per box: lanes 2r, 2r+1 take the children of list[r].  The records a pass needs are staged into shared memory
                // with cp.async while the previous pass is evaluated, and the list entries of the pass after that are in flight behind
                // them, so a pass does not wait for HBM (ncu on the first version: long_scoreboard was the top stall).
                const int npass = (ntodo + 31) >> 5;
                auto entry = [&](int t) { return first ? t : __ldcg(list + (t >> 1)); };
                auto issue = [&](int src_slot, int buf) {        // record -> stage[buf][chunk][lane]
                    if constexpr (CPT) {
                        const double2 *src = reinterpret_cast<const double2 *>(recb + (size_t)src_slot * kCptBytes);
                        hp::cp_async16(stage + ((size_t)buf * NCH + 0) * 32 + lane, src);
                        hp::cp_async16(stage + ((size_t)buf * NCH + 1) * 32 + lane, src + 1);
                    } else {
                        const double2 *src = reinterpret_cast<const double2 *>(rec + (size_t)src_slot * RECW);
#pragma unroll
                        for (int q = 0; q < NCH; q++) hp::cp_async16(stage + ((size_t)buf * NCH + q) * 32 + lane, src + q);
                    }
                };
                int par_cur = lane < ntodo ? entry(lane) : 0;
                if (lane < ntodo) issue(par_cur, 0);
                hp::cp_async_commit();
                int par_next = 32 + lane < ntodo ? entry(32 + lane) : 0;
                for (int ps = 0; ps < npass; ps++) {
                    const int t = (ps << 5) + lane;
                    const bool have = t < ntodo;
                    const int par_next2 = 64 + t < ntodo ? entry(64 + t) : 0;
                    hp::cp_async_wait_all();
                    double a[D], b[D], Ipar = 0.0;
                    int kd = -1;
                    unsigned cidx[CPT ? D : 1], cmeta = 0u;          // CPT: the child's indices and levels
                    if (have) {
                        if constexpr (CPT) {
                            const uint4 c0 = *reinterpret_cast<const uint4 *>(stage + ((size_t)(ps & 1) * NCH + 0) * 32 + lane);
                            const uint4 c1 = *reinterpret_cast<const uint4 *>(stage + ((size_t)(ps & 1) * NCH + 1) * 32 + lane);
                            const unsigned pidx4[5] = {c0.x, c0.y, c0.z, c0.w, c1.y};
                            Ipar = __hiloint2double((int)c1.w, (int)c1.z);
                            cmeta = c1.x & 0x0fffffffu;
                            if (!first) kd = (int)(c1.x >> 28);
                            const unsigned up = (t & 1) == 0 ? 1u : 0u;             // child 0 = upper half
#pragma unroll
                            for (int k = 0; k < D; k++) {
                                const unsigned lev = (cmeta >> (5 * k)) & 31u;
                                const bool mine = (k == kd);
                                if (mine && lev >= (unsigned)p.cpt_max_level) ovf = true;
                                const unsigned nlev = mine ? lev + 1u : lev;
                                cidx[k] = mine ? 2u * pidx4[k] + up : pidx4[k];
                                if (mine) cmeta += 1u << (5 * k);
                                // a = -1 + i 2^(1-l), b = a + 2^(1-l): exact (see the note at kCptBytes)
                                const double w = __hiloint2double((int)((1024u - nlev) << 20), 0);
                                a[k] = fma((double)cidx[k], w, -1.0);
                                b[k] = a[k] + w;
                            }
                        } else {
                            double r[RECW];
#pragma unroll
                            for (int q = 0; q < NCH; q++) { const double2 v = stage[((size_t)(ps & 1) * NCH + q) * 32 + lane]; r[2 * q] = v.x; r[2 * q + 1] = v.y; }
#pragma unroll
                            for (int k = 0; k < D; k++) { a[k] = r[k]; b[k] = r[D + k]; }
                            Ipar = r[2 * D];
                            if (!first) kd = (int)r[2 * D + 1];
                        }
                    }
                    if (32 + t < ntodo) issue(par_next, (ps + 1) & 1);
                    hp::cp_async_commit();
                    __syncwarp();            // every lane holds its parent's record before child 0 overwrites that slot
                    if (have) {
                        // child 0 = upper half (a[kd] += w) keeps the parent's slot, child 1 = lower half (b[kd] -= w) is appended
                        if constexpr (!CPT) {
#pragma unroll
                            for (int k = 0; k < D; k++) if (k == kd) { const double w = (b[k] - a[k]) * 0.5; if ((t & 1) == 0) a[k] = a[k] + w; else b[k] = b[k] - w; }
                        }
                        if (!first && (t & 1) == 0) sI -= Ipar;                   // the parent's integral leaves the sum
                        const int slot = (first || (t & 1) == 0) ? par_cur : nbox0 + (t >> 1);
                        double Ic, Ec; int kc = 0;
                        if constexpr (D > kMaxThreadDim) g3::gm_rule_thread3<FAM, D, ALLFIN>(pc, p.tr, a, b, stage + (size_t)(ps & 1) * NCH * 32 + lane, Ic, Ec, kc);
                        else if constexpr (FAM == IB200_FAM_GENZ_OSC && ALLFIN) hp::gm_rule_thread_osc<D>(pc, a, b, Ic, Ec, kc);
                        else if constexpr (ALLFIN) g2::gm_rule_thread2<FAM, D>(pc, a, b, Ic, Ec, kc);
                        else hp::gm_rule_thread<FAM, D, ALLFIN>(pc, p.tr, a, b, Ic, Ec, kc);
                        if constexpr (CPT) {
                            uint4 *dst = reinterpret_cast<uint4 *>(recb + (size_t)slot * kCptBytes);
                            dst[0] = make_uint4(cidx[0], cidx[1], D > 2 ? cidx[D > 2 ? 2 : 0] : 0u, D > 3 ? cidx[D > 3 ? 3 : 0] : 0u);
                            dst[1] = make_uint4(cmeta | ((unsigned)kc << 28), D > 4 ? cidx[D > 4 ? 4 : 0] : 0u,
                                                (unsigned)__double2loint(Ic), (unsigned)__double2hiint(Ic));
                        } else {
                            double2 *dst = reinterpret_cast<double2 *>(rec + (size_t)slot * RECW);
                            double r[RECW];
#pragma unroll
                            for (int k = 0; k < D; k++) { r[k] = a[k]; r[D + k] = b[k]; }
                            r[2 * D] = Ic; r[2 * D + 1] = (double)kc;
#pragma unroll
                            for (int q = 0; q < NCH; q++) dst[q] = make_double2(r[2 * q], r[2 * q + 1]);
                        }
                        err[slot] = Ec;
                        const int bn = ebin(Ec);
                        if (atomicAdd(&hist[bn], 1u) == 0u) atomicOr(&occ[bn >> 5], 1u << (bn & 31));
                        sI += Ic; sE += Ec; mx = fmax(mx, Ec);       // fmax drops NaN: a non-finite E is caught through sE
                    }
                    par_cur = par_next; par_next = par_next2;
                }
            } else {
                // warp per box: one entry (round 0: a box; later: a parent = two boxes) at a time; the next entry's record is
                // in flight in a register per lane while the rule runs
                const int nit = first ? ntodo : nsplit;
                double *box = W::box(wstage);
                auto entry = [&](int r0) { return first ? r0 : __ldcg(list + r0); };
                auto note = [&](double Ec) {                                      // lane 0
                    const int bn = ebin(Ec);
                    if (atomicAdd(&hist[bn], 1u) == 0u) atomicOr(&occ[bn >> 5], 1u << (bn & 31));
                    sE += Ec; mx = fmax(mx, Ec);
                };
                int src_cur = nit > 0 ? entry(0) : 0;
                double pre = (nit > 0 && lane < RECW) ? __ldcg(rec + (size_t)src_cur * RECW + lane) : 0.0;
                int src_next = nit > 1 ? entry(1) : 0;
                for (int r0 = 0; r0 < nit; r0++) {
                    __syncwarp();
                    if (lane < RECW) prec[lane] = pre;
                    const int src_next2 = r0 + 2 < nit ? entry(r0 + 2) : 0;
                    if (r0 + 1 < nit && lane < RECW) pre = __ldcg(rec + (size_t)src_next * RECW + lane);
                    __syncwarp();
                    if (first) {
                        if (lane < 2 * D) box[lane] = prec[lane];
                        __syncwarp();
                        double Iv, Er; int kc;
                        W::eval(pc, p.tr, finit, jacc, s_w, s_pt, wstage, lane, Iv, Er, kc);
                        if (lane == 0) {
                            rec[(size_t)src_cur * RECW + 2 * D] = Iv; rec[(size_t)src_cur * RECW + 2 * D + 1] = (double)kc;
                            err[src_cur] = Er;
                            sI += Iv; note(Er);
                        }
                    } else {
                        const int kd = (int)prec[2 * D + 1];
                        const double akd = prec[kd], bkd = prec[D + kd], w = (bkd - akd) * 0.5, Ipar = prec[2 * D];
                        // child 0 = upper half (a[kd] += w) keeps the parent's slot
                        if (lane < 2 * D) box[lane] = lane == kd ? akd + w : prec[lane];
                        __syncwarp();
                        double I0, E0; int k0;
                        W::eval(pc, p.tr, finit, jacc, s_w, s_pt, wstage, lane, I0, E0, k0);
                        __syncwarp();
                        // child 1 = lower half (b[kd] -= w) is appended
                        if (lane < 2 * D) box[lane] = lane == D + kd ? bkd - w : prec[lane];
                        __syncwarp();
                        double I1, E1; int k1;
                        W::eval(pc, p.tr, finit, jacc, s_w, s_pt, wstage, lane, I1, E1, k1);
                        const size_t s1 = (size_t)(nbox0 + r0);
                        if (lane < 2 * D) rec[s1 * RECW + lane] = box[lane];
                        if (lane == 0) {
                            double *r = rec + (size_t)src_cur * RECW;
                            r[kd] = akd + w; r[2 * D] = I0; r[2 * D + 1] = (double)k0; err[src_cur] = E0;
                            rec[s1 * RECW + 2 * D] = I1; rec[s1 * RECW + 2 * D + 1] = (double)k1; err[s1] = E1;
                            sI += (I0 + I1) - Ipar; note(E0); note(E1);
                        }
                    }
                    src_cur = src_next; src_next = src_next2;
                }
            }
            first = false;

            // ---- close the round: sums over the new boxes
            sI = wsum(sI); sE = wsum(sE); mx = wmax(mx);
            I += sI; E += sE;
            applied += ntodo;
            __syncwarp();
            if constexpr (CPT) { if (__any_sync(full, ovf)) { status = -1; break; } }      // handed back to the full-record kernel

            // ---- HCubature's stopping test
            const double tol = fmax(rtol * fabs(I), atol);
            if (degenerate || E <= tol) break;
            if (!isfinite(E)) { status = IB200_ST_NONFINITE; break; }
            if (applied >= p.max_apps) { status = IB200_ST_MAXEVALS; break; }

            // ---- threshold (round_threshold above) and selection.  ret[1] = the error retired from this sub-problem's arena so far (0 in batch mode)
            const double theta_prev = theta;
            theta = round_threshold(hist, occ, lane, tol - ret[1], mx, theta_prev, (double)p.max_apps - (double)applied, p.frac * (double)nbox);

            // ---- select: slots with error >= theta, in store order.  The pass is a latency-bound stream of the problem's keys (ncu: a fifth
            //      of the stall samples of the product-peak kernel sat on its first compare), so the keys of the NEXT 512 boxes are already
            //      in flight while the current ones are compacted (HB_KU = 8 independent loads per lane, double buffered; measured against 16 and against
            //      no double buffering: 8 wins on the families with smaller stores, 16 only on the product peak)
            nsplit = 0;
            dE = 0.0;
            {
#ifndef HB_KU
#define HB_KU 8
#endif
                constexpr int KU = HB_KU;
                double en[KU];
#pragma unroll
                for (int u = 0; u < KU; u++) { const int i = 32 * u + lane; en[u] = i < nbox ? __ldcg(err + i) : -1.0; }
                for (int i0 = 0; i0 < nbox; i0 += 32 * KU) {
                    double e[KU];
#pragma unroll
                    for (int u = 0; u < KU; u++) e[u] = en[u];
#ifdef HB_NODB
                    if (i0 > 0) {
#pragma unroll
                        for (int u = 0; u < KU; u++) { const int i = i0 + 32 * u + lane; e[u] = i < nbox ? __ldcg(err + i) : -1.0; }
                    }
                    if (false) {
#else
                    if (i0 + 32 * KU < nbox) {
#endif
#pragma unroll
                        for (int u = 0; u < KU; u++) { const int i = i0 + 32 * KU + 32 * u + lane; en[u] = i < nbox ? __ldcg(err + i) : -1.0; }
                    }
#pragma unroll
                    for (int u = 0; u < KU; u++) {
                        if (i0 + 32 * u >= nbox) break;                // (uniform) short stores: no work for the empty part of the batch
                        const int i = i0 + 32 * u + lane;
                        const bool flag = e[u] >= theta;
                        const unsigned bal = __ballot_sync(full, flag);
                        if (flag) {
                            list[nsplit + __popc(bal & lt)] = i;
                            dE += e[u];
                            const int bn = ebin(e[u]);
                            if (atomicSub(&hist[bn], 1u) == 1u) atomicAnd(&occ[bn >> 5], ~(1u << (bn & 31)));
                        }
                        nsplit += __popc(bal);
                    }
                }
            }
            if (nbox + nsplit > p.cap) {
                // arena full.  Batch mode: leave the boxes alone.  Sub-problem mode: retire the converged boxes and decide the round again
                bool ok = false;
                if constexpr (SUB) {
                    int nb = nbox, ns = 0; double de = 0.0, th = theta;
                    ok = retire_reselect<D>(rec, err, list, hist, occ, ret, p.cap, tol, mx, theta_prev, theta, (double)p.max_apps - (double)applied,
                                            p.frac, &nb, &ns, &de, &th);
                    nbox = nb; nsplit = ns; dE = de; theta = th;
                }
                if (!ok) { status = IB200_ST_MAXEVALS; break; }
            }
            nbox0 = nbox;
            nbox += nsplit; ntodo = 2 * nsplit;
            __syncwarp();                 // the split list is read across lanes
        }

        if constexpr (CPT) {
            if (status < 0) {
                if (lane == 0) { const unsigned long long i = atomicAdd(p.ovf_count, 1ULL); p.ovf_list[i] = (unsigned)prob; }
                continue;
            }
        }
        // ---- HCubature's closing sum over the store
        double pI = 0.0, pE = 0.0;
        if constexpr (CPT) {
            const double *ri = reinterpret_cast<const double *>(recb + 24);      // I of slot s: ri[4 s]
            int s = lane;
            for (; s + 7 * 32 < nbox; s += 8 * 32) {
                double vi[8], ve[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { vi[u] = __ldcg(ri + (size_t)(s + 32 * u) * 4); ve[u] = __ldcg(err + s + 32 * u); }
#pragma unroll
                for (int u = 0; u < 8; u++) { pI += vi[u]; pE += ve[u]; }
            }
            for (; s < nbox; s += 32) { pI += __ldcg(ri + (size_t)s * 4); pE += __ldcg(err + s); }
        } else {
            int s = lane;
            for (; s + 7 * 32 < nbox; s += 8 * 32) {             // 16 loads in flight per lane, added in slot order
                double vi[8], ve[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { vi[u] = __ldcg(rec + (size_t)(s + 32 * u) * RECW + 2 * D); ve[u] = __ldcg(err + s + 32 * u); }
#pragma unroll
                for (int u = 0; u < 8; u++) { pI += vi[u]; pE += ve[u]; }
            }
            for (; s < nbox; s += 32) { pI += __ldcg(rec + (size_t)s * RECW + 2 * D); pE += __ldcg(err + s); }
        }
        pI = wsum(pI) + ret[0]; pE = wsum(pE) + ret[1];          // + what was retired from a full arena (sub-problem mode)
        if (lane == 0) {
            if (sub) {
                if (ret[2] > 0.0) { atomicAdd(p.counter + 4, 1ULL); atomicAdd(p.counter + 5, (unsigned long long)ret[2]); }
                double *o = p.sub_out + (size_t)pidx * 4;
                o[0] = pI; o[1] = pE; o[2] = (double)applied; o[3] = (double)status;
                // diagnostics behind the counter: sub-problems ended by their budget / by a full arena / non-finite
                if (status == IB200_ST_MAXEVALS) atomicAdd(p.counter + (applied >= p.max_apps ? 1 : 2), 1ULL);
                else if (status == IB200_ST_NONFINITE) atomicAdd(p.counter + 3, 1ULL);
            } else {
                p.out_I[prob] = pI;
                if (p.out_E) p.out_E[prob] = pE;
                if (p.out_nevals) p.out_nevals[prob] = (int64_t)applied * NP;
                if (p.out_status) p.out_status[prob] = status;
            }
        }
    }
}

inline size_t arena_bytes_per_box(int D) { return (size_t)8 * (2 * D + 2) + 8 + 4; }

template <int D>
int32_t launch(ib200_ctx *ctx, int family, bool allfin, Params &p, long long arena_mib = 0);

// Sizes the persistent grid from the kernel's occupancy, carves the per-warp arenas out of the context's scratch and launches.
// arena_mib > 0 (sub-problem mode): device memory the arenas may take -- the GRID shrinks to fit, never the arena of a warp, so a
// sub-problem meets the same capacity (and gives the same result) whatever the free memory or the number of ranks.
template <int D>
int32_t launch_impl(ib200_ctx *ctx, int family, bool allfin, Params &p, long long arena_mib) {
    int32_t rc = IB200_ERR_UNSUPPORTED;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, D)) {
            auto go2 = [&](auto kern, size_t smem, Params &p) -> int32_t {
                if (smem > ctx->smem_optin) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_hcubature_batch: shared memory of the round-based kernel exceeds the device limit");
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
                if (blocks < 1) blocks = 1;
                if (arena_mib > 0) {
                    const long long fit = (long long)(((size_t)arena_mib << 20) / ((size_t)kWarps * (size_t)p.cap * arena_bytes_per_box(D)));
                    if (fit < 1) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_single: no device memory left for the level-2 arenas (lower the single_sub_boxes option)");
                    if (blocks > fit) blocks = fit;
                }
                const size_t warps = (size_t)blocks * kWarps;
                const size_t cap = (size_t)p.cap;
                const bool sub = p.sub_rec != nullptr;                                       // level 2 keeps its arenas apart from the level-1 store
                p.rec = (double *)ctx->get(sub ? SL_ARENA0 : SL_WORK2, warps * cap * (2 * D + 2) * sizeof(double));
                p.err = (double *)ctx->get(sub ? SL_ARENA1 : SL_WORK4, warps * cap * sizeof(double));
                p.list = (int *)ctx->get(sub ? SL_ARENA2 : SL_WORK3, warps * cap * sizeof(int));
                if (!p.rec || !p.err || !p.list)
                    return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: box arena allocation failed (lower maxiters or the hcub_capacity option)");
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(p);
                IB_LAUNCH_CHECK(ctx);
                return IB200_OK;
            };
            auto go = [&](auto kern, size_t smem) -> int32_t { return go2(kern, smem, p); };
            const size_t smem_fin = (size_t)(block_doubles<FAM, D>() + kWarps * row_doubles<FAM, D, true>()) * sizeof(double);
            const size_t smem_inf = (size_t)(block_doubles<FAM, D>() + kWarps * row_doubles<FAM, D, false>()) * sizeof(double);
            if (p.sub_rec != nullptr) rc = allfin ? go(hcub_rb_kernel<FAM, D, true, true>, smem_fin) : go(hcub_rb_kernel<FAM, D, false, true>, smem_inf);
            else if (!allfin) rc = go(hcub_rb_kernel<FAM, D, false, false>, smem_inf);
            else {
                bool compact = false;
                if constexpr (D <= kMaxThreadDim && HB_COMPACT) compact = p.initdiv == 1 && p.nprob < ((int64_t)1 << 32) && ctx->opt_hcub_compact != 0;
                if (!compact) rc = go(hcub_rb_kernel<FAM, D, true, false>, smem_fin);
                else if constexpr (D <= kMaxThreadDim && HB_COMPACT) rc = [&]() -> int32_t {
                    // compact records first (see kCptBytes); the problems it hands back -- a box halved 32 times along one axis -- are solved by
                    // the full-record kernel right behind it, from the device-side list, in the same arenas (sized for the full records)
                    unsigned long long *aux = (unsigned long long *)ctx->get(SL_WORK5, 2 * sizeof(unsigned long long) + (size_t)p.nprob * sizeof(unsigned));
                    if (!aux) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: scratch allocation failed");
                    IB_CUDA(ctx, cudaMemsetAsync(aux, 0, 2 * sizeof(unsigned long long), ctx->stream));
                    p.ovf_count = aux; p.ovf_list = reinterpret_cast<unsigned *>(aux + 2);
                    p.cpt_max_level = (int)ctx->opt_hcub_compact_level;
                    const int32_t rc1 = go(hcub_rb_kernel<FAM, D, true, false, true>, smem_fin);
                    if (rc1 != IB200_OK) return rc1;
                    Params q = p;
                    q.order = p.ovf_list; q.norder = p.ovf_count; q.counter = aux + 1; q.ovf_list = nullptr; q.ovf_count = nullptr;
                    return go2(hcub_rb_kernel<FAM, D, true, false>, smem_fin, q);
                }();
            }
        }
    });
    if (rc == IB200_ERR_UNSUPPORTED && ctx->err.empty())
        return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_batch: family not compiled in for this dimension (round-based mode)");
    return rc;
}

}  // namespace hb
