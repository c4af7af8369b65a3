// hcub_rounds.cuh -- the round-synchronous refinement of ONE integral on a device-resident box store: state block, the
// rule evaluation kernel for the registered families and the histogram / decide / count / scan / scatter bookkeeping
// kernels.  Shared by hcub_single.cu (registered integrands, 1-8 GPUs; the scheme is described there) and
// batch_bridge.cu (caller-evaluated integrands: the same rounds with the evaluation step handed to the caller).
#pragma once
#include "hcubature_impl.cuh"

namespace hs {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int kRoundsPerCheck = 4;
constexpr int kBins = 2048;                   // one bin per binary exponent of a double
constexpr int kXchg = 8;                      // doubles exchanged per rank per round
constexpr int kMaxWorld = 8;
constexpr int kFin = 1024;                    // threads of the single-block bookkeeping kernels

struct State {
    long long nbox;          // boxes in the local store
    long long ntodo;         // boxes waiting for evaluation (indices in todo[])
    long long nsplit;        // boxes split in the current round
    long long napplied;      // local rule applications so far
    double I, E;             // local running sums over the store
    double I_g, E_g;         // global sums after the last exchange
    double gapplied;         // global rule applications
    double bound;            // upper bound of the largest box error anywhere
    double theta;            // split threshold of the current round
    int done, status, full, round;
};

struct Params {
    const double *par; const double *lb, *ub; int tr;
    double rtol, atol; long long max_boxes; long long cap;
    double frac;             // largest share of the stored boxes bisected in one round
    double *rec;             // [cap][2D+2]  a[D], b[D], I, split axis
    double *err;             // [cap]
    unsigned *todo;          // [cap]
    double *partial;         // [grid][4]    per-block sums (fixed-order finish)
    long long *bcount;       // [grid + 1]   per-block split counts / offsets
    unsigned long long *hist;  // [kBins]    local exponent histogram of err[0..nbox)
    double *sendbuf, *recvbuf;
    State *st;
    int rank, world;
    // result exchange (world > 1): this rank's results of a round are listed compactly in `outbox` ({slot, I, E, axis}: 32 B),
    // copied with coalesced 16-byte stores into every peer's `inbox` (CUDA IPC mappings; two round-parity halves, one
    // segment per sender) and applied to the peer's store after the round's all-gather
    double *outbox;                // [obcap][4]
    double *inbox;                 // [2][world][obcap][4]   (this rank's, exported)
    double *inbox_peer[kMaxWorld]; // the peers' inboxes
    long long obcap;
    int grid;
    // two-level solve (hcub_single.cu): hand the store over to the batched sub-problem driver once it holds >= switch_boxes boxes
    // and the error is within switch_factor of the tolerance (or it holds 64 x switch_boxes boxes); 0 = single level
    long long switch_boxes = 0;
    double switch_factor = 0.0;
    double switch_cap = 0.0;      // ... or once it holds this many boxes: min(64 x switch_boxes, half the store's capacity) -- a round can grow
                                  // the store by half, so the hand-over always comes before the store is full
};
constexpr int kStatusHandover = 4;            // internal: level 1 ended by the hand-over rule (never returned to the caller)

// ---- kernels ---------------------------------------------------------------------------------------
// the initial box: the whole transformed domain (identical on every rank)
template <int D, bool ALLFIN>
__global__ void init_kernel(Params p) {
    State *st = p.st;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double *r = p.rec;
#pragma unroll
        for (int k = 0; k < D; k++) {
            double a = -1.0, b = 1.0;
            if (!ALLFIN) { ib::AxisMap m; ib::make_axis(p.lb[k], p.ub[k], m); a = m.tlo; b = m.thi; }
            r[k] = a; r[D + k] = b;
        }
        r[2 * D] = 0.0; r[2 * D + 1] = 0.0;
        p.err[0] = 0.0;
        p.todo[0] = 0u;
        st->nbox = 1; st->ntodo = 1; st->nsplit = 0; st->napplied = 0;
        st->I = 0.0; st->E = 0.0; st->I_g = 0.0; st->E_g = 0.0; st->gapplied = 0.0;
        st->bound = 0.0; st->theta = 0.0; st->done = 0; st->status = 0; st->full = 0; st->round = 0;
    }
}

// block-level fixed-order reduction of (v0, v1, v2=max): result in thread 0
__device__ __forceinline__ void block_reduce3(double &v0, double &v1, double &v2, double *sh /* 3*kWarps */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v0 += __shfl_down_sync(0xffffffffu, v0, off);
        v1 += __shfl_down_sync(0xffffffffu, v1, off);
        v2 = fmax(v2, __shfl_down_sync(0xffffffffu, v2, off));
    }
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sh[w] = v0; sh[kWarps + w] = v1; sh[2 * kWarps + w] = v2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        v0 = sh[0]; v1 = sh[kWarps]; v2 = sh[2 * kWarps];
        for (int k = 1; k < kWarps; k++) { v0 += sh[k]; v1 += sh[kWarps + k]; v2 = fmax(v2, sh[2 * kWarps + k]); }
    }
}

// ---- the Genz-Malik rule on ONE box by ONE warp, in two phases: per-axis term tables (the 7 abscissae of every axis through the
//      variable transform and the family's separable term), then the points over the lanes.  Shared by the single-domain evaluation
//      kernel below and by the warp-per-box variant of the round-based batch driver (hcubature_rb.cuh).
template <int FAM, int D, bool ALLFIN>
struct WarpRule {
    using F = ib::Family<FAM, D>;
    static constexpr int NP = hc::gm_npoints(D);
    static constexpr int NAX = 1 + 4 * D;
    static constexpr int NPAR = ib::fam_nparam(FAM, D);
    static constexpr bool FUSE = F::kProd && !ALLFIN;      // product families: one table of T_k * J_k (finish is the identity)
    static constexpr bool SEPJ = !ALLFIN && !FUSE;         // sum families on infinite axes: separate Jacobian table
    static constexpr int DL = D / 2, DH = D - DL, NL = 1 << DL, NH = 1 << DH, NCORN = 1 << D;
    static constexpr int kStage = 14 * D + 14 * D + NAX + D + 2 * D + 2 * (NL + NH);       // doubles of per-warp staging
    static constexpr int kProb = NPAR + 3 * D;                                              // par | lb | ub | den

    // rule weights s_w[10] and the point table s_pt[NP] (3 bits per axis = abscissa id, class in bits 24-26): `nthreads` threads cooperate
    __device__ static void setup_tables(double *s_w, unsigned *s_pt, int tid, int nthreads) {
        if (tid == 0) {
            const double twon = (double)(1 << D);
            s_w[0] = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0); s_w[1] = twon * (980 / 6561.0);
            s_w[2] = twon * ((1820 - 400 * D) / 19683.0); s_w[3] = twon * (200 / 19683.0); s_w[4] = 6859 / 19683.0;
            s_w[5] = twon * ((729 - 950 * D + 50 * D * D) / 729.0); s_w[6] = twon * (245 / 486.0);
            s_w[7] = twon * ((265 - 100 * D) / 1458.0); s_w[8] = twon * (25 / 729.0); s_w[9] = 0.0;
        }
        for (int q = tid; q < NP; q += nthreads) {
            unsigned e = 0, cls = 0;
            if (q == 0) cls = 0;
            else if (q < 1 + 2 * D) { const int r = q - 1; cls = 1; e = (1u + (r & 1)) << (3 * (r >> 1)); }
            else if (q < 1 + 4 * D) { const int r = q - 1 - 2 * D; cls = 2; e = (3u + (r & 1)) << (3 * (r >> 1)); }
            else if (q < 1 + 4 * D + 2 * D * (D - 1)) {
                int r = q - 1 - 4 * D; const int s = r & 3; r >>= 2;
                int i = 0, j = 1;
                for (int cnt = 0; cnt < r; cnt++) { if (++j == D) { ++i; j = i + 1; } }
                cls = 3; e = ((3u + ((s >> 1) & 1)) << (3 * i)) | ((3u + (s & 1)) << (3 * j));
            } else {
                const unsigned r = q - (1 + 4 * D + 2 * D * (D - 1)); cls = 4;
                for (int k = 0; k < D; k++) e |= (5u + ((r >> k) & 1u)) << (3 * k);
            }
            s_pt[q] = e | (cls << 24);
        }
    }

    // box(stage) = the 2 D doubles a[D], b[D] of the box to evaluate, written by the caller (followed by __syncwarp)
    __device__ static double *box(double *stage) { return stage + 14 * D + 14 * D + NAX + D; }

    // prob = par[NPAR] | lb[D] | ub[D] | den[D] (shared or global); finit = F::sep_init(par); jacc = prod den (ALLFIN).
    // Result (integral, error, split axis) valid in LANE 0.
    __device__ static void eval(const double *prob, int tr, double finit, double jacc, const double *s_w, const unsigned *s_pt,
                                double *stage, int lane, double &Iv, double &Er, int &kdiv) {
        const double *par = prob, *lbs = prob + NPAR, *ubs = lbs + D, *dens = ubs + D;
        const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
        double *tabT = stage, *tabJ = tabT + 14 * D, *fst = tabJ + 14 * D, *hh = fst + NAX, *bx = hh + D;
        double *cornA = bx + 2 * D, *cornJ = cornA + (NL + NH);     // half-products of the corner abscissae: [NL low | NH high]
        for (int idx = lane; idx < 7 * D; idx += 32) {                  // phase 1: tables
            const int k = idx / 7, o = idx - 7 * k;
            const double a = bx[k], b = bx[D + k];
            const double c = 0.5 * (a + b), h = 0.5 * fabs(b - a);
            const double lam = o == 0 ? 0.0 : (o <= 2 ? l2 : (o <= 4 ? l3 : l5));
            const double dl = h * lam;
            const double tt = o == 0 ? c : ((o & 1) ? c + dl : c - dl);
            double u, jac;
            if (ALLFIN) { u = lbs[k] + (1.0 + tt) * dens[k]; jac = 1.0; }
            else { ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m); ib::t2ujac(tr, tt, m, u, jac); }
            const double T = F::sep_term(k, u, par);
            tabT[idx] = FUSE ? T * jac : T;
            if (SEPJ) tabJ[idx] = jac;
            if (o == 0) hh[k] = h;
        }
        __syncwarp();
        // phase 1b: the 2^D corner points (+-lambda5 on every axis) are products of one entry per axis; tabulate the
        // 2^(D/2) combinations of the low and of the high axes once, a corner is then ONE combination of each
        for (int idx = lane; idx < NL + NH; idx += 32) {
            const bool high = idx >= NL;
            const int c = high ? idx - NL : idx, k0 = high ? DL : 0, nk = high ? DH : DL;
            double acc = F::kProd ? 1.0 : 0.0, jac = 1.0;
            for (int t = nk - 1; t >= 0; t--) {
                const int ent = 7 * (k0 + t) + 5 + ((c >> t) & 1);
                const double T = tabT[ent];
                acc = F::kProd ? acc * T : acc + T;
                if (SEPJ) jac *= tabJ[ent];
            }
            cornA[idx] = acc;
            if (SEPJ) cornJ[idx] = jac;
        }
        __syncwarp();
        double sI = 0.0, sP = 0.0;
        for (int q = lane; q < NP - NCORN; q += 32) {                    // phase 2a: centre, axis and pair points
            const unsigned e = s_pt[q];
            double acc = finit, jac = 1.0;
#pragma unroll
            for (int k = D - 1; k >= 0; k--) {
                const int sel = (e >> (3 * k)) & 7u;
                const double T = tabT[7 * k + sel];
                acc = F::kProd ? acc * T : acc + T;
                if (SEPJ) { const double J = tabJ[7 * k + sel]; jac = (k == D - 1) ? J : jac * J; }
            }
            const double f = F::sep_finish(acc, par) * (ALLFIN ? jacc : jac);
            if (q < NAX) fst[q] = f;
            const unsigned cls = e >> 24;
            sI = fma(s_w[cls], f, sI); sP = fma(s_w[5 + cls], f, sP);
        }
        for (int r = lane; r < NCORN; r += 32) {                         // phase 2b: corner points (rule weight class 4)
            const int lo = r & (NL - 1), hi = NL + (r >> DL);
            double acc = F::kProd ? finit * (cornA[hi] * cornA[lo]) : finit + (cornA[hi] + cornA[lo]);
            const double jac = SEPJ ? cornJ[hi] * cornJ[lo] : 1.0;
            const double f = F::sep_finish(acc, par) * (ALLFIN ? jacc : jac);
            sI = fma(s_w[4], f, sI); sP = fma(s_w[9], f, sP);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sI += __shfl_xor_sync(0xffffffffu, sI, off);
            sP += __shfl_xor_sync(0xffffffffu, sP, off);
        }
        __syncwarp();
        Iv = 0.0; Er = 0.0; kdiv = 0;
        if (lane == 0) {
            double V = hh[0];
#pragma unroll
            for (int k = 1; k < D; k++) V *= hh[k];
            Iv = V * sI;
            const double Ip = V * sP;
            Er = fabs(Iv - Ip);
            double p10 = 1.0;
#pragma unroll
            for (int k = 0; k < D; k++) p10 *= 10.0;
            const double deltaf = Er / (p10 * V), twelvef1 = 12 * fst[0];
            int kd = 0; double maxdd = 0.0, hkd = hh[0];
#pragma unroll
            for (int k = 0; k < D; k++) {
                const double f2k = fst[1 + 2 * k] + fst[2 + 2 * k];
                const double f3k = fst[1 + 2 * D + 2 * k] + fst[2 + 2 * D + 2 * k];
                const double dd = fabs(f3k + twelvef1 - 7 * f2k);
                const double delta = dd - maxdd;
                if (delta > deltaf) { kd = k; maxdd = dd; hkd = hh[k]; }
                else if (fabs(delta) <= deltaf && hh[k] > hkd) { kd = k; hkd = hh[k]; }
            }
            kdiv = kd;
        }
    }
};

// evaluate: one warp per todo box (WarpRule above)
template <int FAM, int D, bool ALLFIN>
__global__ void __launch_bounds__(kBlock)
eval_kernel(Params p) {
    using W = WarpRule<FAM, D, ALLFIN>;
    using F = ib::Family<FAM, D>;
    constexpr int NPAR = W::NPAR, NP = W::NP;
    constexpr int RECW = 2 * D + 2;
    const State *st = p.st;
    if (st->done) return;
    __shared__ double s_par[W::kProb];
    __shared__ double s_w[10];
    __shared__ unsigned s_pt[NP];
    __shared__ double s_stage[kWarps][W::kStage];
    __shared__ double s_red[3 * kWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *par = s_par, *lbs = s_par + NPAR, *ubs = lbs + D, *dens = ubs + D;
    for (int t = threadIdx.x; t < NPAR; t += kBlock) par[t] = p.par[t];
    for (int k = threadIdx.x; k < D; k += kBlock) { lbs[k] = p.lb[k]; ubs[k] = p.ub[k]; dens[k] = (p.ub[k] - p.lb[k]) * 0.5; }
    W::setup_tables(s_w, s_pt, threadIdx.x, kBlock);
    __syncthreads();
    double jacc = 1.0;
    if (ALLFIN) {
#pragma unroll
        for (int k = 0; k < D; k++) jacc = (k == 0) ? dens[0] : jacc * dens[k];
    }
    const double finit = F::sep_init(par);
    double *stage = s_stage[wib], *box = W::box(stage);

    double sumI = 0.0, sumE = 0.0, maxE = 0.0;           // lane 0 accumulates this warp's boxes in todo order
    const long long ntodo = st->ntodo;
    const long long nwarps = (long long)gridDim.x * kWarps;
    for (long long t = ((long long)blockIdx.x * kWarps + wib) * p.world + p.rank; t < ntodo; t += nwarps * p.world) {
        const unsigned slot = p.todo[t];
        double *r = p.rec + (size_t)slot * RECW;
        __syncwarp();
        if (lane < 2 * D) box[lane] = r[lane];
        __syncwarp();
        double Iv, Er; int kd;
        W::eval(s_par, p.tr, finit, jacc, s_w, s_pt, stage, lane, Iv, Er, kd);
        if (lane == 0) {
            r[2 * D] = Iv; r[2 * D + 1] = (double)kd;
            p.err[slot] = Er;
            if (p.world > 1) {                                   // t == rank (mod world): entry t / world of this rank's list
                double2 *ob = reinterpret_cast<double2 *>(p.outbox + (size_t)(t / p.world) * 4);
                ob[0] = make_double2(__longlong_as_double((long long)slot), Iv);
                ob[1] = make_double2(Er, (double)kd);
            }
            sumI += Iv; sumE += Er; maxE = fmax(maxE, Er);       // fmax drops NaN: non-finite E is caught through sumE
        }
    }
    block_reduce3(sumI, sumE, maxE, s_red);
    if (threadIdx.x == 0) { double *o = p.partial + (size_t)blockIdx.x * 4; o[0] = sumI; o[1] = sumE; o[2] = maxE; }
}

// exponent histogram of the stored error estimates (integer atomics: order-independent, exact)
static __global__ void __launch_bounds__(kBlock)
hist_kernel(Params p) {
    const State *st = p.st;
    if (st->done) return;
    __shared__ unsigned s_h[kBins];
    for (int b = threadIdx.x; b < kBins; b += kBlock) s_h[b] = 0u;
    __syncthreads();
    const long long n = st->nbox;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
        atomicAdd(&s_h[(unsigned)((unsigned long long)__double_as_longlong(p.err[i]) >> 52) & 0x7ffu], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < kBins; b += kBlock) if (s_h[b]) atomicAdd(&p.hist[b], (unsigned long long)s_h[b]);
}

// single-block fixed-shape tree over (v0 + v1, max v2): same association whatever the data
__device__ __forceinline__ void fin_reduce3(double &v0, double &v1, double &v2, double *sh /* 3*32 */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v0 += __shfl_down_sync(0xffffffffu, v0, off);
        v1 += __shfl_down_sync(0xffffffffu, v1, off);
        v2 = fmax(v2, __shfl_down_sync(0xffffffffu, v2, off));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sh[w] = v0; sh[32 + w] = v1; sh[64 + w] = v2; }
    __syncthreads();
    if (w == 0) {
        v0 = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0; v1 = l < (int)(blockDim.x >> 5) ? sh[32 + l] : 0.0; v2 = l < (int)(blockDim.x >> 5) ? sh[64 + l] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v0 += __shfl_down_sync(0xffffffffu, v0, off);
            v1 += __shfl_down_sync(0xffffffffu, v1, off);
            v2 = fmax(v2, __shfl_down_sync(0xffffffffu, v2, off));
        }
    }
}

// number of boxes of a round that rank r evaluates (todo indices t == r mod world)
__device__ __forceinline__ long long share_of(long long ntodo, int r, int world) { return ntodo > r ? (ntodo - r + world - 1) / world : 0; }

// copy this rank's result list into every peer's inbox: coalesced 16-byte stores over NVLink
static __global__ void __launch_bounds__(kBlock)
publish_kernel(Params p) {
    const State *st = p.st;
    if (st->done) return;
    const long long n2 = 2 * share_of(st->ntodo, p.rank, p.world);                 // double2 elements
    const size_t half = (size_t)(st->round & 1) * p.world * p.obcap * 4;
    const double2 *src = reinterpret_cast<const double2 *>(p.outbox);
    for (int q = 0; q < p.world; q++) {
        if (q == p.rank) continue;
        double2 *dst = reinterpret_cast<double2 *>(p.inbox_peer[q] + half + (size_t)p.rank * p.obcap * 4);
        for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n2; i += (long long)gridDim.x * kBlock) dst[i] = src[i];
    }
}

// after the round's all-gather: scatter the peers' results into this rank's store
static __global__ void __launch_bounds__(kBlock)
apply_kernel(Params p, int recw) {
    const State *st = p.st;
    if (st->done) return;
    const size_t half = (size_t)(st->round & 1) * p.world * p.obcap * 4;
    for (int q = 0; q < p.world; q++) {
        if (q == p.rank) continue;
        const long long n = share_of(st->ntodo, q, p.world);
        const double2 *src = reinterpret_cast<const double2 *>(p.inbox + half + (size_t)q * p.obcap * 4);
        for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
            const double2 e0 = src[2 * i], e1 = src[2 * i + 1];
            const size_t slot = (size_t)__double_as_longlong(e0.x);
            double *r = p.rec + slot * recw;
            r[recw - 2] = e0.y; r[recw - 1] = e1.y;
            p.err[slot] = e1.x;
        }
    }
}

// fold the block partials of this rank's share of the round into the exchange buffer (one block of kFin threads)
static __global__ void __launch_bounds__(kFin)
eval_finish_kernel(Params p) {
    State *st = p.st;
    if (st->done) return;
    __shared__ double sh[96];
    double sI = 0.0, sE = 0.0, mE = 0.0;
    for (int b = threadIdx.x; b < p.grid; b += kFin) { sI += p.partial[4 * b]; sE += p.partial[4 * b + 1]; mE = fmax(mE, p.partial[4 * b + 2]); }
    fin_reduce3(sI, sE, mE, sh);
    if (threadIdx.x == 0) {
        double *s = p.sendbuf;
        s[0] = sI; s[1] = sE; s[2] = mE; s[3] = 0.0; s[4] = 0.0; s[5] = 0.0; s[6] = 0.0; s[7] = 0.0;
        if (p.world == 1) { for (int k = 0; k < kXchg; k++) p.recvbuf[k] = s[k]; }
    }
}

// global reduction in rank order + HCubature's stopping test + next threshold (one block of kFin threads)
static __global__ void __launch_bounds__(kFin)
decide_kernel(Params p) {
    State *st = p.st;
    if (st->done) return;
    __shared__ double s_cnt[kBins];
    __shared__ int s_bmin, s_bmax;
    if (threadIdx.x == 0) { s_bmin = kBins; s_bmax = -1; }
    __syncthreads();
    for (int b = threadIdx.x; b < kBins; b += kFin) {            // replicated store: the local histogram is the global one
        const unsigned long long c = p.hist[b];
        s_cnt[b] = (double)c; p.hist[b] = 0ull;
        if (c) { atomicMin(&s_bmin, b); atomicMax(&s_bmax, b); }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double dI = 0.0, dE = 0.0, mx = 0.0;
    for (int r = 0; r < p.world; r++) {                      // rank order: the same sums on every rank
        const double *v = p.recvbuf + (size_t)r * kXchg;
        dI += v[0]; dE += v[1]; mx = fmax(mx, v[2]);
    }
    st->I += dI; st->E += dE;
    st->napplied += st->ntodo;
    const double I = st->I, E = st->E, ap = (double)st->napplied, full = (double)st->full, nb = (double)st->nbox;
    st->I_g = I; st->E_g = E; st->gapplied = ap;
    st->round++;
    st->ntodo = 0;
    double rtol = p.rtol;
    if (rtol == 0.0 && p.atol == 0.0) rtol = 1.4901161193847656e-08;
    const double tol = fmax(rtol * fabs(I), p.atol);
    if (E <= tol) { st->done = 1; st->status = IB200_ST_CONVERGED; return; }
    if (!isfinite(E)) { st->done = 1; st->status = IB200_ST_NONFINITE; return; }
    if (ap >= (double)p.max_boxes || full > 0.0) { st->done = 1; st->status = IB200_ST_MAXEVALS; return; }
    if (p.switch_boxes > 0 && nb >= (double)p.switch_boxes && (E <= p.switch_factor * tol || nb >= p.switch_cap)) {
        st->done = 1; st->status = kStatusHandover; return;
    }
    st->bound = fmax(mx, st->theta);          // boxes left unsplit by the previous round are below its threshold
    double theta = 0.5 * st->bound;
    // The three scans below are written for all kBins bins (oracle: orc_hcubature_rounds); empty bins change no sum, so
    // only the occupied range [b0, b1] is walked, serially and in the same order -- the decision is bit-identical.
    const int b0 = s_bmin, b1 = s_bmax < kBins - 2 ? s_bmax : kBins - 2;
    // (1) longest ascending prefix of bins whose lower-edge mass stays <= tol: bins 0..bh may stay unsplit
    //     edge(b) = 0 for b = 0, 2^(b-1023) otherwise
    int bh = b0 - 1; double low = 0.0; bool broke = false;
    for (int b = b0; b <= b1; b++) {
        const double c = s_cnt[b];
        if (c != 0.0) { low += c * (b == 0 ? 0.0 : scalbn(1.0, b - 1023)); if (!(low <= tol)) { broke = true; break; } }
        bh = b;
    }
    if (!broke) bh = kBins - 2;
    // (2) budget: lowest bin bb such that bisecting every box in bins >= bb fits the remaining rule applications
    const double remaining = (double)p.max_boxes - ap;
    int bb = b1 + 1; double above = 0.0; broke = false;         // the (empty) bins above b1 have already lowered bb to b1 + 1
    for (int b = b1; b >= b0; b--) {
        above += s_cnt[b];
        if (2.0 * above > remaining) { broke = true; break; }
        bb = b;
    }
    if (!broke) bb = 0;
    // (3) selectivity: at most `frac` of all stored boxes per round, largest errors first -- keeps the refinement
    //     close to the reference's worst-first order when the budget, not the tolerance, ends the solve
    int bf = b1 + 1; double top = 0.0; broke = false;
    for (int b = b1; b >= b0; b--) {
        top += s_cnt[b];
        if (top > p.frac * nb) { broke = true; break; }
        bf = b;
    }
    if (!broke) bf = 0;
    if (bf > bb) bb = bf;
    const int bs = (bh + 1 > bb) ? bh + 1 : bb;               // bisect bins >= bs
    const double theta_h = bs <= 0 ? 0.0 : (bs >= kBins - 1 ? 1.7976931348623157e308 : scalbn(1.0, bs - 1023));
    st->theta = fmin(theta, theta_h);
}

// split pass 1: per-block count of boxes with err >= theta, and the sums of their (I, E)
template <int D>
__global__ void __launch_bounds__(kBlock)
count_kernel(Params p) {
    constexpr int RECW = 2 * D + 2;
    const State *st = p.st;
    if (st->done) return;
    __shared__ double s_red[3 * kWarps];
    __shared__ long long s_cnt[kWarps];
    const long long n = st->nbox;
    const double theta = st->theta;
    // contiguous chunk per block so that slots are assigned in box order
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    long long cnt = 0; double sI = 0.0, sE = 0.0, dummy = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += kBlock) {
        const double e = p.err[i];
        if (e >= theta) { cnt++; sI += p.rec[(size_t)i * RECW + 2 * D]; sE += e; }
    }
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    block_reduce3(sI, sE, dummy, s_red);
    if (threadIdx.x == 0) {
        long long c = 0;
        for (int k = 0; k < kWarps; k++) c += s_cnt[k];
        p.bcount[blockIdx.x] = c;
        double *o = p.partial + (size_t)blockIdx.x * 4; o[0] = sI; o[1] = sE;
    }
}

// split pass 2: exclusive scan of the block counts, capacity check, running sums (one block of kFin threads, grid <= kFin)
static __global__ void __launch_bounds__(kFin)
scan_kernel(Params p) {
    State *st = p.st;
    if (st->done) return;
    __shared__ double sh[96];
    __shared__ long long s_w[32];
    const int t = threadIdx.x, l = t & 31, w = t >> 5;
    const long long c = t < p.grid ? p.bcount[t] : 0;
    long long inc = c;                                        // inclusive warp scan
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, inc, off); if (l >= off) inc += v; }
    if (l == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        long long v = s_w[l], iv = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, iv, off); if (l >= off) iv += u; }
        s_w[l] = iv - v;                                      // exclusive prefix of the warp totals
    }
    __syncthreads();
    const long long excl = s_w[w] + inc - c;
    if (t < p.grid) p.bcount[t] = excl;
    double sI = t < p.grid ? p.partial[4 * t] : 0.0, sE = t < p.grid ? p.partial[4 * t + 1] : 0.0, d = 0.0;
    fin_reduce3(sI, sE, d, sh);
    __shared__ long long s_total;
    if (t == p.grid - 1) s_total = excl + c;
    __syncthreads();
    if (t == 0) {
        const long long run = s_total;
        p.bcount[p.grid] = run;
        if (st->nbox + run > p.cap) {             // store full: leave the boxes alone, tell everybody in the next exchange
            st->full = 1; st->nsplit = 0; st->ntodo = 0;
        } else {
            st->nsplit = run;
            st->I -= sI; st->E -= sE;
        }
    }
}

// split pass 3: children to their deterministic slots, todo list for the next round
template <int D>
__global__ void __launch_bounds__(kBlock)
scatter_kernel(Params p) {
    constexpr int RECW = 2 * D + 2;
    State *st = p.st;
    if (st->done || st->full) return;
    __shared__ long long s_woff[kWarps + 1];
    const long long n = st->nbox;
    const double theta = st->theta;
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long base = p.bcount[blockIdx.x];
    for (long long i0 = lo; i0 < hi; i0 += kBlock) {
        const long long i = i0 + threadIdx.x;
        const bool flag = i < hi && p.err[i] >= theta;
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_woff[w + 1] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { s_woff[0] = 0; for (int k = 0; k < kWarps; k++) s_woff[k + 1] += s_woff[k]; }
        __syncthreads();
        if (flag) {
            const long long rank_in = base + s_woff[w] + __popc(bal & ((1u << lane) - 1u));
            const long long child = n + rank_in;
            double *r = p.rec + (size_t)i * RECW, *c = p.rec + (size_t)child * RECW;
            const int kd = (int)r[2 * D + 1];
            double a[D], b[D];
#pragma unroll
            for (int k = 0; k < D; k++) { a[k] = r[k]; b[k] = r[D + k]; }
            double akd = 0.0, bkd = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) if (k == kd) { akd = a[k]; bkd = b[k]; }
            const double wd = (bkd - akd) * 0.5;
            // child 0 (upper half, a[kd] += w) stays in the parent's slot; child 1 (lower half, b[kd] -= w) is appended
#pragma unroll
            for (int k = 0; k < D; k++) { c[k] = a[k]; c[D + k] = (k == kd) ? bkd - wd : b[k]; }
            c[2 * D] = 0.0; c[2 * D + 1] = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) if (k == kd) r[k] = akd + wd;
            p.err[child] = 0.0;
            p.todo[2 * rank_in] = (unsigned)i;
            p.todo[2 * rank_in + 1] = (unsigned)child;
        }
        base += s_woff[kWarps];
        __syncthreads();
    }
}

static __global__ void split_finish_kernel(Params p) {
    State *st = p.st;
    if (st->done || st->full || threadIdx.x != 0) return;
    st->nbox += st->nsplit;
    st->ntodo = 2 * st->nsplit;
}

// final re-sum over the whole store (HCubature's closing sum), fixed order
template <int D>
__global__ void __launch_bounds__(kBlock)
resum_kernel(Params p) {
    constexpr int RECW = 2 * D + 2;
    __shared__ double s_red[3 * kWarps];
    const long long n = p.st->nbox;
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    double sI = 0.0, sE = 0.0, d = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += kBlock) { sI += p.rec[(size_t)i * RECW + 2 * D]; sE += p.err[i]; }
    block_reduce3(sI, sE, d, s_red);
    if (threadIdx.x == 0) { double *o = p.partial + (size_t)blockIdx.x * 4; o[0] = sI; o[1] = sE; }
}
static __global__ void __launch_bounds__(kFin)
resum_finish_kernel(Params p, double *out /* I, E, applied, boxes, rounds, status */) {
    State *st = p.st;
    __shared__ double sh[96];
    double sI = 0.0, sE = 0.0, d = 0.0;
    for (int b = threadIdx.x; b < p.grid; b += kFin) { sI += p.partial[4 * b]; sE += p.partial[4 * b + 1]; }
    fin_reduce3(sI, sE, d, sh);
    if (threadIdx.x == 0) {
        out[0] = sI; out[1] = sE; out[2] = (double)st->napplied; out[3] = (double)st->nbox; out[4] = (double)st->round; out[5] = (double)st->status;
    }
}

// ---- two-level solve: the weights of the level-1 boxes, sum_i err_i^alpha, in a fixed order (chunks of the store per block, block
//      partials folded by one block): the same double on every rank
static __global__ void __launch_bounds__(kBlock)
powsum_kernel(Params p, double alpha) {
    __shared__ double s_red[3 * kWarps];
    const long long n = p.st->nbox;
    const long long chunk = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    double s = 0.0, d1 = 0.0, d2 = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += kBlock) s += pow(p.err[i], alpha);
    block_reduce3(s, d1, d2, s_red);
    if (threadIdx.x == 0) p.partial[(size_t)blockIdx.x * 4] = s;
}
static __global__ void __launch_bounds__(kFin)
powsum_finish_kernel(Params p, double *out) {
    __shared__ double sh[96];
    double s = 0.0, d1 = 0.0, d2 = 0.0;
    for (int b = threadIdx.x; b < p.grid; b += kFin) s += p.partial[4 * b];
    fin_reduce3(s, d1, d2, sh);
    if (threadIdx.x == 0) out[0] = s;
}
// level-2 results -> per-chunk sums.  A chunk = kSubChunk consecutive boxes of the level-1 store, solved on ONE rank (chunk c on rank
// c mod world), summed by one warp in a fixed order: the chunk sums, and so the final result, do not depend on the number of ranks.
constexpr int kSubChunk = 64;
static __global__ void __launch_bounds__(kBlock)
chunk_sum_kernel(const double *sub_out /* [nlocal][4] */, long long nchunk_local, double *chunk_out /* [nchunk_local][4] */) {
    const int lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (c >= nchunk_local) return;
    const double *o = sub_out + (size_t)c * kSubChunk * 4;
    double sI = 0.0, sE = 0.0, sA = 0.0, mS = 0.0;
    for (int j = lane; j < kSubChunk; j += 32) { sI += o[4 * j]; sE += o[4 * j + 1]; sA += o[4 * j + 2]; mS = fmax(mS, o[4 * j + 3]); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sI += __shfl_xor_sync(0xffffffffu, sI, off); sE += __shfl_xor_sync(0xffffffffu, sE, off);
        sA += __shfl_xor_sync(0xffffffffu, sA, off); mS = fmax(mS, __shfl_xor_sync(0xffffffffu, mS, off));
    }
    if (lane == 0) { double *q = chunk_out + (size_t)c * 4; q[0] = sI; q[1] = sE; q[2] = sA; q[3] = mS; }
}
// all chunks in chunk order (chunk c sits at [c % world][c / world] of the gathered buffer), one block, fixed shape
static __global__ void __launch_bounds__(kFin)
chunk_total_kernel(const double *gathered, long long nchunk, long long nchunk_local, int world, double *out /* I, E, applications, status */) {
    __shared__ double sh[96];
    __shared__ double sh2[96];
    const long long per = (nchunk + kFin - 1) / kFin;                        // consecutive chunks per thread
    const long long lo = (long long)threadIdx.x * per, hi = (lo + per < nchunk) ? lo + per : nchunk;
    double sI = 0.0, sE = 0.0, sA = 0.0, mS = 0.0, d = 0.0;
    for (long long c = lo; c < hi; c++) {
        const double *q = gathered + ((size_t)(c % world) * nchunk_local + (size_t)(c / world)) * 4;
        sI += q[0]; sE += q[1]; sA += q[2]; mS = fmax(mS, q[3]);
    }
    fin_reduce3(sI, sE, mS, sh);
    fin_reduce3(sA, d, d, sh2);
    if (threadIdx.x == 0) { out[0] = sI; out[1] = sE; out[2] = sA; out[3] = mS; }
}

}  // namespace hs
