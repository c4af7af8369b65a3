// batch_bridge.cu -- BatchIntegralFunction bridge: adaptive integration of ONE integral whose
// integrand is evaluated by the CALLER, a batch of points at a time.
//
// Reference: BatchIntegralFunction (src/Integrals.jl:274-315: QuadGKJL hands QuadGK a BatchIntegrand, f(y, u) on a vector
// of points; test/interface_tests.jl:189-258 "Batched ... Integrands").  The registered-family kernels cannot run a Julia
// (or Python) closure, but a closure that maps `pts` (dim x n) to n values can run anywhere -- on the host, or as a
// CuArray / torch broadcast on the same GPU.  So the engine keeps what is data-parallel and stateful on the device -- the
// box store, the rule's points, the weighted sums, error estimates, split axes and the whole refinement bookkeeping of
// hcub_rounds.cuh -- and exposes the rounds as a pull-style C ABI without callbacks:
//
//   ib200_batch_open(ctx, dim, transform, lb, ub, reltol, abstol, max_boxes, store_boxes)
//   loop:  ib200_batch_next(ctx, &n)          n = points of the next round, 0 = finished
//          ib200_batch_points(ctx, x, n)      x[dim x n] (column-major, ORIGINAL coordinates) <- the round's points
//          ib200_batch_values(ctx, y, n)      y[n] = f(x[:, j]); combine, decide, split
//   ib200_batch_result(ctx, &I, &E, stats)
//
// x and y may be host or device pointers (device: no copy at all).  The automatic ChangeOfVariables (src/common.jl:97-103)
// stays inside the engine: points are handed out in the caller's coordinates u(t), and the Jacobian of every point is
// kept on the device and applied when the values come back (infinity_handling.jl:11-25).
//
// Rule and refinement: Genz-Malik degree 7/5 per box for dim >= 2, GK(7,15) per segment in 1-D (HCubature's rules,
// oracle gm_rule / gk_rule_hc), refined in the rounds of hcub_single.cu -- every box the reference would provably bisect
// is bisected in the same round, so the caller sees O(log #boxes) large batches instead of one 15- or 57-point call per
// refinement step.  Oracle: orc_hcubature_rounds with the same integrand; same rounds, boxes and rule applications,
// (I, E) equal to summation-order rounding.
#include "hcub_rounds.cuh"
#include <vector>

namespace bb {

using hs::Params;
using hs::State;
using hs::kBlock;
using hs::kWarps;
using hs::kFin;
using hs::kBins;
using hs::kXchg;

constexpr int kBufs = 12;

// host-side state of an open bridge solve; the device buffers are the bridge's own (a solve spans several C-ABI calls,
// so it cannot live in the context's per-call scratch) and are reused by the next solve
struct Bridge {
    Params p;
    int dim = 0;
    int phase = 0;              // 0 closed, 1 round ready (next / points), 2 points handed out (values), 3 finished
    long long ntodo = 0;        // boxes of the round in flight
    int np = 0;                 // points per box
    int status = 0;
    struct Buf { void *p = nullptr; size_t cap = 0; } buf[kBufs];
    void *get(int i, size_t bytes) {
        Buf &b = buf[i];
        if (bytes == 0) bytes = 8;
        if (b.cap < bytes) {
            if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
            if (cudaMalloc(&b.p, bytes) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return nullptr; }
            b.cap = bytes;
        }
        return b.p;
    }
    void release() { for (auto &b : buf) if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; } }
};
enum { B_LB = 0, B_UB, B_REC, B_ERR, B_TODO, B_PARTIAL, B_BCOUNT, B_XB, B_JAC, B_X, B_Y, B_OUT };

// Genz-Malik point table: 3 bits per axis = abscissa index (0 centre, 1/2 +-l2, 3/4 +-l3, 5/6 +-l5); class in bits 24-26
template <int D>
__device__ __forceinline__ void fill_point_table(unsigned *pt) {
    constexpr int NP = hc::gm_npoints(D);
    for (int q = threadIdx.x; q < NP; q += blockDim.x) {
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
        pt[q] = e | (cls << 24);
    }
}

// the round's points: one thread per (box, point).  x[j * D + k] = coordinate k of point j in the caller's variables,
// jac[j] = prod_k du_k/dt_k (left to right, infinity_handling.jl:11-17); point j = (todo index) * NP + (point of the rule)
template <int D>
__global__ void __launch_bounds__(kBlock)
points_kernel(Params p, double *__restrict__ x, double *__restrict__ jac) {
    constexpr int NP = hc::gm_npoints(D);
    constexpr int RECW = 2 * D + 2;
    const State *st = p.st;
    if (st->done) return;
    __shared__ unsigned s_pt[NP];
    __shared__ double s_lb[D], s_ub[D];
    if (D >= 2) fill_point_table<D>(s_pt);
    for (int k = threadIdx.x; k < D; k += kBlock) { s_lb[k] = p.lb[k]; s_ub[k] = p.ub[k]; }
    __syncthreads();
    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
    const long long n = st->ntodo * NP;
    for (long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < n; j += (long long)gridDim.x * kBlock) {
        const long long t = j / NP;
        const int q = (int)(j - t * NP);
        const double *r = p.rec + (size_t)p.todo[t] * RECW;
        double J = 1.0;
        if constexpr (D == 1) {
            // HCubature's 1-D rule: GK(7,15) about the centre with signed half-width (oracle gk_rule_hc)
            const double a = r[0], b = r[1];
            const double c = (a + b) * 0.5, Dl = (b - a) * 0.5;
            const double dx = q < 14 ? Dl * hc::c_x[q < 7 ? q : q - 7] : 0.0;
            const double tt = q < 7 ? c + dx : (q < 14 ? c - dx : c);
            ib::AxisMap m; ib::make_axis(s_lb[0], s_ub[0], m);
            double u;
            ib::t2ujac(p.tr, tt, m, u, J);
            x[j] = u;
        } else {
            const unsigned e = s_pt[q];
#pragma unroll
            for (int k = 0; k < D; k++) {
                const int o = (e >> (3 * k)) & 7u;
                const double a = r[k], b = r[D + k];
                const double c = 0.5 * (a + b), h = 0.5 * fabs(b - a);
                const double lam = o == 0 ? 0.0 : (o <= 2 ? l2 : (o <= 4 ? l3 : l5));
                const double dl = h * lam;
                const double tt = o == 0 ? c : ((o & 1) ? c + dl : c - dl);
                ib::AxisMap m; ib::make_axis(s_lb[k], s_ub[k], m);
                double u, jk;
                ib::t2ujac(p.tr, tt, m, u, jk);
                x[(size_t)j * D + k] = u;
                J = (k == 0) ? jk : J * jk;
            }
        }
        jac[j] = J;
    }
}

// the caller's values -> (I, E, split axis) of every box of the round: one warp per box, the points over the lanes
template <int D>
__global__ void __launch_bounds__(kBlock)
combine_kernel(Params p, const double *__restrict__ y, const double *__restrict__ jac) {
    constexpr int NP = hc::gm_npoints(D);
    constexpr int NAX = 1 + 4 * D;
    constexpr int RECW = 2 * D + 2;
    const State *st = p.st;
    if (st->done) return;
    __shared__ double s_w[10];
    __shared__ unsigned s_pt[NP];
    __shared__ double s_fst[kWarps][NAX];
    __shared__ double s_red[3 * kWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (D >= 2) {
        if (threadIdx.x == 0) {
            const double twon = (double)(1 << D);
            s_w[0] = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0); s_w[1] = twon * (980 / 6561.0);
            s_w[2] = twon * ((1820 - 400 * D) / 19683.0); s_w[3] = twon * (200 / 19683.0); s_w[4] = 6859 / 19683.0;
            s_w[5] = twon * ((729 - 950 * D + 50 * D * D) / 729.0); s_w[6] = twon * (245 / 486.0);
            s_w[7] = twon * ((265 - 100 * D) / 1458.0); s_w[8] = twon * (25 / 729.0); s_w[9] = 0.0;
        }
        fill_point_table<D>(s_pt);
    }
    __syncthreads();
    double *fst = s_fst[wib];
    double sumI = 0.0, sumE = 0.0, maxE = 0.0;           // lane 0 accumulates this warp's boxes in todo order
    const long long ntodo = st->ntodo;
    const long long nwarps = (long long)gridDim.x * kWarps;
    for (long long t = (long long)blockIdx.x * kWarps + wib; t < ntodo; t += nwarps) {
        const unsigned slot = p.todo[t];
        double *r = p.rec + (size_t)slot * RECW;
        const double *yt = y + (size_t)t * NP, *jt = jac + (size_t)t * NP;
        double sI = 0.0, sP = 0.0;
        __syncwarp();
        for (int q = lane; q < NP; q += 32) {
            const double f = yt[q] * jt[q];
            if constexpr (D == 1) {
                const double wk = q < 14 ? hc::c_w[q < 7 ? q : q - 7] : hc::c_w[7];
                const int i7 = q < 7 ? q : q - 7;
                const double wg = q < 14 ? ((i7 & 1) ? hc::c_wg[i7 >> 1] : 0.0) : hc::c_wg[3];
                sI = fma(wk, f, sI); sP = fma(wg, f, sP);
            } else {
                if (q < NAX) fst[q] = f;
                const unsigned cls = s_pt[q] >> 24;
                sI = fma(s_w[cls], f, sI); sP = fma(s_w[5 + cls], f, sP);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sI += __shfl_xor_sync(0xffffffffu, sI, off);
            sP += __shfl_xor_sync(0xffffffffu, sP, off);
        }
        __syncwarp();
        if (lane == 0) {
            double Iv, Er; int kd = 0;
            if constexpr (D == 1) {
                const double Dl = (r[1] - r[0]) * 0.5;
                Iv = sI * Dl;
                Er = fabs(Iv - sP * Dl);
            } else {
                double hh[D];
#pragma unroll
                for (int k = 0; k < D; k++) hh[k] = 0.5 * fabs(r[D + k] - r[k]);
                double V = hh[0];
#pragma unroll
                for (int k = 1; k < D; k++) V *= hh[k];
                Iv = V * sI;
                Er = fabs(Iv - V * sP);
                double p10 = 1.0;
#pragma unroll
                for (int k = 0; k < D; k++) p10 *= 10.0;
                const double deltaf = Er / (p10 * V), twelvef1 = 12 * fst[0];
                double maxdd = 0.0, hkd = hh[0];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const double f2k = fst[1 + 2 * k] + fst[2 + 2 * k];
                    const double f3k = fst[1 + 2 * D + 2 * k] + fst[2 + 2 * D + 2 * k];
                    const double dd = fabs(f3k + twelvef1 - 7 * f2k);
                    const double delta = dd - maxdd;
                    if (delta > deltaf) { kd = k; maxdd = dd; hkd = hh[k]; }
                    else if (fabs(delta) <= deltaf && hh[k] > hkd) { kd = k; hkd = hh[k]; }
                }
            }
            r[2 * D] = Iv; r[2 * D + 1] = (double)kd;
            p.err[slot] = Er;
            sumI += Iv; sumE += Er; maxE = fmax(maxE, Er);       // fmax drops NaN: non-finite E is caught through sumE
        }
    }
    hs::block_reduce3(sumI, sumE, maxE, s_red);
    if (threadIdx.x == 0) { double *o = p.partial + (size_t)blockIdx.x * 4; o[0] = sumI; o[1] = sumE; o[2] = maxE; }
}

static void bridge_free(void *b) {
    Bridge *B = static_cast<Bridge *>(b);
    if (!B) return;
    B->release();
    delete B;
}

// read the device state block (synchronises the stream)
static int32_t read_state(ib200_ctx *ctx, Bridge *B, State &h) {
    IB_CUDA(ctx, cudaMemcpyAsync(&h, B->p.st, sizeof(State), cudaMemcpyDeviceToHost, ctx->stream));
    IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}

template <int D>
static int32_t do_open(ib200_ctx *ctx, Bridge *B) {
    hs::init_kernel<D, false><<<1, 32, 0, ctx->stream>>>(B->p);
    IB_LAUNCH_CHECK(ctx);
    return IB200_OK;
}
template <int D>
static int32_t do_points(ib200_ctx *ctx, Bridge *B, double *x, double *jac) {
    points_kernel<D><<<B->p.grid, kBlock, 0, ctx->stream>>>(B->p, x, jac);
    IB_LAUNCH_CHECK(ctx);
    return IB200_OK;
}
template <int D>
static int32_t do_values(ib200_ctx *ctx, Bridge *B, const double *y, const double *jac) {
    Params &p = B->p;
    cudaStream_t s = ctx->stream;
    const int grid = p.grid;
    combine_kernel<D><<<grid, kBlock, 0, s>>>(p, y, jac);
    hs::eval_finish_kernel<<<1, kFin, 0, s>>>(p);
    hs::hist_kernel<<<grid, kBlock, 0, s>>>(p);
    hs::decide_kernel<<<1, kFin, 0, s>>>(p);
    hs::count_kernel<D><<<grid, kBlock, 0, s>>>(p);
    hs::scan_kernel<<<1, kFin, 0, s>>>(p);
    hs::scatter_kernel<D><<<grid, kBlock, 0, s>>>(p);
    hs::split_finish_kernel<<<1, 32, 0, s>>>(p);
    ctx->launches += 7;
    IB_LAUNCH_CHECK(ctx);
    return IB200_OK;
}
template <int D>
static int32_t do_result(ib200_ctx *ctx, Bridge *B, double *out6) {
    hs::resum_kernel<D><<<B->p.grid, kBlock, 0, ctx->stream>>>(B->p);
    hs::resum_finish_kernel<<<1, kFin, 0, ctx->stream>>>(B->p, out6);
    ctx->launches += 1;
    IB_LAUNCH_CHECK(ctx);
    return IB200_OK;
}

// dimension dispatch of the four steps
#define BB_DISPATCH(call)                                                                          \
    [&]() -> int32_t {                                                                             \
        int32_t rc__ = IB200_ERR_UNSUPPORTED;                                                      \
        ib200_dispatch_dim<IB200_MAX_DIM>(B->dim, [&](auto dc) { constexpr int D = decltype(dc)::value; rc__ = call; }); \
        return rc__;                                                                               \
    }()

}  // namespace bb

extern "C" int32_t ib200_batch_open(ib200_ctx *ctx, int32_t dim, int32_t transform, const double *lb, const double *ub,
                                    double reltol, double abstol, int64_t max_boxes, int64_t store_boxes) {
    using namespace bb;
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_batch_open: ctx is NULL");
    IB_REQUIRE(ctx, dim >= 1 && dim <= IB200_MAX_DIM, "ib200_batch_open: needs 1 <= dim <= 8");
    IB_REQUIRE(ctx, transform >= 0 && transform <= 2, "ib200_batch_open: unknown transform id");
    IB_REQUIRE(ctx, !(reltol < 0) && !(abstol < 0), "invalid negative tolerance");
    IB_REQUIRE(ctx, max_boxes >= 1, "ib200_batch_open: max_boxes must be positive");
    IB_REQUIRE(ctx, lb && ub, "ib200_batch_open: NULL data pointer");
    IB_REQUIRE(ctx, !ib200_is_device_ptr(lb) && !ib200_is_device_ptr(ub), "ib200_batch_open: lb, ub are host arrays of `dim` numbers");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->batch) { ctx->batch = new Bridge(); ctx->batch_free = bridge_free; }
    Bridge *B = static_cast<Bridge *>(ctx->batch);
    B->phase = 0; B->dim = dim; B->np = hc::gm_npoints(dim); B->ntodo = 0; B->status = 0;
    Params &p = B->p;
    memset(&p, 0, sizeof(p));
    p.tr = transform; p.rtol = reltol; p.atol = abstol; p.max_boxes = max_boxes;
    p.rank = 0; p.world = 1; p.frac = ctx->opt_single_round_fraction;
    p.grid = ctx->sm_count * 4;
    if (p.grid > kFin) p.grid = kFin;
    // box store: what the budget can create (a round is handed to the caller, so stores stay modest), else store_boxes
    long long want = store_boxes > 0 ? store_boxes : (max_boxes < (1LL << 22) ? max_boxes + 1024 : (1LL << 22));
    if (want > 4000000000LL) want = 4000000000LL;
    if (want < 16) want = 16;
    p.cap = want;
    double *dlb = (double *)B->get(B_LB, dim * sizeof(double)), *dub = (double *)B->get(B_UB, dim * sizeof(double));
    p.rec = (double *)B->get(B_REC, (size_t)p.cap * (2 * dim + 2) * sizeof(double));
    p.err = (double *)B->get(B_ERR, (size_t)p.cap * sizeof(double));
    p.todo = (unsigned *)B->get(B_TODO, (size_t)p.cap * sizeof(unsigned));
    p.partial = (double *)B->get(B_PARTIAL, (size_t)p.grid * 4 * sizeof(double));
    p.bcount = (long long *)B->get(B_BCOUNT, (size_t)(p.grid + 1) * sizeof(long long));
    double *xb = (double *)B->get(B_XB, (size_t)((kXchg + 1) * 2 + kBins) * sizeof(double) + sizeof(State) + 64);
    if (!dlb || !dub || !p.rec || !p.err || !p.todo || !p.partial || !p.bcount || !xb)
        return ib200_fail(ctx, IB200_ERR_OOM, "ib200_batch_open: box store allocation failed (pass a smaller store_boxes)");
    IB_CUDA(ctx, cudaMemcpyAsync(dlb, lb, dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    IB_CUDA(ctx, cudaMemcpyAsync(dub, ub, dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // lb, ub may be temporaries of the caller
    p.lb = dlb; p.ub = dub; p.par = nullptr;
    p.sendbuf = xb; p.recvbuf = xb + kXchg + 1;
    p.hist = reinterpret_cast<unsigned long long *>(xb + (kXchg + 1) * 2);
    p.st = reinterpret_cast<State *>(xb + (kXchg + 1) * 2 + kBins);
    p.outbox = p.inbox = nullptr; p.obcap = 0;
    IB_CUDA(ctx, cudaMemsetAsync(p.hist, 0, kBins * sizeof(unsigned long long), ctx->stream));
    int32_t rc = BB_DISPATCH(do_open<D>(ctx, B));
    if (rc) return rc == IB200_ERR_UNSUPPORTED ? ib200_fail(ctx, rc, "ib200_batch_open: dimension not compiled in") : rc;
    B->ntodo = 1;
    B->phase = 1;
    return IB200_OK;
}

extern "C" int32_t ib200_batch_next(ib200_ctx *ctx, int64_t *npoints) {
    using namespace bb;
    IB_REQUIRE(ctx, ctx != nullptr && npoints != nullptr, "ib200_batch_next: NULL argument");
    Bridge *B = static_cast<Bridge *>(ctx->batch);
    IB_REQUIRE(ctx, B && (B->phase == 1 || B->phase == 3), "ib200_batch_next: no round pending (call ib200_batch_open, or ib200_batch_values for the points already handed out)");
    *npoints = B->phase == 3 ? 0 : (int64_t)B->ntodo * B->np;
    return IB200_OK;
}

extern "C" int32_t ib200_batch_points(ib200_ctx *ctx, double *x, int64_t npoints) {
    using namespace bb;
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_batch_points: ctx is NULL");
    Bridge *B = static_cast<Bridge *>(ctx->batch);
    IB_REQUIRE(ctx, B && B->phase == 1, "ib200_batch_points: no round pending");
    IB_REQUIRE(ctx, x != nullptr && npoints == (int64_t)B->ntodo * B->np, "ib200_batch_points: x must hold dim x n doubles, n as returned by ib200_batch_next");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));
    double *jac = (double *)B->get(B_JAC, (size_t)npoints * sizeof(double));
    const bool dev = ib200_is_device_ptr(x);
    double *xd = dev ? x : (double *)B->get(B_X, (size_t)npoints * B->dim * sizeof(double));
    if (!jac || !xd) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_batch_points: point buffer allocation failed (lower max_boxes to bound the round size)");
    ib200_time_begin(ctx);
    int32_t rc = BB_DISPATCH(do_points<D>(ctx, B, xd, jac));
    if (rc) return rc;
    ib200_time_end(ctx);
    if (!dev) IB_CUDA(ctx, cudaMemcpyAsync(x, xd, (size_t)npoints * B->dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // the points are complete on return, whatever stream the caller evaluates on
    B->phase = 2;
    return IB200_OK;
}

extern "C" int32_t ib200_batch_values(ib200_ctx *ctx, const double *y, int64_t npoints) {
    using namespace bb;
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_batch_values: ctx is NULL");
    Bridge *B = static_cast<Bridge *>(ctx->batch);
    IB_REQUIRE(ctx, B && B->phase == 2, "ib200_batch_values: no points handed out (call ib200_batch_points first)");
    IB_REQUIRE(ctx, y != nullptr && npoints == (int64_t)B->ntodo * B->np, "ib200_batch_values: y must hold the n values of the points handed out");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double *yd = y;
    if (!ib200_is_device_ptr(y)) {
        double *s = (double *)B->get(B_Y, (size_t)npoints * sizeof(double));
        if (!s) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_batch_values: value buffer allocation failed");
        IB_CUDA(ctx, cudaMemcpyAsync(s, y, (size_t)npoints * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        yd = s;
    }
    ib200_time_begin(ctx);
    int32_t rc = BB_DISPATCH(do_values<D>(ctx, B, yd, (const double *)B->buf[B_JAC].p));
    if (rc) return rc;
    ib200_time_end(ctx);
    State h;
    if ((rc = read_state(ctx, B, h))) return rc;            // also: y (host) has been consumed when this returns
    // A round may bisect nothing: the selectivity rule can put the threshold above every stored error when few boxes share
    // the top histogram bin; the threshold then halves per round (hcub_single.cu just runs such rounds, the oracle counts
    // them too).  They need no integrand values, so they are run here, not handed to the caller.
    for (int empty = 0; !h.done && h.ntodo == 0; empty++) {
        if (empty > 4096) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_batch_values: refinement stalled (internal error)");
        if ((rc = BB_DISPATCH(do_values<D>(ctx, B, yd, (const double *)B->buf[B_JAC].p)))) return rc;
        if ((rc = read_state(ctx, B, h))) return rc;
    }
    B->ntodo = h.ntodo;
    B->status = h.status;
    B->phase = h.done ? 3 : 1;
    return IB200_OK;
}

extern "C" int32_t ib200_batch_result(ib200_ctx *ctx, double *out_I, double *out_E, int64_t *out_stats) {
    using namespace bb;
    IB_REQUIRE(ctx, ctx != nullptr && out_I != nullptr, "ib200_batch_result: NULL argument");
    Bridge *B = static_cast<Bridge *>(ctx->batch);
    IB_REQUIRE(ctx, B && B->phase == 3, "ib200_batch_result: the solve has not finished (ib200_batch_next must have returned 0 points)");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));
    double *out6 = (double *)B->get(B_OUT, 8 * sizeof(double));
    if (!out6) return ib200_fail(ctx, IB200_ERR_OOM, "scratch allocation failed");
    int32_t rc = BB_DISPATCH(do_result<D>(ctx, B, out6));
    if (rc) return rc;
    double res[6];
    IB_CUDA(ctx, cudaMemcpyAsync(res, out6, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_I = res[0];
    if (out_E) *out_E = res[1];
    if (out_stats) {
        out_stats[0] = (int64_t)res[2] * B->np;   // integrand evaluations
        out_stats[1] = (int64_t)res[2];           // rule applications
        out_stats[2] = (int64_t)res[3];           // boxes in the store at the end
        out_stats[3] = (int64_t)res[4];           // refinement rounds
        out_stats[4] = (int64_t)res[5];           // status (IB200_ST_*)
        out_stats[5] = 1;
    }
    return IB200_OK;
}
