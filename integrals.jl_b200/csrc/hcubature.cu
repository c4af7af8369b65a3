// hcubature.cu -- C entry point of the batched adaptive cubature (kernels: hcubature_impl.cuh)
#include "hcubature_impl.cuh"
#include "hcubature_pair.cuh"
#include "hcubature_rb.cuh"

namespace hc {
template <> int32_t launch<1>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<2>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<3>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<4>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<5>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<6>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<7>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<8>(ib200_ctx *, int, bool, Params &);
}
namespace hp {
template <> int32_t launch<2>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<3>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<4>(ib200_ctx *, int, bool, Params &);
template <> int32_t launch<5>(ib200_ctx *, int, bool, Params &);
}

namespace hb {
template <> int32_t launch<2>(ib200_ctx *, int, bool, Params &, long long);
template <> int32_t launch<3>(ib200_ctx *, int, bool, Params &, long long);
template <> int32_t launch<4>(ib200_ctx *, int, bool, Params &, long long);
template <> int32_t launch<5>(ib200_ctx *, int, bool, Params &, long long);
template <> int32_t launch<6>(ib200_ctx *, int, bool, Params &, long long);
template <> int32_t launch<7>(ib200_ctx *, int, bool, Params &, long long);
template <> int32_t launch<8>(ib200_ctx *, int, bool, Params &, long long);
}

namespace hp {
__global__ void lpt_hist_kernel(const double *I, const double *E, const int32_t *st, int64_t n, double rtol, double atol, unsigned *hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (st[i] == IB200_ST_MAXEVALS) atomicAdd(&hist[lpt_bucket(I[i], E[i], rtol, atol)], 1u);
}
__global__ void lpt_scan_kernel(unsigned *hist, unsigned *cursor, unsigned long long *norder, unsigned long long *counter) {
    __shared__ unsigned s[kLptBuckets];
    for (int b = threadIdx.x; b < kLptBuckets; b += blockDim.x) s[b] = hist[b];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int b = 0; b < kLptBuckets; b++) { const unsigned c = s[b]; s[b] = run; run += c; }
        *norder = run; *counter = 0ULL;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kLptBuckets; b += blockDim.x) cursor[b] = s[b];
}
__global__ void lpt_scatter_kernel(const double *I, const double *E, const int32_t *st, int64_t n, double rtol, double atol,
                                   unsigned *cursor, unsigned *order) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (st[i] == IB200_ST_MAXEVALS) order[atomicAdd(&cursor[lpt_bucket(I[i], E[i], rtol, atol)], 1u)] = (unsigned)i;
}
}  // namespace hp

extern "C" int32_t ib200_hcubature_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                         const double *lb, const double *ub, int32_t bpp,
                                         const double *params, int32_t nparam, int64_t nprob,
                                         double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t mode,
                                         double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_hcubature_batch: ctx is NULL");
    IB_REQUIRE(ctx, mode == IB200_HCUB_REFERENCE_ORDER || mode == IB200_HCUB_ROUNDS, "ib200_hcubature_batch: mode must be 0 (reference order) or 1 (rounds)");
    IB_REQUIRE(ctx, mode != IB200_HCUB_ROUNDS || dim >= 2, "ib200_hcubature_batch: the round-based mode is built for 2 <= dim <= 8 (use mode 0 in 1-D)");
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, "ib200_hcubature_batch: unknown integrand family");
    IB_REQUIRE(ctx, transform >= 0 && transform <= 2, "ib200_hcubature_batch: unknown transform id");
    IB_REQUIRE(ctx, ib::fam_dim_ok(family, dim), "ib200_hcubature_batch: dimension not supported for this family");
    IB_REQUIRE(ctx, nparam == ib::fam_nparam(family, dim), "ib200_hcubature_batch: nparam does not match the family");
    IB_REQUIRE(ctx, !(reltol < 0) && !(abstol < 0), "invalid negative tolerance");        // HCubature ArgumentError
    IB_REQUIRE(ctx, maxevals >= 0, "invalid negative maxevals");
    IB_REQUIRE(ctx, initdiv >= 1, "initdiv must be positive");
    IB_REQUIRE(ctx, nprob >= 0, "ib200_hcubature_batch: negative nprob");
    if (nprob == 0) return IB200_OK;
    IB_REQUIRE(ctx, lb && ub && params && out_I, "ib200_hcubature_batch: NULL data pointer");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t nb = (size_t)dim * (bpp ? (size_t)nprob : 1);
    const bool allfin = ib200_all_finite(ctx, lb, ub, nb);
    DevIn<double> dlb, dub, dpar;
    DevOut<double> oI, oE; DevOut<int64_t> oN; DevOut<int32_t> oS;
    int32_t rc;
    if ((rc = dlb.init(ctx, lb, nb, SL_IN0))) return rc;
    if ((rc = dub.init(ctx, ub, nb, SL_IN1))) return rc;
    if ((rc = dpar.init(ctx, params, (size_t)nparam * nprob, SL_IN2))) return rc;
    if ((rc = oI.init(ctx, out_I, (size_t)nprob, SL_OUT0))) return rc;
    if ((rc = oE.init(ctx, out_E, (size_t)nprob, SL_OUT1))) return rc;
    if ((rc = oN.init(ctx, out_nevals, (size_t)nprob, SL_OUT2))) return rc;
    if ((rc = oS.init(ctx, out_status, (size_t)nprob, SL_OUT3))) return rc;

    // box capacity per in-flight problem: initial grid + one box per refinement allowed by maxevals
    const int64_t np = hc::gm_npoints(dim);
    double init_boxes_d = 1.0;
    for (int k = 0; k < dim; k++) init_boxes_d *= initdiv;
    IB_REQUIRE(ctx, init_boxes_d <= 65536.0, "ib200_hcubature_batch: initdiv^dim exceeds the box capacity");
    const int64_t init_boxes = (int64_t)init_boxes_d;
    int64_t cap_max = ctx->opt_hcub_capacity > 0 ? ctx->opt_hcub_capacity : 65536;
    int64_t need = cap_max;
    if (maxevals < ((int64_t)1 << 40)) {
        const int64_t rem = maxevals - init_boxes * np;
        need = init_boxes + (rem > 0 ? (rem + 2 * np - 1) / (2 * np) : 0) + 1;
    }
    int64_t cap = need < cap_max ? need : cap_max;
    if (cap < init_boxes + 1) cap = init_boxes + 1;
    cap = (cap + hc::kGroup - 1) / hc::kGroup * hc::kGroup;

    unsigned long long *counter = (unsigned long long *)ctx->get(SL_WORK1, sizeof(unsigned long long));
    if (!counter) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: scratch allocation failed");
    IB_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));

    if (mode == IB200_HCUB_ROUNDS) {
        // round-based driver (hcubature_rb.cuh): budget in rule applications; a round bisects at least every box within a
        // factor two of the worst, so the last round may overshoot the budget -- the arena holds that slack
        hb::Params q;
        q.params = dpar.d; q.nparam = nparam; q.nprob = nprob; q.lb = dlb.d; q.ub = dub.d; q.bpp = bpp; q.tr = transform;
        q.rtol = reltol; q.atol = abstol; q.initdiv = initdiv; q.frac = ctx->opt_single_round_fraction;
        q.max_apps = maxevals < ((int64_t)1 << 60) ? (maxevals + np - 1) / np : ((int64_t)1 << 60);
        int64_t rcap = ctx->opt_hcub_capacity > 0 ? ctx->opt_hcub_capacity : 262144;
        if (q.max_apps < ((int64_t)1 << 40)) {
            const int64_t want = init_boxes + q.max_apps / 2 + q.max_apps / 8 + 1024;
            if (want < rcap) rcap = want;
        }
        if (rcap < init_boxes + 64) rcap = init_boxes + 64;
        q.cap = (int)((rcap + 127) / 128 * 128);
        q.rec = q.err = nullptr; q.list = nullptr; q.counter = counter;
        q.sub_rec = q.sub_err = nullptr; q.sub_out = nullptr; q.sub_tol = q.sub_wsum = q.sub_alpha = 0.0; q.sub_rank = 0; q.sub_world = 1;
        q.out_I = oI.d; q.out_E = oE.d; q.out_nevals = oN.d; q.out_status = oS.d;
        ctx->err.clear();
        ib200_time_begin(ctx);
        switch (dim) {
        case 2: rc = hb::launch<2>(ctx, family, allfin, q, 0); break;
        case 3: rc = hb::launch<3>(ctx, family, allfin, q, 0); break;
        case 4: rc = hb::launch<4>(ctx, family, allfin, q, 0); break;
        case 5: rc = hb::launch<5>(ctx, family, allfin, q, 0); break;
        case 6: rc = hb::launch<6>(ctx, family, allfin, q, 0); break;
        case 7: rc = hb::launch<7>(ctx, family, allfin, q, 0); break;
        default: rc = hb::launch<8>(ctx, family, allfin, q, 0); break;
        }
        if (rc) return rc;
        ib200_time_end(ctx);
        if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
        if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return IB200_OK;
    }

    hc::Params p;
    p.params = dpar.d; p.nparam = nparam; p.nprob = nprob; p.lb = dlb.d; p.ub = dub.d; p.bpp = bpp; p.tr = transform;
    p.rtol = reltol; p.atol = abstol; p.maxevals = maxevals; p.initdiv = initdiv; p.cap = (int)cap;
    p.keys = nullptr; p.recs = nullptr; p.counter = counter;
    p.out_I = oI.d; p.out_E = oE.d; p.out_nevals = oN.d; p.out_status = oS.d;

    ctx->err.clear();
    // pair mode (thread per box, hcubature_pair.cuh) for 2 <= dim <= 5 when the batch can fill the GPU's pair slots;
    // warp-per-problem mode (hcubature_impl.cuh) for 1-D, dim >= 6 and small batches
    const bool pair_ok = dim >= 2 && dim <= 5 && cap <= 65536 - 1023;
    const bool use_pair = pair_ok && (ctx->opt_hcub_mode == 2 || (ctx->opt_hcub_mode == 0 && nprob >= 4096));
    if (use_pair) {
        hp::Params q;
        q.params = dpar.d; q.nparam = nparam; q.nprob = nprob; q.lb = dlb.d; q.ub = dub.d; q.bpp = bpp; q.tr = transform;
        q.rtol = reltol; q.atol = abstol; q.maxevals = maxevals; q.initdiv = initdiv;
        q.cap = (int)((cap + 1023) / 1024 * 1024);
        q.keys = q.l1 = q.rec = nullptr; q.l1off = nullptr; q.counter = counter;
        q.order = nullptr; q.norder = nullptr; q.gk_order = 0; q.gk_x = q.gk_w = q.gk_wg = nullptr;
        q.out_I = oI.d; q.out_E = oE.d; q.out_nevals = oN.d; q.out_status = oS.d;
        auto run = [&](hp::Params &pp) -> int32_t {
            switch (dim) {
            case 2: return hp::launch<2>(ctx, family, allfin, pp);
            case 3: return hp::launch<3>(ctx, family, allfin, pp);
            case 4: return hp::launch<4>(ctx, family, allfin, pp);
            default: return hp::launch<5>(ctx, family, allfin, pp);
            }
        };
        ib200_time_begin(ctx);
        // longest-expected-first: pilot pass + counting sort (hcubature_pair.cuh); needs I, E and status of every problem
        const int64_t pilot_evals = init_boxes * np + 2 * np * 24;
        const bool lpt = ctx->opt_hcub_lpt != 0 && nprob >= 8192 && nprob < ((int64_t)1 << 32) && maxevals > 16 * pilot_evals &&
                         oE.d && oS.d;
        if (lpt) {
            hp::Params pilot = q;
            pilot.maxevals = pilot_evals;
            pilot.cap = (int)((init_boxes + 24 + 1 + 1023) / 1024 * 1024);
            if ((rc = run(pilot))) return rc;
            unsigned *lw = (unsigned *)ctx->get(SL_WORK3, ((size_t)nprob + 2 * hp::kLptBuckets + 4) * sizeof(unsigned));
            if (!lw) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: scratch allocation failed");
            unsigned *hist = lw, *cursor = lw + hp::kLptBuckets;
            unsigned long long *norder = reinterpret_cast<unsigned long long *>(lw + 2 * hp::kLptBuckets);
            unsigned *order = lw + 2 * hp::kLptBuckets + 4;
            IB_CUDA(ctx, cudaMemsetAsync(hist, 0, hp::kLptBuckets * sizeof(unsigned), ctx->stream));
            const int gb = ctx->sm_count * 4;
            hp::lpt_hist_kernel<<<gb, 256, 0, ctx->stream>>>(oI.d, oE.d, oS.d, nprob, reltol, abstol, hist);
            hp::lpt_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, cursor, norder, counter);
            hp::lpt_scatter_kernel<<<gb, 256, 0, ctx->stream>>>(oI.d, oE.d, oS.d, nprob, reltol, abstol, cursor, order);
            ctx->launches += 3;
            IB_LAUNCH_CHECK(ctx);
            q.order = order; q.norder = norder;
        }
        rc = run(q);
        if (rc) return rc;
        ib200_time_end(ctx);
        if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
        if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return IB200_OK;
    }
    ib200_time_begin(ctx);
    switch (dim) {
    case 1: rc = hc::launch<1>(ctx, family, allfin, p); break;
    case 2: rc = hc::launch<2>(ctx, family, allfin, p); break;
    case 3: rc = hc::launch<3>(ctx, family, allfin, p); break;
    case 4: rc = hc::launch<4>(ctx, family, allfin, p); break;
    case 5: rc = hc::launch<5>(ctx, family, allfin, p); break;
    case 6: rc = hc::launch<6>(ctx, family, allfin, p); break;
    case 7: rc = hc::launch<7>(ctx, family, allfin, p); break;
    case 8: rc = hc::launch<8>(ctx, family, allfin, p); break;
    default: rc = ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_batch: dim must be 1..8");
    }
    if (rc) return rc;
    ib200_time_end(ctx);
    if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
    if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}
