// hcubature_vec.cu -- C entry point of the batched adaptive cubature of vector-valued integrands (kernel: hcubature_vec.cuh)
#include "hcubature_vec.cuh"

namespace hv {
template <> int32_t launch<1>(ib200_ctx *, int, Params &);
template <> int32_t launch<2>(ib200_ctx *, int, Params &);
template <> int32_t launch<3>(ib200_ctx *, int, Params &);
template <> int32_t launch<4>(ib200_ctx *, int, Params &);
template <> int32_t launch<5>(ib200_ctx *, int, Params &);
template <> int32_t launch<6>(ib200_ctx *, int, Params &);
template <> int32_t launch<7>(ib200_ctx *, int, Params &);
template <> int32_t launch<8>(ib200_ctx *, int, Params &);
}

static int32_t hcubature_vec_impl(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                  const double *lb, const double *ub, int32_t bpp,
                                  const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                                  double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t normkind, const ib::SensIdx sens,
                                  double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

extern "C" int32_t ib200_hcubature_vec_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                             const double *lb, const double *ub, int32_t bpp,
                                             const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                                             double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t normkind,
                                             double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_hcubature_vec_batch: ctx is NULL");
    IB_REQUIRE(ctx, normkind == 0 || normkind == 1, "ib200_hcubature_vec_batch: norm must be 0 (2-norm) or 1 (maximum norm)");
    ib::SensIdx none; none.n = 0;
    return hcubature_vec_impl(ctx, family, dim, transform, lb, ub, bpp, params, nparam, ncomp, nprob, reltol, abstol, maxevals, initdiv, normkind,
                              none, out_I, out_E, out_nevals, out_status);
}

// I and dI/dp as one vector-valued integral [f, df/dp_s...] (see ib::Grad, family.cuh)
extern "C" int32_t ib200_hcubature_sens_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                              const double *lb, const double *ub, int32_t bpp,
                                              const double *params, int32_t nparam, int64_t nprob,
                                              const int32_t *sens_idx, int32_t nsens,
                                              double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t normkind,
                                              double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_hcubature_sens_batch: ctx is NULL");
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, "ib200_hcubature_sens_batch: unknown integrand family");
    IB_REQUIRE(ctx, nsens >= 1 && nsens <= 7 && sens_idx != nullptr, "ib200_hcubature_sens_batch: 1 <= nsens <= 7 parameters to differentiate with respect to");
    IB_REQUIRE(ctx, normkind >= 0 && normkind <= 2, "ib200_hcubature_sens_batch: norm must be 0 (2-norm), 1 (maximum norm) or 2 (value only)");
    IB_REQUIRE(ctx, family != IB200_FAM_GENZ_DISC, "ib200_hcubature_sens_batch: the discontinuous family has no parameter sensitivities (its support moves with u)");
    ib::SensIdx si; si.n = nsens;
    for (int s = 0; s < 7; s++) {
        si.idx[s] = s < nsens ? sens_idx[s] : 0;
        IB_REQUIRE(ctx, si.idx[s] >= 0 && si.idx[s] < nparam, "ib200_hcubature_sens_batch: parameter index out of range");
    }
    return hcubature_vec_impl(ctx, family, dim, transform, lb, ub, bpp, params, nparam, 1 + nsens, nprob, reltol, abstol, maxevals, initdiv, normkind,
                              si, out_I, out_E, out_nevals, out_status);
}

static int32_t hcubature_vec_impl(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                  const double *lb, const double *ub, int32_t bpp,
                                  const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                                  double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t normkind, const ib::SensIdx sens,
                                  double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_hcubature_vec_batch: ctx is NULL");
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, "ib200_hcubature_vec_batch: unknown integrand family");
    IB_REQUIRE(ctx, transform >= 0 && transform <= 2, "ib200_hcubature_vec_batch: unknown transform id");
    IB_REQUIRE(ctx, ib::fam_dim_ok(family, dim), "ib200_hcubature_vec_batch: dimension not supported for this family");
    IB_REQUIRE(ctx, nparam == ib::fam_nparam(family, dim), "ib200_hcubature_vec_batch: nparam does not match the family");
    IB_REQUIRE(ctx, ncomp >= 1 && ncomp <= hv::NC, "ib200_hcubature_vec_batch: 1 <= ncomp <= 8 components");
    IB_REQUIRE(ctx, !(reltol < 0) && !(abstol < 0), "invalid negative tolerance");        // HCubature ArgumentError
    IB_REQUIRE(ctx, maxevals >= 0, "invalid negative maxevals");
    IB_REQUIRE(ctx, initdiv >= 1, "initdiv must be positive");
    IB_REQUIRE(ctx, nprob >= 0, "ib200_hcubature_vec_batch: negative nprob");
    if (nprob == 0) return IB200_OK;
    IB_REQUIRE(ctx, lb && ub && params && out_I, "ib200_hcubature_vec_batch: NULL data pointer");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t nb = (size_t)dim * (bpp ? (size_t)nprob : 1);
    DevIn<double> dlb, dub, dpar;
    DevOut<double> oI, oE; DevOut<int64_t> oN; DevOut<int32_t> oS;
    int32_t rc;
    if ((rc = dlb.init(ctx, lb, nb, SL_IN0))) return rc;
    if ((rc = dub.init(ctx, ub, nb, SL_IN1))) return rc;
    if ((rc = dpar.init(ctx, params, (size_t)nparam * (sens.n > 0 ? 1 : ncomp) * nprob, SL_IN2))) return rc;
    if ((rc = oI.init(ctx, out_I, (size_t)ncomp * nprob, SL_OUT0))) return rc;
    if ((rc = oE.init(ctx, out_E, (size_t)nprob, SL_OUT1))) return rc;
    if ((rc = oN.init(ctx, out_nevals, (size_t)nprob, SL_OUT2))) return rc;
    if ((rc = oS.init(ctx, out_status, (size_t)nprob, SL_OUT3))) return rc;

    // box capacity per in-flight problem: initial grid + one box per refinement allowed by maxevals (as ib200_hcubature_batch)
    const int64_t np = hc::gm_npoints(dim);
    double init_boxes_d = 1.0;
    for (int k = 0; k < dim; k++) init_boxes_d *= initdiv;
    IB_REQUIRE(ctx, init_boxes_d <= 65536.0, "ib200_hcubature_vec_batch: initdiv^dim exceeds the box capacity");
    const int64_t init_boxes = (int64_t)init_boxes_d;
    const int64_t cap_max = ctx->opt_hcub_capacity > 0 ? ctx->opt_hcub_capacity : 65536;
    int64_t need = cap_max;
    if (maxevals < ((int64_t)1 << 40)) {
        const int64_t rem = maxevals - init_boxes * np;
        need = init_boxes + (rem > 0 ? (rem + 2 * np - 1) / (2 * np) : 0) + 1;
    }
    int64_t cap = need < cap_max ? need : cap_max;
    if (cap < init_boxes + 1) cap = init_boxes + 1;
    cap = (cap + hv::kGroup - 1) / hv::kGroup * hv::kGroup;

    unsigned long long *counter = (unsigned long long *)ctx->get(SL_WORK1, sizeof(unsigned long long));
    if (!counter) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_vec_batch: scratch allocation failed");
    IB_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));

    hv::Params p;
    p.params = dpar.d; p.nparam = nparam; p.ncomp = ncomp; p.nprob = nprob; p.lb = dlb.d; p.ub = dub.d; p.bpp = bpp; p.tr = transform;
    p.rtol = reltol; p.atol = abstol; p.maxevals = maxevals; p.initdiv = initdiv; p.normkind = normkind; p.sens = sens; p.cap = (int)cap;
    p.keys = p.ivals = p.recs = nullptr; p.counter = counter;
    p.out_I = oI.d; p.out_E = oE.d; p.out_nevals = oN.d; p.out_status = oS.d;

    ctx->err.clear();
    ib200_time_begin(ctx);
    switch (dim) {
    case 1: rc = hv::launch<1>(ctx, family, p); break;
    case 2: rc = hv::launch<2>(ctx, family, p); break;
    case 3: rc = hv::launch<3>(ctx, family, p); break;
    case 4: rc = hv::launch<4>(ctx, family, p); break;
    case 5: rc = hv::launch<5>(ctx, family, p); break;
    case 6: rc = hv::launch<6>(ctx, family, p); break;
    case 7: rc = hv::launch<7>(ctx, family, p); break;
    case 8: rc = hv::launch<8>(ctx, family, p); break;
    default: rc = ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_vec_batch: dim must be 1..8");
    }
    if (rc) return rc;
    ib200_time_end(ctx);
    if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
    if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}
