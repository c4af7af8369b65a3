// fixed_rule.cu -- batched fixed quadrature rules on the device.
//
// Replaces evalrule(f, p, lb, ub, nodes, weights) (src/quadrules.jl:1-19), as reached through the
// automatic ChangeOfVariables (src/common.jl:97-103, src/Integrals.jl:192-224), and the 1-D
// gauss_legendre / composite_gauss_legendre loops (ext/IntegralsFastGaussQuadratureExt.jl:11-32):
//     I = prod(scale) * sum_i w_i * g(scale .* x_i + shift),   g(t) = f(u(t), p) * prod_k du_k/dt_k
// where (scale, shift) map the rule's [-1,1]^d onto the transformed domain (tlo, thi).
//
// tensor_kernel  one CTA per problem.  The rule is the d-fold tensor product of a 1-D rule; the
//                per-axis pieces of the integrand (family "separable form", family.cuh) and the
//                weight*jacobian products are tabulated once per problem in shared memory, so the
//                innermost loop is: 1 LDS.128, one add/mul, the family's transcendental, one DFMA.
// nodes_kernel   explicit N x d node list (any rule); a warp or a CTA per problem depending on N.
// Both reduce in a fixed order (deterministic run to run).
#include "common.cuh"
#include "family.cuh"

namespace {

constexpr int kBlock = 256;
#ifndef IB200_TENSOR_ROWS
#define IB200_TENSOR_ROWS 8      // items per thread and pass in the angle-addition loop of tensor_kernel (measured, see there)
#endif

__device__ __forceinline__ double block_reduce_sum(double v, double *red /* >= kBlock/32 doubles */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        s = red[0];
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) s += red[k];
    }
    return s;   // valid in thread 0
}

template <int FAM, int D>
__global__ void __launch_bounds__(kBlock)
tensor_kernel(const double *__restrict__ params, int nparam, int64_t nprob,
              const double *__restrict__ lb, const double *__restrict__ ub, int bpp, int tr,
              const double *__restrict__ x1d, const double *__restrict__ w1d, int n, double *__restrict__ out) {
    using F = ib::Family<FAM, D>;
    extern __shared__ double sm[];
    // layout: tab[D][n] as double2 {tau, w*jac}; par[nparam]; red[8]; pscale
    double2 *tab = reinterpret_cast<double2 *>(sm);
    double *par = sm + 2 * (size_t)D * n;
    double *red = par + nparam;
    double *pscale = red + kBlock / 32;
    // oscillatory family only: {cos, sin} of the innermost axis' phase term, for the angle-addition inner loop
    double2 *cs = reinterpret_cast<double2 *>(pscale + IB200_MAX_DIM + ((nparam + kBlock / 32 + IB200_MAX_DIM) & 1));
    constexpr bool kAngle = (FAM == IB200_FAM_GENZ_OSC);

    int64_t outer = 1;
#pragma unroll
    for (int k = 0; k < D - 1; k++) outer *= n;
    int isplit = 1;
    while (outer * isplit < kBlock && isplit < n) isplit <<= 1;
    const int chunk = (n + isplit - 1) / isplit;
    const int64_t items = outer * isplit;

    for (int64_t prob = blockIdx.x; prob < nprob; prob += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < nparam; i += kBlock) par[i] = params[prob * nparam + i];
        __syncthreads();
        const double *lbp = bpp ? lb + prob * D : lb;
        const double *ubp = bpp ? ub + prob * D : ub;
        for (int idx = threadIdx.x; idx < D * n; idx += kBlock) {
            const int k = idx / n, j = idx - k * n;
            ib::AxisMap m;
            ib::make_axis(lbp[k], ubp[k], m);
            const double sc = (m.thi - m.tlo) / 2, sh = (m.tlo + m.thi) / 2;   // quadrules.jl:2-3
            const double t = sc * x1d[j] + sh;
            double u, jac;
            ib::t2ujac(tr, t, m, u, jac);
            const double term = F::sep_term(k, u, par);
            tab[idx] = make_double2(term, w1d[j] * jac);
            if (kAngle && k == D - 1) { double sn, cn; sincos(term, &sn, &cn); cs[j] = make_double2(cn, sn); }
            if (j == 0) pscale[k] = sc;
        }
        __syncthreads();
        const double init = F::sep_init(par);
        const double2 *tin = tab + (size_t)(D - 1) * n;
        double total = 0.0;
        if constexpr (kAngle) {
            // Oscillatory family: cos(phi + theta_j) = cos(phi) cos(theta_j) - sin(phi) sin(theta_j): phi (the outer axes' phase) costs one
            // sincos per item, every node of the innermost axis one multiply and two FMAs (angle addition from per-axis
            // tables; the integrand is still formed at every node of the rule).  A thread takes kRows items at a time through the same
            // inner loop, so the table entries of a node ({cos, sin} and w * jac: two shared-memory loads) serve kRows nodes -- ncu on
            // the one-item version: LSU 54.7 % against FP64 45.6 %, the loads, not the arithmetic, set the time
            // (profiles/r01b_tensor_c3_b_full_summary.txt; with 4 items: LSU 17 %, FP64 pipe 53.7 %, profiles/r02_tensor_c3_rows4_full_summary.txt).
            // isplit is a power of two <= kBlock, so the items of a thread share their chunk.
            // Measured on BASELINE config 3 (65,536 integrals): 1 item per pass 7.36 ms, 2 or 4 items 5.95 ms, 8 items 5.67 ms.  Full groups of
            // kRows items first, single items for what is left (small rules have fewer items than kRows per thread).
            auto rows = [&](auto rc, int64_t it0) {
                constexpr int R = decltype(rc)::value;
                double qs[R], qc[R], wo[R], s0[R];
                const int c = (int)(it0 % isplit);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    int64_t o = (it0 + (int64_t)r * kBlock) / isplit;
                    double acc = init, w = 1.0;
#pragma unroll
                    for (int k = 0; k < D - 1; k++) {
                        const int ik = (int)(o % n); o /= n;
                        const double2 e = tab[k * n + ik];
                        acc = acc + e.x;
                        w *= e.y;
                    }
                    sincos(acc, &qs[r], &qc[r]);
                    wo[r] = w;
                    s0[r] = 0.0;
                }
                const int k0 = c * chunk, k1 = min(n, k0 + chunk);
                for (int k = k0; k < k1; k++) {
                    const double2 cj = cs[k];
                    const double wj = tin[k].y;
#pragma unroll
                    for (int r = 0; r < R; r++) s0[r] = fma(wj, fma(qc[r], cj.x, -(qs[r] * cj.y)), s0[r]);
                }
#pragma unroll
                for (int r = 0; r < R; r++) total = fma(wo[r], s0[r], total);
            };
            constexpr int kRows = IB200_TENSOR_ROWS;
            int64_t it0 = threadIdx.x;
            for (; it0 + (int64_t)(kRows - 1) * kBlock < items; it0 += (int64_t)kRows * kBlock) rows(std::integral_constant<int, kRows>{}, it0);
            for (; it0 < items; it0 += kBlock) rows(std::integral_constant<int, 1>{}, it0);
        } else
        for (int64_t it = threadIdx.x; it < items; it += kBlock) {
            int64_t o = it / isplit;
            const int c = (int)(it - o * isplit);
            double acc = init, wo = 1.0;
#pragma unroll
            for (int k = 0; k < D - 1; k++) {
                const int ik = (int)(o % n); o /= n;
                const double2 e = tab[k * n + ik];
                acc = F::kProd ? acc * e.x : acc + e.x;
                wo *= e.y;
            }
            const int k0 = c * chunk, k1 = min(n, k0 + chunk);
            double s0 = 0.0, s1 = 0.0;
            int k = k0;
            for (; k + 1 < k1; k += 2) {
                const double2 e0 = tin[k], e1 = tin[k + 1];
                const double a0 = F::kProd ? acc * e0.x : acc + e0.x;
                const double a1 = F::kProd ? acc * e1.x : acc + e1.x;
                s0 = fma(e0.y, F::sep_finish(a0, par), s0);
                s1 = fma(e1.y, F::sep_finish(a1, par), s1);
            }
            if (k < k1) {
                const double2 e0 = tin[k];
                const double a0 = F::kProd ? acc * e0.x : acc + e0.x;
                s0 = fma(e0.y, F::sep_finish(a0, par), s0);
            }
            total = fma(wo, s0 + s1, total);
        }
        const double s = block_reduce_sum(total, red);
        if (threadIdx.x == 0) {
            double ps = pscale[0];
#pragma unroll
            for (int k = 1; k < D; k++) ps *= pscale[k];
            out[prob] = ps * s;
        }
    }
}

// explicit nodes; G threads per problem (G = 32: one warp; G = kBlock: one CTA)
template <int FAM, int D, bool ALLFIN, int G>
__global__ void __launch_bounds__(kBlock)
nodes_kernel(const double *__restrict__ params, int nparam, int64_t nprob,
             const double *__restrict__ lb, const double *__restrict__ ub, int bpp, int tr,
             const double *__restrict__ nodes, const double *__restrict__ weights, int64_t N,
             double *__restrict__ out) {
    __shared__ double red[kBlock / 32];
    constexpr int groups = kBlock / G;
    const int g = threadIdx.x / G, lane = threadIdx.x % G;
    for (int64_t base = (int64_t)blockIdx.x * groups; base < nprob; base += (int64_t)gridDim.x * groups) {
        const int64_t prob = base + g;
        double total = 0.0, ps = 1.0;
        if (prob < nprob) {
            ib::Problem<FAM, D, ALLFIN> P;
            P.load(params + prob * nparam, bpp ? lb + prob * D : lb, bpp ? ub + prob * D : ub, tr);
            double sc[D], sh[D];
#pragma unroll
            for (int k = 0; k < D; k++) { sc[k] = P.tscale(k); sh[k] = P.tshift(k); ps = (k == 0) ? sc[k] : ps * sc[k]; }
            for (int64_t i = lane; i < N; i += G) {
                double t[D];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const double xn = __ldg(nodes + i * D + k);
                    t[k] = ALLFIN ? xn : sc[k] * xn + sh[k];
                }
                total = total + __ldg(weights + i) * P(t);
            }
        }
        if (G == 32) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) total += __shfl_down_sync(0xffffffffu, total, off);
            if (lane == 0 && prob < nprob) out[prob] = ps * total;
        } else {
            const double s = block_reduce_sum(total, red);
            if (threadIdx.x == 0 && prob < nprob) out[prob] = ps * s;
        }
    }
}

struct BatchArgs {
    DevIn<double> lb, ub, params;
    DevOut<double> out;
    bool allfin = false;
};

int32_t stage_batch(ib200_ctx *ctx, const char *who, int32_t family, int32_t dim, int32_t transform,
                    const double *lb, const double *ub, int32_t bpp, const double *params, int32_t nparam,
                    int64_t nprob, double *out_I, BatchArgs &a) {
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, std::string(who) + ": unknown integrand family");
    IB_REQUIRE(ctx, ib::fam_dim_ok(family, dim), std::string(who) + ": dimension not supported for this family");
    IB_REQUIRE(ctx, transform >= 0 && transform <= 2, std::string(who) + ": unknown transform id");
    IB_REQUIRE(ctx, nparam == ib::fam_nparam(family, dim), std::string(who) + ": nparam does not match the family");
    IB_REQUIRE(ctx, nprob >= 0, std::string(who) + ": negative nprob");
    if (nprob == 0) return IB200_OK;
    IB_REQUIRE(ctx, lb && ub && params && out_I, std::string(who) + ": NULL data pointer");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nb = (size_t)dim * (bpp ? (size_t)nprob : 1);
    a.allfin = ib200_all_finite(ctx, lb, ub, nb);
    int32_t rc;
    if ((rc = a.lb.init(ctx, lb, nb, SL_IN0))) return rc;
    if ((rc = a.ub.init(ctx, ub, nb, SL_IN1))) return rc;
    if ((rc = a.params.init(ctx, params, (size_t)nparam * nprob, SL_IN2))) return rc;
    if ((rc = a.out.init(ctx, out_I, (size_t)nprob, SL_OUT0))) return rc;
    return IB200_OK;
}

int32_t finish_batch(ib200_ctx *ctx, BatchArgs &a) {
    int32_t rc;
    if ((rc = a.out.finish(ctx))) return rc;
    if (a.out.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}

int32_t launch_nodes(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform, int32_t bpp,
                     int32_t nparam, int64_t nprob, const BatchArgs &a, const double *dnodes,
                     const double *dweights, int64_t N) {
    bool launched = false;
    const bool big = N >= 2048;
    const int64_t groups = big ? 1 : kBlock / 32;
    int64_t blocks = (nprob + groups - 1) / groups;
    const int64_t maxb = (int64_t)ctx->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        ib200_dispatch_dim<IB200_MAX_DIM>(dim, [&](auto dc) {
            constexpr int D = decltype(dc)::value;
            if constexpr (ib::fam_dim_ok(FAM, D)) {
                auto go = [&](auto kern) {
                    kern<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(a.params.d, nparam, nprob, a.lb.d, a.ub.d, bpp,
                                                                      transform, dnodes, dweights, N, a.out.d);
                    launched = true;
                };
                if (a.allfin) { if (big) go(nodes_kernel<FAM, D, true, kBlock>); else go(nodes_kernel<FAM, D, true, 32>); }
                else          { if (big) go(nodes_kernel<FAM, D, false, kBlock>); else go(nodes_kernel<FAM, D, false, 32>); }
            }
        });
    });
    if (!launched) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "fixed rule: family/dimension not compiled in");
    IB_LAUNCH_CHECK(ctx);
    return IB200_OK;
}

}  // namespace

extern "C" int32_t ib200_fixed_rule_tensor_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                                 const double *lb, const double *ub, int32_t bpp,
                                                 const double *params, int32_t nparam, int64_t nprob,
                                                 const double *x1d, const double *w1d, int32_t n1d, double *out_I) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_fixed_rule_tensor_batch: ctx is NULL");
    IB_REQUIRE(ctx, n1d >= 1 && x1d && w1d, "empty quadrature rule");                 // quadrules.jl:9
    IB_REQUIRE(ctx, dim >= 1 && dim <= 6, "ib200_fixed_rule_tensor_batch: tensor rules support 1 <= dim <= 6");
    BatchArgs a;
    int32_t rc = stage_batch(ctx, "ib200_fixed_rule_tensor_batch", family, dim, transform, lb, ub, bpp, params, nparam, nprob, out_I, a);
    if (rc || nprob == 0) return rc;
    DevIn<double> dx, dw;
    if ((rc = dx.init(ctx, x1d, (size_t)n1d, SL_IN3))) return rc;
    if ((rc = dw.init(ctx, w1d, (size_t)n1d, SL_IN4))) return rc;
    const size_t smem = (2 * (size_t)dim * n1d + nparam + kBlock / 32 + IB200_MAX_DIM + 1 + 2 * (size_t)n1d) * sizeof(double);
    IB_REQUIRE(ctx, smem <= ctx->smem_optin, "ib200_fixed_rule_tensor_batch: 1-D rule too long for shared memory");
    int64_t blocks = nprob;
    const int64_t maxb = (int64_t)ctx->sm_count * 4;
    if (blocks > maxb) blocks = maxb;
    bool launched = false;
    cudaError_t attr_err = cudaSuccess;
    ib200_time_begin(ctx);
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        ib200_dispatch_dim<6>(dim, [&](auto dc) {
            constexpr int D = decltype(dc)::value;
            if constexpr (ib::fam_dim_ok(FAM, D)) {
                auto kern = tensor_kernel<FAM, D>;
                if (smem > 48 * 1024) attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(a.params.d, nparam, nprob, a.lb.d, a.ub.d, bpp, transform,
                                                                     dx.d, dw.d, n1d, a.out.d);
                launched = true;
            }
        });
    });
    if (attr_err != cudaSuccess) return ib200_fail(ctx, IB200_ERR_CUDA, "cudaFuncSetAttribute failed");
    if (!launched) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_fixed_rule_tensor_batch: family/dimension not compiled in");
    IB_LAUNCH_CHECK(ctx);
    ib200_time_end(ctx);
    return finish_batch(ctx, a);
}

extern "C" int32_t ib200_fixed_rule_nodes_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                                const double *lb, const double *ub, int32_t bpp,
                                                const double *params, int32_t nparam, int64_t nprob,
                                                const double *nodes, const double *weights, int64_t N, double *out_I) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_fixed_rule_nodes_batch: ctx is NULL");
    IB_REQUIRE(ctx, N >= 1 && nodes && weights, "empty quadrature rule");             // quadrules.jl:9
    BatchArgs a;
    int32_t rc = stage_batch(ctx, "ib200_fixed_rule_nodes_batch", family, dim, transform, lb, ub, bpp, params, nparam, nprob, out_I, a);
    if (rc || nprob == 0) return rc;
    DevIn<double> dn, dw;
    if ((rc = dn.init(ctx, nodes, (size_t)N * dim, SL_IN3))) return rc;
    if ((rc = dw.init(ctx, weights, (size_t)N, SL_IN4))) return rc;
    ib200_time_begin(ctx);
    if ((rc = launch_nodes(ctx, family, dim, transform, bpp, nparam, nprob, a, dn.d, dw.d, N))) return rc;
    ib200_time_end(ctx);
    return finish_batch(ctx, a);
}

// 1-D (composite) Gauss-Legendre: the sub-interval nodes scale_i*x_j + shift_i and weights
// w_j*scale_i of ext/IntegralsFastGaussQuadratureExt.jl:11-32 are expanded on the host (n*sub
// doubles) and evaluated by nodes_kernel on the transformed domain.
extern "C" int32_t ib200_gauss_legendre_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                              const double *lb, const double *ub, int32_t bpp,
                                              const double *params, int32_t nparam, int64_t nprob,
                                              const double *nodes, const double *weights, int32_t n,
                                              int32_t subintervals, double *out_I) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_gauss_legendre_batch: ctx is NULL");
    IB_REQUIRE(ctx, n >= 1 && nodes && weights, "ib200_gauss_legendre_batch: empty rule");
    IB_REQUIRE(ctx, subintervals >= 1, "ib200_gauss_legendre_batch: subintervals must be positive");
    IB_REQUIRE(ctx, !ib200_is_device_ptr(nodes) && !ib200_is_device_ptr(weights),
               "ib200_gauss_legendre_batch: nodes/weights must be host arrays (they are set at algorithm construction, src/algorithms.jl:244-249)");
    BatchArgs a;
    int32_t rc = stage_batch(ctx, "ib200_gauss_legendre_batch", family, 1, transform, lb, ub, bpp, params, nparam, nprob, out_I, a);
    if (rc || nprob == 0) return rc;
    // composite rule on the reference interval [-1,1]: nodes_kernel applies the (tlo,thi) affine map
    // and the variable change; h, _lb, _ub as in composite_gauss_legendre with lb=-1, ub=1.
    const int64_t N = (int64_t)n * subintervals;
    std::vector<double> xs((size_t)N), ws((size_t)N);
    if (subintervals == 1) {
        for (int j = 0; j < n; j++) { xs[j] = nodes[j]; ws[j] = weights[j]; }
    } else {
        const double h = 2.0 / subintervals;
        for (int i = 0; i < subintervals; i++) {
            const double _lb = (i == 0) ? -1.0 : -1.0 + i * h;
            const double _ub = _lb + h;
            const double scale = (_ub - _lb) / 2, shift = (_lb + _ub) / 2;
            for (int j = 0; j < n; j++) { xs[(size_t)i * n + j] = scale * nodes[j] + shift; ws[(size_t)i * n + j] = weights[j] * scale; }
        }
    }
    DevIn<double> dn, dw;
    if ((rc = dn.init(ctx, xs.data(), (size_t)N, SL_IN3))) return rc;
    if ((rc = dw.init(ctx, ws.data(), (size_t)N, SL_IN4))) return rc;
    IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // xs/ws are stack-owned: finish the H2D before they die
    ib200_time_begin(ctx);
    if ((rc = launch_nodes(ctx, family, 1, transform, bpp, nparam, nprob, a, dn.d, dw.d, N))) return rc;
    ib200_time_end(ctx);
    return finish_batch(ctx, a);
}
