/*
 * integrals_b200.h -- C ABI of libintegrals_b200.so (CUDA, sm_100a).
 *
 * This is the drop-in boundary for the Integrals.jl hot path: each entry point is what a
 * Julia `ccall` from the new algorithm types (QuadGKB200, HCubatureB200, QuadratureRuleB200 /
 * GaussLegendreB200, TrapezoidalB200 / SimpsonsB200) binds, and it replaces the reference
 * function cited beside it (paths relative to the Integrals.jl v5.4.4 tree).  INTEGRATION.md
 * shows the Julia-side binding.  Plain C types only; no callbacks into the caller.
 *
 * Conventions
 *  - Every call returns 0 on success or a negative IB200_ERR_* code; the message is available
 *    from ib200_last_error(ctx).  No C++ exception crosses the boundary.
 *  - Data pointers may be HOST or DEVICE pointers (resolved per pointer with
 *    cudaPointerGetAttributes).  Host inputs are staged to the device inside the call and host
 *    outputs are copied back before it returns.  When every output pointer is a device pointer
 *    the call is asynchronous on the context's stream (ib200_synchronize to wait).
 *  - The caller owns all buffers; the library keeps no caller pointer after a call returns.
 *  - `params` is nparam x nprob, column-major (Julia Matrix): problem k's parameters are the
 *    contiguous doubles params[k*nparam .. (k+1)*nparam).  See ib200_family_nparam.
 *  - `lb`,`ub` are `dim` doubles shared by all problems, or dim x nprob column-major when
 *    bounds_per_problem != 0.  +-INFINITY selects the infinity_handling.jl variable change.
 *  - `transform`: which ChangeOfVariables map the reference would apply (src/common.jl:97-103):
 *    0 transformation_if_inf, 1 transformation_tan_inf, 2 transformation_cot_inf
 *    (src/infinity_handling.jl:165-170, 277-282, 405-410).  Applied on the device.
 *  - One context is bound to one GPU and may be used by one host thread at a time
 *    (one process per GPU; multi-GPU jobs shard problems across processes).
 */
#ifndef INTEGRALS_B200_H
#define INTEGRALS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ib200_ctx ib200_ctx;

#define IB200_OK               0
#define IB200_ERR_INVALID     (-1)  /* bad argument (message says which)                     */
#define IB200_ERR_CUDA        (-2)  /* CUDA runtime error                                     */
#define IB200_ERR_OOM         (-3)  /* device memory exhausted                                */
#define IB200_ERR_UNSUPPORTED (-4)  /* family/dimension/transform combination not compiled in */
#define IB200_ERR_NCCL        (-5)  /* NCCL error (multi-GPU single-domain solve)             */

/* registered device-side integrand families; x in R^dim, per-problem params */
#define IB200_FAM_SINXP      0   /* [p]               sin(x*p)                         dim 1   */
#define IB200_FAM_GENZ_OSC   1   /* [a(d),u(d)]       cos(2 pi u1 + sum a_i x_i)               */
#define IB200_FAM_GENZ_PROD  2   /* [a,u]             1/prod(a_i^-2 + (x_i-u_i)^2)             */
#define IB200_FAM_GENZ_CORN  3   /* [a,u]             (1 + sum a_i x_i)^-(d+1)                 */
#define IB200_FAM_GENZ_GAUSS 4   /* [a,u]             exp(-sum a_i^2 (x_i-u_i)^2)              */
#define IB200_FAM_GENZ_C0    5   /* [a,u]             exp(-sum a_i |x_i-u_i|)                  */
#define IB200_FAM_GENZ_DISC  6   /* [a,u]             0 if x1>u1 or x2>u2 else exp(sum a_i x_i) */
#define IB200_FAM_COSPROD    7   /* [p(d)]            prod cos(p_i x_i)                        */
#define IB200_FAM_GAUSSPOLY  8   /* [a,u,k(d),c]      c prod x_i^k_i exp(-sum a_i^2 (x_i-u_i)^2) */
#define IB200_FAM_EXPLIN     9   /* [a(d)]            exp(sum a_i x_i)                         */
#define IB200_FAM_GLTEST     10  /* [p]               5x + sin(x) - p exp(x)            dim 1  */
#define IB200_FAM_COUNT      11
#define IB200_MAX_DIM        8

#define IB200_TR_IF_INF 0
#define IB200_TR_TAN    1
#define IB200_TR_COT    2

#define IB200_RULE_TRAPEZOID 0
#define IB200_RULE_SIMPSON   1

/* per-problem status of the adaptive solves */
#define IB200_ST_CONVERGED  0  /* E <= max(abstol, reltol*|I|) (QuadGK: E<=abstol || E<=reltol*|I|) */
#define IB200_ST_MAXEVALS   1  /* evaluation budget (maxiters) or region capacity exhausted         */
#define IB200_ST_NONFINITE  2  /* non-finite error estimate (QuadGK DomainError / HCubature stop)   */
#define IB200_ST_ROUNDOFF   3  /* QuadGK: segment can no longer be bisected                          */

int32_t     ib200_version(void);
/* context: device ordinal, stream, reusable device scratch (the analogue of QuadGK's segbuf,
   src/Integrals.jl:241-250, and HCubature's buffer, :341-346, kept in cache.cacheval) */
int32_t     ib200_create(int32_t device, ib200_ctx **out);
void        ib200_destroy(ib200_ctx *ctx);
const char *ib200_last_error(ib200_ctx *ctx);            /* ctx may be NULL (create failures) */
int32_t     ib200_set_stream(ib200_ctx *ctx, void *cuda_stream);   /* NULL -> ctx's own stream */
int32_t     ib200_synchronize(ib200_ctx *ctx);
int64_t     ib200_kernel_launches(ib200_ctx *ctx);       /* kernels launched so far (bench)   */
int32_t     ib200_family_nparam(int32_t family, int32_t dim);   /* doubles per problem, <0 bad */
/* timing of the most recent call's kernels on the context's stream, ms (CUDA events) */
double      ib200_last_kernel_ms(ib200_ctx *ctx);
/* optional knobs: key in {"adaptive_warps_per_sm","hcub_capacity","sampled_slabs","hcub_mode",
   "hcub_lpt", "quadgk_mode" (0 auto, 1 warp per problem, 2 thread per segment, 3 thread per problem), "single_round_percent",
   "single_switch_boxes", "single_switch_factor", "single_sub_boxes", "single_arena_mib" (ib200_hcubature_single, two-level solves),
   "hcub_compact" (round-based batches: 1 = 32-byte dyadic box records where they apply [default], 0 = full records; same results),
   "hcub_compact_level" (1..31, halvings per axis such a record may hold before the problem goes to the full-record kernel)} */
int32_t     ib200_set_option(ib200_ctx *ctx, const char *key, int64_t value);

/* -- sampled data: replaces evalrule(data, weights, dim) src/sampled.jl:51-124 with the weights of
      src/trapezoidal.jl:16-45 / src/simpsons.jl:16-87, as dispatched by
      __solvebp_call(::SampledIntegralCache) src/sampled.jl:135-168.
      y is column-major [inner, n, outer] (the array with the integration axis in the middle:
      inner = prod(size(y)[1:dim-1]), outer = prod(size(y)[dim+1:end])); out is [inner, outer].
      x == NULL: uniform grid with step h (AbstractRange branch, trapezoidal.jl:43, simpsons.jl:85);
      else x holds the n grid points and h is ignored.  n == 0 -> IB200_ERR_INVALID
      ("No points to integrate", sampled.jl:55). */
int32_t ib200_sampled(ib200_ctx *ctx, int32_t rule, const double *y, int64_t inner, int64_t n,
                      int64_t outer, double h, const double *x, double *out);
/* -- the same rule on a SLAB of the integration axis (one long sampled array spread over several GPUs along the
      axis that is integrated).  y_slab holds the points offset .. offset + n_slab - 1 of a grid of n_total points ([inner, n_slab, outer]
      column-major); the weights are those of the whole grid (end corrections only at the global ends; x, if given, is the whole grid of
      n_total points).  Without a communicator `out` receives this slab's contribution; with one (ib200_comm_init) every rank calls this with
      its own slab and the contributions are summed over the ranks by ONE all-reduce of inner x outer doubles (NCCL over NVLink), so every
      rank returns the whole integral.  n_slab == 0 is allowed (a rank without points). */
int32_t ib200_sampled_slab(ib200_ctx *ctx, int32_t rule, const double *y_slab, int64_t inner, int64_t n_slab, int64_t outer,
                           int64_t n_total, int64_t offset, double h, const double *x, double *out);
/* Float32 samples on a Float32 grid ("Float32 type preservation", test/interface_tests.jl:524-539): weights and the
   products w[i]*y[i] in float as in the reference, accumulation in double, result rounded to float.
   Complex samples need no entry point of their own: the weights are real, so a Complex{Float64} array
   [inner, n, outer] is the Float64 array [2*inner, n, outer] (test/interface_tests.jl:501-509). */
int32_t ib200_sampled_f32(ib200_ctx *ctx, int32_t rule, const float *y, int64_t inner, int64_t n,
                          int64_t outer, float h, const float *x, float *out);

/* -- fixed rules: replaces evalrule(f,p,lb,ub,nodes,weights) src/quadrules.jl:1-19 under the
      automatic ChangeOfVariables (src/Integrals.jl:192-224), for a batch of problems.
      tensor: the rule is the dim-fold tensor product of (x1d, w1d), n1d^dim nodes
      (test/quadrule_tests.jl:41-44).  nodes: explicit N x dim row-major nodes, N weights. */
int32_t ib200_fixed_rule_tensor_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                      const double *lb, const double *ub, int32_t bounds_per_problem,
                                      const double *params, int32_t nparam, int64_t nprob,
                                      const double *x1d, const double *w1d, int32_t n1d, double *out_I);
int32_t ib200_fixed_rule_nodes_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                     const double *lb, const double *ub, int32_t bounds_per_problem,
                                     const double *params, int32_t nparam, int64_t nprob,
                                     const double *nodes, const double *weights, int64_t N, double *out_I);
/* -- 1-D (composite) Gauss-Legendre: replaces gauss_legendre / composite_gauss_legendre
      ext/IntegralsFastGaussQuadratureExt.jl:11-32 as called from :34-76. */
int32_t ib200_gauss_legendre_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                   const double *lb, const double *ub, int32_t bounds_per_problem,
                                   const double *params, int32_t nparam, int64_t nprob,
                                   const double *nodes, const double *weights, int32_t n,
                                   int32_t subintervals, double *out_I);

/* -- adaptive 1-D: replaces QuadGK.quadgk(f, lb, ub; rtol, atol, maxevals, order = 7) as called at
      src/Integrals.jl:327-335 (scalar out-of-place branch of __solvebp_call(::QuadGKJL) :252-339).
      out_E, out_nevals, out_status may be NULL. */
int32_t ib200_quadgk_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                           const double *lb, const double *ub, int32_t bounds_per_problem,
                           const double *params, int32_t nparam, int64_t nprob,
                           double reltol, double abstol, int64_t maxevals,
                           double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);
/* -- adaptive N-D: replaces HCubature.hcubature / hquadrature(f, lb, ub; rtol, atol, maxevals,
      initdiv) as called at src/Integrals.jl:380-393 (__solvebp_call(::HCubatureJL) :347-397).
      mode: which refinement driver runs the reference's algorithm
        IB200_HCUB_REFERENCE_ORDER  one box per step, largest error first -- the reference's own sequence of boxes
                                    (csrc/hcubature_pair.cuh, hcubature_impl.cuh; 1 <= dim <= 8)
        IB200_HCUB_ROUNDS           every box whose error is above a per-problem threshold is bisected in the same round
                                    (csrc/hcubature_rb.cuh; 2 <= dim <= 8): same rule, split axis, stopping test and
                                    closing sum, a different ORDER of refinement, hence different evaluation counts;
                                    (I, E) agree with mode 0 within the requested tolerance.  maxevals is honoured as a
                                    budget of rule applications, ceil(maxevals / points per rule); the last round may
                                    overshoot it (a round always bisects the boxes within a factor 2 of the worst). */
#define IB200_HCUB_REFERENCE_ORDER 0
#define IB200_HCUB_ROUNDS          1
int32_t ib200_hcubature_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                              const double *lb, const double *ub, int32_t bounds_per_problem,
                              const double *params, int32_t nparam, int64_t nprob,
                              double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t mode,
                              double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

/* -- the same solve with the Gauss-Kronrod rule of another order (QuadGKJL(order = n), src/algorithms.jl:51-56, forwarded
      as `order = alg.order` at src/Integrals.jl:331-334).  x, w, wg: host arrays in the layout of QuadGK.kronrod(n) --
      x[n+1] the non-positive Kronrod nodes in increasing order (x[n] = 0), w[n+1] their weights, wg[(n+1)/2] the Gauss
      weights at x[1], x[3], ...  2 <= order <= 30. */
int32_t ib200_quadgk_rule_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                const double *lb, const double *ub, int32_t bounds_per_problem,
                                const double *params, int32_t nparam, int64_t nprob,
                                double reltol, double abstol, int64_t maxevals,
                                int32_t order, const double *x, const double *w, const double *wg,
                                double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

/* -- vector-valued integrands: QuadGK.quadgk!(f!, result, lb, ub; ..., norm) as called from
      __solvebp_call(::QuadGKJL) for array-valued integrands, src/Integrals.jl:316-326, with QuadGKJL.norm
      (src/algorithms.jl:10-11,56; default LinearAlgebra.norm).  Component c of problem k is the registered family with
      parameter column params[:, c, k] (params: nparam x ncomp x nprob, column-major); all components share the nodes and
      the subdivision; a segment's error is norm(K - G) over the components; stop when E <= max(atol, rtol * norm(I)).
      normkind: 0 = 2-norm, 1 = maximum norm.  1 <= ncomp <= 8.  out_I: ncomp x nprob; out_E, out_nevals, out_status: nprob. */
int32_t ib200_quadgk_vec_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                               const double *lb, const double *ub, int32_t bounds_per_problem,
                               const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                               double reltol, double abstol, int64_t maxevals, int32_t normkind,
                               double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

/* -- the same for N-D: HCubature.hcubature / hquadrature(f, lb, ub; norm, rtol, atol, maxevals, initdiv) with an array-valued
      integrand, as called at src/Integrals.jl:380-393 with `norm = alg.norm` (HCubatureJL.norm, src/algorithms.jl:66-67,
      108-115).  params: nparam x ncomp x nprob; the components share the points and the subdivision; a box's error is
      norm(I - I') over the components, the split axis is chosen from the norms of the fourth differences; stop when
      E <= max(rtol * norm(I), atol).  1 <= dim <= 8, 1 <= ncomp <= 8.  out_I: ncomp x nprob. */
int32_t ib200_hcubature_vec_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                  const double *lb, const double *ub, int32_t bounds_per_problem,
                                  const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                                  double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t normkind,
                                  double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

/* -- parameter sensitivities: I and dI/dp_s for chosen parameters s of a registered family, computed as ONE
      vector-valued integral of [f, df/dp_s1, ..., df/dp_sm] on shared nodes and subdivision.  Replaces differentiating solve() through
      the integrand: ext/IntegralsForwardDiffExt.jl:55-134 ("a vector-valued integral of the primal and dual simultaneously"; with
      QuadGKJL / HCubatureJL the dual numbers flow through the quadrature, whose norm sees the value only -> normkind 2) and
      ext/IntegralsZygoteExt.jl:144-178 (dI/dp as a vector-valued integral of df/dp with the same algorithm -> normkind 0 or 1).
      params: nparam x nprob (ONE column per problem); sens_idx[nsens]: 0-based indices into the parameter column, 1 <= nsens <= 7;
      normkind: 0 = 2-norm over all components, 1 = maximum norm, 2 = the value alone drives the refinement.
      out_I: (1 + nsens) x nprob, column k = [I, dI/dp_s1, ...] of problem k.  The discontinuous Genz family is rejected (the integral
      of its pointwise derivative is not the derivative of its integral). */
int32_t ib200_quadgk_sens_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                const double *lb, const double *ub, int32_t bounds_per_problem,
                                const double *params, int32_t nparam, int64_t nprob,
                                const int32_t *sens_idx, int32_t nsens,
                                double reltol, double abstol, int64_t maxevals, int32_t normkind,
                                double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);
int32_t ib200_hcubature_sens_batch(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                                   const double *lb, const double *ub, int32_t bounds_per_problem,
                                   const double *params, int32_t nparam, int64_t nprob,
                                   const int32_t *sens_idx, int32_t nsens,
                                   double reltol, double abstol, int64_t maxevals, int32_t initdiv, int32_t normkind,
                                   double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

/* -- one large domain (BASELINE config 5): the same HCubature call, src/Integrals.jl:386-393, for a single
      integral that is too expensive for one warp: region-parallel rounds on a device-resident box queue
      (csrc/hcub_single.cu).  With a communicator (ib200_comm_init) every rank of the job calls this with the
      same arguments; every rank keeps the whole box store, evaluates its share of each round's new boxes, pushes
      the results into the peers' inboxes over NVLink (CUDA IPC) and the ranks exchange (integral, error) partial
      sums once per refinement round over NCCL; every rank returns the same result bit for bit.  max_boxes: budget
      of rule applications over all ranks (the analogue of maxiters / points-per-rule); store_boxes: box capacity
      (the same on every rank); 0 = what the budget can create, at most 2^28 boxes (2^26 for an unbounded budget)
      and 70 % of the free HBM (a full store ends the solve with IB200_ST_MAXEVALS).
      TWO LEVELS.  A solve whose store reaches `single_switch_boxes` boxes (option, default 2^20) with the error within
      `single_switch_factor` (default 4096) of the tolerance -- or 16 times that many boxes -- is finished in a second level: every
      stored box becomes an independent sub-problem with the share eps_i = tol E_i^a / sum_j E_j^a (a = dim / (dim + 6)) of the
      tolerance, refined by the round-based batch driver (csrc/hcubature_rb.cuh) in per-warp arenas that are re-used from box to box,
      its boxes summed and dropped when it is done.  Memory is then bounded by the level-1 store however many boxes the integral needs
      (BASELINE config 5 at reltol 1e-9 needs ~10^11), sum eps_i < tol guarantees HCubature's stopping test for the whole integral, and
      the sub-problems need no communication: with a communicator the store is dealt to the ranks in chunks of 64 boxes and the chunk
      sums are gathered once.  The result is the same double on every rank and for every number of ranks.
      out_stats (may be NULL): int64[8] = integrand evaluations, rule applications, boxes (leaves of the whole refinement), level-1
      rounds, status (IB200_ST_*), world size, level-1 boxes handed to level 2 (0 = single level), level-2 sub-problems of this rank that
      did not converge (budget / full arena; the split is in ib200_last_error). */
int32_t ib200_hcubature_single(ib200_ctx *ctx, int32_t family, int32_t dim, int32_t transform,
                               const double *lb, const double *ub, const double *params, int32_t nparam,
                               double reltol, double abstol, int64_t max_boxes, int64_t store_boxes,
                               double *out_I, double *out_E, int64_t *out_stats);
/* multi-GPU plumbing for ib200_hcubature_single: one process per GPU.  Rank 0 obtains a 128-byte id
   (ncclUniqueId) and distributes it by any means (torch.distributed / MPI broadcast); every rank then
   calls ib200_comm_init on its own context.  NCCL is loaded at run time (dlopen of libnccl.so.2). */
int32_t ib200_comm_unique_id(void *id128);
int32_t ib200_comm_init(ib200_ctx *ctx, int32_t nranks, int32_t rank, const void *id128);
int32_t ib200_comm_destroy(ib200_ctx *ctx);

/* -- BatchIntegralFunction bridge: adaptive integration of ONE integral whose integrand is evaluated
      by the CALLER on batches of points -- the engine-side half of BatchIntegralFunction (src/Integrals.jl:274-315, where
      QuadGKJL hands QuadGK a BatchIntegrand; pts is dim x n, one point per column, values a vector of n).  The engine
      keeps the box store, the rule's points, the weighted sums, error estimates, split axes and the refinement rounds of
      ib200_hcubature_single on the device and hands out each round's points; there are no callbacks:
          ib200_batch_open(ctx, dim, transform, lb, ub, reltol, abstol, max_boxes, store_boxes)
          for (;;) { ib200_batch_next(ctx, &n); if (n == 0) break;
                     ib200_batch_points(ctx, x, n);      x[dim x n] column-major, in the caller's ORIGINAL coordinates
                     y[j] = f(x[:, j]) for j < n         (host loop, CuArray broadcast, ...)
                     ib200_batch_values(ctx, y, n); }
          ib200_batch_result(ctx, &I, &E, stats)
      x and y may be host or device pointers (device x is complete when ib200_batch_points returns; device y must be complete --
      its producer stream synchronised or shared through ib200_set_stream -- when ib200_batch_values is called).  The ChangeOfVariables of src/common.jl:97-103 stays inside the engine (the
      Jacobians of a round's points are kept on the device and applied to the values).  Rule: Genz-Malik 7/5 (dim >= 2) or
      GK(7,15) (1-D) per box, as HCubature; a round bisects every box the reference's worst-first loop would provably
      bisect too (csrc/hcub_single.cu), so n grows geometrically and the number of rounds is O(log #boxes).  max_boxes:
      budget of rule applications (maxiters / points per rule); store_boxes: box capacity (0 = derived from max_boxes, at
      most 2^22).  One bridge solve per context at a time; other calls on the context may be interleaved.
      out_stats (may be NULL): int64[6] = integrand evaluations, rule applications, boxes stored, rounds, status, 1. */
int32_t ib200_batch_open(ib200_ctx *ctx, int32_t dim, int32_t transform, const double *lb, const double *ub,
                         double reltol, double abstol, int64_t max_boxes, int64_t store_boxes);
int32_t ib200_batch_next(ib200_ctx *ctx, int64_t *npoints);
int32_t ib200_batch_points(ib200_ctx *ctx, double *x, int64_t npoints);
int32_t ib200_batch_values(ib200_ctx *ctx, const double *y, int64_t npoints);
int32_t ib200_batch_result(ib200_ctx *ctx, double *out_I, double *out_E, int64_t *out_stats);

/* -- device-resident USER integrands (second half).  For integrands outside the registered families the reference
      calls a Julia closure f(x, p) point by point (src/Integrals.jl:327-335, 380-393) or on batches (BatchIntegralFunction, :274-315).
      Here the caller states the integrand in CUDA C++ -- `body` is the body of  __device__ double f(const double *x, const double *p)
      (x: the point, `dim` coordinates in the caller's ORIGINAL variables; p: the parameters) -- and the engine compiles it at run time
      (NVRTC, sm_100a) into one elementwise kernel that it runs between the two kernels of the bridge above: the same rounds, rule,
      change of variables and stopping test, with every evaluation on the device.  ib200_integrand_compile needs no GPU (ctx may be
      NULL); a compile error is returned as IB200_ERR_INVALID with the compiler's log in ib200_last_error.  params: nparam doubles, host
      or device.  out_stats as ib200_batch_result. */
int32_t ib200_integrand_compile(ib200_ctx *ctx, const char *body, int32_t dim, int32_t *handle_out);
int32_t ib200_integrand_release(ib200_ctx *ctx, int32_t handle);
int32_t ib200_integrate_compiled(ib200_ctx *ctx, int32_t handle, int32_t dim, int32_t transform,
                                 const double *lb, const double *ub, const double *params, int32_t nparam,
                                 double reltol, double abstol, int64_t max_boxes, int64_t store_boxes,
                                 double *out_I, double *out_E, int64_t *out_stats);

/* -- measurement helper: dependency-free DFMA chains on every SM; returns TFLOP/s (2 flops/DFMA).
      This is the FP64 roofline denominator. */
int32_t ib200_measure_fp64_peak(ib200_ctx *ctx, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif
