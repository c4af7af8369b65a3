/*
 * oracle.h -- CPU restatement of the Integrals.jl hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for integrals.jl_b200.  It is NOT a product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The CUDA library never links or calls anything in here.
 *
 * Each function cites the reference file:line (relative to the Integrals.jl v5.4.4
 * tree) it restates.  QuadGK.jl (compat "2.11") and HCubature.jl (compat "1.7") are
 * un-vendored third-party dependencies of the reference (Project.toml:60,65); their
 * published algorithms are restated here and anchored on the reference's call sites
 * (src/Integrals.jl:327-335, 380-393) and on the closed-form answers the reference's
 * own tests use (tests/test_oracle_pins.py).  See DESIGN.md "Oracle".
 *
 * PARITY PINNING: the reference is Julia and cannot run here (no Julia, no C sources to build
 * into oracle/_ref), and its tests hold closed-form answers at loose tolerances only -- so at the
 * <= 1e-12 level this oracle is "parity unpinned by reference execution", with one exception: the
 * QuadGK.jl README's published output quadgk(x -> exp(-x^2), 0, 1, rtol=1e-8) =
 * (0.746824132812427, 7.887024366937112e-13), which orc_quadgk reproduces bit for bit.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp; no -ffast-math:
 * the reference does separate multiplies and adds, src/sampled.jl:84).
 */
#ifndef IB200_ORACLE_H
#define IB200_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* generic scalar integrand f(x[0..dim), user) */
typedef double (*orc_integrand)(const double *x, int dim, void *user);

/* registered integrand family ids -- must match include/integrals_b200.h */
enum {
    ORC_FAM_SINXP = 0, ORC_FAM_GENZ_OSC = 1, ORC_FAM_GENZ_PRODPEAK = 2,
    ORC_FAM_GENZ_CORNER = 3, ORC_FAM_GENZ_GAUSS = 4, ORC_FAM_GENZ_C0 = 5,
    ORC_FAM_GENZ_DISCONT = 6, ORC_FAM_COSPROD = 7, ORC_FAM_GAUSSPOLY = 8,
    ORC_FAM_EXPLIN = 9, ORC_FAM_GLTEST = 10, ORC_FAM_COUNT = 11
};
/* transform ids: infinity_handling.jl transformation_if_inf / _tan_inf / _cot_inf */
enum { ORC_TR_IF_INF = 0, ORC_TR_TAN = 1, ORC_TR_COT = 2 };

int    orc_family_nparam(int fam, int dim);
double orc_family_eval(int fam, int dim, const double *x, const double *par);

/* infinity_handling.jl:91-107 (u2t; u2t_tan :182-199 and u2t_cot :293-310 are identical) */
void orc_u2t(double lb, double ub, double *tlb, double *tub);
/* infinity_handling.jl:109-125 (t2ujac), :201-237 (t2ujac_tan), :312-356 (t2ujac_cot) */
void orc_t2ujac(int tr, double t, double lb, double ub, double *u, double *jac);

/* change-of-variables integrand g(t) = f(u(t)) * prod(jac)  (infinity_handling.jl:11-25) */
typedef struct {
    orc_integrand f; void *user; int dim; int tr;
    const double *lb; const double *ub;
} orc_cov;
double orc_cov_eval(const double *t, int dim, void *cov);

/* registered family as an orc_integrand */
typedef struct { int fam; int dim; const double *par; } orc_family_user;
double orc_family_integrand(const double *x, int dim, void *user);

/* src/quadrules.jl:1-19  evalrule(f,p,lb,ub,nodes,weights); nodes is N x dim row-major */
double orc_evalrule(orc_integrand f, void *user, int dim, const double *lb, const double *ub,
                    const double *nodes, const double *weights, int64_t N);
/* same rule where the N = n1d^dim nodes are the tensor product of a 1-D rule, first index
   fastest (Iterators.product order, test/quadrule_tests.jl:41-44); w = w_i*w_j*w_k */
double orc_evalrule_tensor(orc_integrand f, void *user, int dim, const double *lb, const double *ub,
                           const double *x1d, const double *w1d, int n1d);
/* ext/IntegralsFastGaussQuadratureExt.jl:11-22 and :23-32 */
double orc_gauss_legendre(orc_integrand f, void *user, double lb, double ub,
                          const double *nodes, const double *weights, int n);
double orc_composite_gauss_legendre(orc_integrand f, void *user, double lb, double ub,
                                    const double *nodes, const double *weights, int n, int subintervals);

/* sampled weights: trapezoidal.jl:16-20,35-40  simpsons.jl:16-25,40-82 ; i is 1-based */
double orc_trap_uniform_w(int64_t n, double h, int64_t i);
double orc_trap_nonuniform_w(const double *x, int64_t n, int64_t i);
double orc_simpson_uniform_w(int64_t n, double h, int64_t i);
double orc_simpson_nonuniform_w(const double *x, int64_t n, int64_t i);
/* fill w[0..n) ; rule 0 trapezoid 1 simpson ; x==NULL -> uniform(n,h) */
int orc_sampled_weights(int rule, int64_t n, double h, const double *x, double *w);
/* src/sampled.jl:51-124 ; y is column-major [inner, n, outer]; out is [inner, outer];
   strictly sequential in i, separate multiply and add.  nthreads>1 splits (inner,outer). */
int orc_sampled_evalrule(int rule, const double *y, int64_t inner, int64_t n, int64_t outer,
                         double h, const double *x, double *out, int nthreads);

/* the same reduction for Float32 samples on a Float32 grid, all arithmetic in float like the reference's generic code
   (test/interface_tests.jl:524-539); w_out (may be NULL) receives the n float weights */
int orc_sampled_evalrule_f32(int rule, const float *y, int64_t inner, int64_t n, int64_t outer,
                             float h, const float *x, float *out, float *w_out);

/* QuadGK.jl quadgk, order 7 (GK(7,15)), scalar integrand, norm = abs.
   status: 0 ok, 1 maxevals reached, 2 non-finite error estimate (QuadGK DomainError),
   3 endpoint roundoff stopped refinement */
int orc_quadgk(orc_integrand f, void *user, double a, double b, double rtol, double atol,
               int64_t maxevals, double *I, double *E, int64_t *nevals, int64_t *nsegs);
/* QuadGK with the Gauss-Kronrod rule of order n (QuadGKJL(order = n)); x, w, wg in QuadGK.kronrod(n)'s layout */
int orc_quadgk_rule(orc_integrand f, void *user, double a, double b, double rtol, double atol, int64_t maxevals,
                    int n, const double *x, const double *w, const double *wg, double *I, double *E, int64_t *nevals);
int orc_quadgk_rule_batch(int fam, const double *lb, const double *ub, int bounds_per_problem, int tr,
                          const double *params, int nparam, int64_t nprob, double reltol, double abstol, int64_t maxiters,
                          int n, const double *x, const double *w, const double *wg,
                          double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);

/* QuadGK with a vector-valued integrand and a norm (quadgk!(f!, result, ...; norm), src/Integrals.jl:316-326):
   f writes ncomp values; normkind 0 = 2-norm (LinearAlgebra.norm, the QuadGKJL default), 1 = maximum norm */
typedef void (*orc_vec_integrand)(const double *x, int dim, void *user, double *out);
int orc_quadgk_vec(orc_vec_integrand f, void *user, int ncomp, double a, double b, double rtol, double atol,
                   int64_t maxevals, int normkind, double *I, double *E, int64_t *nevals);
/* params: nparam x ncomp x nprob (column-major); I: ncomp x nprob; E: nprob */
int orc_quadgk_vec_batch(int fam, const double *lb, const double *ub, int bounds_per_problem, int tr,
                         const double *params, int nparam, int ncomp, int64_t nprob,
                         double reltol, double abstol, int64_t maxiters, int normkind,
                         double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);

/* HCubature.jl hcubature (dim>=2 Genz-Malik) / hquadrature (dim==1, GK15), norm = abs */
int orc_hcubature(orc_integrand f, void *user, int dim, const double *a, const double *b,
                  double rtol, double atol, int64_t maxevals, int initdiv,
                  double *I, double *E, int64_t *nevals, int64_t *nboxes);
/* HCubature with a vector-valued integrand and a norm (HCubatureJL.norm, src/algorithms.jl:66-67,108-115; forwarded at
   src/Integrals.jl:380-393): box error = norm(I - I'), fourth differences = norms, stop when E <= max(rtol*norm(I), atol) */
int orc_hcubature_vec(orc_vec_integrand f, void *user, int ncomp, int dim, const double *a, const double *b,
                      double rtol, double atol, int64_t maxevals, int initdiv, int normkind,
                      double *I, double *E, int64_t *nevals);
/* params: nparam x ncomp x nprob (column-major); I: ncomp x nprob; E: nprob */
int orc_hcubature_vec_batch(int fam, int dim, const double *lb, const double *ub, int bounds_per_problem, int tr,
                            const double *params, int nparam, int ncomp, int64_t nprob,
                            double reltol, double abstol, int64_t maxiters, int initdiv, int normkind,
                            double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);
/* one application of the cubature rule on box [a,b] (for unit tests) */
void orc_cubrule(orc_integrand f, void *user, int dim, const double *a, const double *b,
                 double *I, double *E, int *kdiv);
void orc_gk15(orc_integrand f, void *user, double a, double b, double *I, double *E);
/* round-synchronous refinement (checker for ib200_hcubature_single; not in the reference -- see oracle.c) */
int orc_hcubature_rounds(orc_integrand f, void *user, int dim, const double *a, const double *b,
                         double rtol, double atol, int64_t max_boxes, int m0,
                         double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds);
int orc_solve_hcubature_rounds(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                               double reltol, double abstol, int64_t max_boxes, int m0,
                               double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds);
/* two-level refinement: rounds until the store holds switch_boxes boxes, then every box as an independent sub-problem with its share of
   the tolerance (checker for ib200_hcubature_single's second level; see oracle.c) */
int orc_hcubature_two_level(orc_integrand f, void *user, int dim, const double *a, const double *b,
                            double rtol, double atol, int64_t max_boxes, int64_t switch_boxes, double switch_factor,
                            int64_t sub_max_boxes, double alpha,
                            double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds, int64_t *nsub);
int orc_hcubature_two_level_family(int fam, int dim, const double *lb, const double *ub, int tr, const double *params,
                                   double reltol, double abstol, int64_t max_boxes, int64_t switch_boxes, double switch_factor,
                                   int64_t sub_max_boxes, double alpha,
                                   double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds, int64_t *nsub);
/* the same rounds started from HCubature's initdiv grid with the budget given in evaluations (checker for
   ib200_hcubature_batch, mode IB200_HCUB_ROUNDS; csrc/hcubature_rb.cuh) */
int orc_hcubature_rounds_init(orc_integrand f, void *user, int dim, const double *a, const double *b,
                              double rtol, double atol, int64_t maxevals, int initdiv,
                              double *I, double *E, int64_t *nevals, int64_t *nboxes, int64_t *rounds);
int orc_hcubature_rounds_family(int fam, int dim, const double *lb, const double *ub, int tr, const double *params,
                                double reltol, double abstol, int64_t max_boxes, int m0,
                                double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds);

/* --- Integrals.jl-level solves with the automatic ChangeOfVariables (common.jl:97-103,
       Integrals.jl:192-224): domain mapped by u2t, integrand by t2ujac, then back-end --- */
int orc_solve_quadgk(orc_integrand f, void *user, double lb, double ub, int tr,
                     double reltol, double abstol, int64_t maxiters,
                     double *I, double *E, int64_t *nevals);
int orc_solve_hcubature(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                        double reltol, double abstol, int64_t maxiters, int initdiv,
                        double *I, double *E, int64_t *nevals);
double orc_solve_quadrule(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                          const double *nodes, const double *weights, int64_t N);
double orc_solve_quadrule_tensor(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                                 const double *x1d, const double *w1d, int n1d);
double orc_solve_gauss_legendre(orc_integrand f, void *user, double lb, double ub, int tr,
                                const double *nodes, const double *weights, int n, int subintervals);

/* --- batch drivers over the registered family (OpenMP, one problem per task): these are
       the CPU baseline legs of bench.py.  params is nparam x nprob column-major (one
       problem per column); lb/ub are dim doubles shared by all problems, or dim x nprob
       when bounds_per_problem != 0. --- */
int orc_quadgk_batch(int fam, const double *lb, const double *ub, int bounds_per_problem, int tr,
                     const double *params, int nparam, int64_t nprob,
                     double reltol, double abstol, int64_t maxiters,
                     double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);
int orc_hcubature_batch(int fam, int dim, const double *lb, const double *ub, int bounds_per_problem, int tr,
                        const double *params, int nparam, int64_t nprob,
                        double reltol, double abstol, int64_t maxiters, int initdiv,
                        double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);
int orc_hcubature_rounds_batch(int fam, int dim, const double *lb, const double *ub, int bounds_per_problem, int tr,
                               const double *params, int nparam, int64_t nprob,
                               double reltol, double abstol, int64_t maxiters, int initdiv,
                               double *I, double *E, int64_t *nevals, int32_t *status, int64_t *rounds, int nthreads);
int orc_fixed_tensor_batch(int fam, int dim, const double *lb, const double *ub, int bounds_per_problem, int tr,
                           const double *params, int nparam, int64_t nprob,
                           const double *x1d, const double *w1d, int n1d, double *I, int nthreads);
int orc_fixed_nodes_batch(int fam, int dim, const double *lb, const double *ub, int bounds_per_problem, int tr,
                          const double *params, int nparam, int64_t nprob,
                          const double *nodes, const double *weights, int64_t N, double *I, int nthreads);
int orc_gauss_legendre_batch(int fam, const double *lb, const double *ub, int bounds_per_problem, int tr,
                             const double *params, int nparam, int64_t nprob,
                             const double *nodes, const double *weights, int n, int subintervals,
                             double *I, int nthreads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
/* parameter sensitivities: value + gradient of a family at a point, and I, dI/dp as one vector-valued integral */
int orc_family_grad(int fam, int dim, const double *x, const double *par, double *out);
int orc_quadgk_sens_batch(int fam, const double *lb, const double *ub, int bounds_per_problem, int tr,
                          const double *params, int nparam, int64_t nprob, const int *idx, int nsens,
                          double reltol, double abstol, int64_t maxiters, int normkind,
                          double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);
int orc_hcubature_sens_batch(int fam, int dim, const double *lb, const double *ub, int bounds_per_problem, int tr,
                             const double *params, int nparam, int64_t nprob, const int *idx, int nsens,
                             double reltol, double abstol, int64_t maxiters, int initdiv, int normkind,
                             double *I, double *E, int64_t *nevals, int32_t *status, int nthreads);

#endif
