/*
 * oracle.c -- CPU restatement of the Integrals.jl hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  Plain C99; compile with -ffp-contract=off so that a*b+c stays a
 * separate multiply and add exactly where the reference (Julia, no auto-FMA) has them.
 * Where this file wants a fused multiply-add (the registered integrand family, which is
 * OUR definition and is mirrored instruction-for-instruction by the device code) it
 * calls fma() explicitly.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXDIM 16

/* ------------------------------------------------------------------------------------------
 * Registered integrand family (not in the reference -- the reference only
 * cites Genz & Malik at src/algorithms.jl:96-98).  Parameter layouts (doubles per problem):
 *   0 SINXP      [p]                         sin(x*p)                      (README.md:30-40)
 *   1 OSC        [a_1..a_d, u_1..u_d]        cos(2 pi u_1 + sum a_i x_i)
 *   2 PRODPEAK   [a, u]                      1 / prod(a_i^-2 + (x_i-u_i)^2)
 *   3 CORNER     [a, u]                      (1 + sum a_i x_i)^-(d+1)
 *   4 GAUSS      [a, u]                      exp(-sum a_i^2 (x_i-u_i)^2)
 *   5 C0         [a, u]                      exp(-sum a_i |x_i-u_i|)
 *   6 DISCONT    [a, u]                      0 if x_1>u_1 or x_2>u_2 else exp(sum a_i x_i)
 *   7 COSPROD    [p_1..p_d]                  prod cos(p_i x_i)             (test/quadrule_tests.jl:6)
 *   8 GAUSSPOLY  [a, u, k_1..k_d, c]         c * prod x_i^k_i * exp(-sum a_i^2 (x_i-u_i)^2)
 *   9 EXPLIN     [a_1..a_d]                  exp(sum a_i x_i)              (test/alt_transformation_tests.jl:22,28)
 *  10 GLTEST     [p]                         5x + sin(x) - p*exp(x)        (test/gaussian_quadrature_tests.jl:48)
 * The evaluation order below is the definition; csrc/family.cuh mirrors it.
 * ------------------------------------------------------------------------------------------ */
int orc_family_nparam(int fam, int dim)
{
    switch (fam) {
    case ORC_FAM_SINXP: case ORC_FAM_GLTEST: return 1;
    case ORC_FAM_GENZ_OSC: case ORC_FAM_GENZ_PRODPEAK: case ORC_FAM_GENZ_CORNER:
    case ORC_FAM_GENZ_GAUSS: case ORC_FAM_GENZ_C0: case ORC_FAM_GENZ_DISCONT: return 2 * dim;
    case ORC_FAM_COSPROD: case ORC_FAM_EXPLIN: return dim;
    case ORC_FAM_GAUSSPOLY: return 3 * dim + 1;
    default: return -1;
    }
}

static const double ORC_TWO_PI = 6.283185307179586476925286766559;

double orc_family_eval(int fam, int dim, const double *x, const double *par)
{
    const double *a = par, *u = par + dim;
    switch (fam) {
    case ORC_FAM_SINXP:
        return sin(x[0] * par[0]);
    case ORC_FAM_GENZ_OSC: {
        double s = ORC_TWO_PI * u[0];
        for (int i = 0; i < dim; i++) s = fma(a[i], x[i], s);
        return cos(s);
    }
    case ORC_FAM_GENZ_PRODPEAK: {
        double pr = 1.0;
        for (int i = 0; i < dim; i++) {
            double t = x[i] - u[i];
            double ainv2 = 1.0 / (a[i] * a[i]);
            pr *= fma(t, t, ainv2);
        }
        return 1.0 / pr;
    }
    case ORC_FAM_GENZ_CORNER: {
        double s = 1.0;
        for (int i = 0; i < dim; i++) s = fma(a[i], x[i], s);
        double r = 1.0 / s, pw = r;
        for (int i = 0; i < dim; i++) pw *= r;
        return pw;
    }
    case ORC_FAM_GENZ_GAUSS: {
        double s = 0.0;
        for (int i = 0; i < dim; i++) {
            double t = a[i] * (x[i] - u[i]);
            s = fma(t, t, s);
        }
        return exp(-s);
    }
    case ORC_FAM_GENZ_C0: {
        double s = 0.0;
        for (int i = 0; i < dim; i++) s = fma(a[i], fabs(x[i] - u[i]), s);
        return exp(-s);
    }
    case ORC_FAM_GENZ_DISCONT: {
        if (x[0] > u[0]) return 0.0;
        if (dim > 1 && x[1] > u[1]) return 0.0;
        double s = 0.0;
        for (int i = 0; i < dim; i++) s = fma(a[i], x[i], s);
        return exp(s);
    }
    case ORC_FAM_COSPROD: {
        double pr = 1.0;
        for (int i = 0; i < dim; i++) pr *= cos(par[i] * x[i]);
        return pr;
    }
    case ORC_FAM_GAUSSPOLY: {
        const double *k = par + 2 * dim;
        double s = 0.0, pr = par[3 * dim];
        for (int i = 0; i < dim; i++) {
            double t = a[i] * (x[i] - u[i]);
            s = fma(t, t, s);
            int ki = (int)k[i];
            for (int j = 0; j < ki; j++) pr *= x[i];
        }
        return pr * exp(-s);
    }
    case ORC_FAM_EXPLIN: {
        double s = 0.0;
        for (int i = 0; i < dim; i++) s = fma(par[i], x[i], s);
        return exp(s);
    }
    case ORC_FAM_GLTEST:
        return fma(5.0, x[0], sin(x[0])) - par[0] * exp(x[0]);
    default:
        return NAN;
    }
}

/* value and gradient with respect to the raw parameters at one point: out = [f, df/dpar[0..nparam-1]] (the
   derivative the reference obtains by differentiating the integrand, ext/IntegralsForwardDiffExt.jl:55-134,
   ext/IntegralsZygoteExt.jl:144-178).  Returns -1 for families without one (discontinuous: its support moves with u). */
int orc_family_grad(int fam, int dim, const double *x, const double *par, double *out)
{
    const double *a = par, *u = par + dim;
    switch (fam) {
    case ORC_FAM_SINXP:
        out[0] = sin(x[0] * par[0]); out[1] = x[0] * cos(x[0] * par[0]);
        return 0;
    case ORC_FAM_GLTEST:
        out[0] = fma(5.0, x[0], sin(x[0])) - par[0] * exp(x[0]); out[1] = -exp(x[0]);
        return 0;
    case ORC_FAM_GENZ_OSC: {
        double s = ORC_TWO_PI * u[0];
        for (int i = 0; i < dim; i++) s = fma(a[i], x[i], s);
        out[0] = cos(s);
        for (int i = 0; i < dim; i++) { out[1 + i] = -(x[i] * sin(s)); out[1 + dim + i] = 0.0; }
        out[1 + dim] = -(ORC_TWO_PI * sin(s));
        return 0;
    }
    case ORC_FAM_GENZ_PRODPEAK: {
        double T[ORC_MAXDIM], f = 1.0;
        for (int i = 0; i < dim; i++) { double t = x[i] - u[i]; T[i] = 1.0 / fma(t, t, 1.0 / (a[i] * a[i])); f *= T[i]; }
        out[0] = f;
        for (int i = 0; i < dim; i++) {
            double t = x[i] - u[i];
            out[1 + i] = f * (2.0 * T[i]) / (a[i] * a[i] * a[i]);
            out[1 + dim + i] = f * (2.0 * t * T[i]);
        }
        return 0;
    }
    case ORC_FAM_GENZ_CORNER: {
        double s = 1.0;
        for (int i = 0; i < dim; i++) s = fma(a[i], x[i], s);
        double r = 1.0 / s, f = r;
        for (int i = 0; i < dim; i++) f *= r;
        out[0] = f;
        double g = -(double)(dim + 1) * f * r;
        for (int i = 0; i < dim; i++) { out[1 + i] = g * x[i]; out[1 + dim + i] = 0.0; }
        return 0;
    }
    case ORC_FAM_GENZ_GAUSS: {
        double s = 0.0;
        for (int i = 0; i < dim; i++) { double t = a[i] * (x[i] - u[i]); s = fma(t, t, s); }
        double f = exp(-s);
        out[0] = f;
        for (int i = 0; i < dim; i++) {
            double d = x[i] - u[i];
            out[1 + i] = -2.0 * a[i] * d * d * f;
            out[1 + dim + i] = 2.0 * a[i] * a[i] * d * f;
        }
        return 0;
    }
    case ORC_FAM_GENZ_C0: {
        double s = 0.0;
        for (int i = 0; i < dim; i++) s = fma(a[i], fabs(x[i] - u[i]), s);
        double f = exp(-s);
        out[0] = f;
        for (int i = 0; i < dim; i++) {
            double d = x[i] - u[i];
            out[1 + i] = -fabs(d) * f;
            out[1 + dim + i] = (d > 0.0 ? a[i] : (d < 0.0 ? -a[i] : 0.0)) * f;
        }
        return 0;
    }
    case ORC_FAM_COSPROD: {
        double f = 1.0;
        for (int i = 0; i < dim; i++) f *= cos(par[i] * x[i]);
        out[0] = f;
        for (int i = 0; i < dim; i++) {
            double o = -(x[i] * sin(par[i] * x[i]));
            for (int k = 0; k < dim; k++) if (k != i) o *= cos(par[k] * x[k]);
            out[1 + i] = o;
        }
        return 0;
    }
    case ORC_FAM_GAUSSPOLY: {
        const double *k = par + 2 * dim;
        double s = 0.0, pr = 1.0;
        for (int i = 0; i < dim; i++) {
            double t = a[i] * (x[i] - u[i]);
            s = fma(t, t, s);
            int ki = (int)k[i];
            for (int j = 0; j < ki; j++) pr *= x[i];
        }
        double g = pr * exp(-s), f = par[3 * dim] * g;
        out[0] = f;
        for (int i = 0; i < dim; i++) {
            double d = x[i] - u[i];
            out[1 + i] = -2.0 * a[i] * d * d * f;
            out[1 + dim + i] = 2.0 * a[i] * a[i] * d * f;
            out[1 + 2 * dim + i] = 0.0;
        }
        out[1 + 3 * dim] = g;
        return 0;
    }
    case ORC_FAM_EXPLIN: {
        double s = 0.0;
        for (int i = 0; i < dim; i++) s = fma(par[i], x[i], s);
        double f = exp(s);
        out[0] = f;
        for (int i = 0; i < dim; i++) out[1 + i] = x[i] * f;
        return 0;
    }
    default:
        return -1;
    }
}

double orc_family_integrand(const double *x, int dim, void *user)
{
    const orc_family_user *fu = (const orc_family_user *)user;
    return orc_family_eval(fu->fam, dim, x, fu->par);
}

/* ------------------------------------------------------------------------------------------
 * Transforms -- src/infinity_handling.jl
 * ------------------------------------------------------------------------------------------ */
static inline double flipsign(double x, double y) { return signbit(y) ? -x : x; }

/* infinity_handling.jl:91-107 */
void orc_u2t(double lb, double ub, double *tlb, double *tub)
{
    if (isinf(lb) && isinf(ub)) { *tlb = flipsign(1.0, lb); *tub = flipsign(1.0, ub); }
    else if (isinf(lb))         { *tlb = flipsign(1.0, lb); *tub = 0.0; }
    else if (isinf(ub))         { *tlb = 0.0;               *tub = flipsign(1.0, ub); }
    else                        { *tlb = -1.0;              *tub = 1.0; }
}

static const double ORC_HALF_PI = 1.5707963267948966; /* Float64(pi/2) */

/* infinity_handling.jl:109-125, 201-237, 312-356 */
void orc_t2ujac(int tr, double t, double lb, double ub, double *u, double *jac)
{
    int il = isinf(lb), iu = isinf(ub);
    if (!il && !iu) {                         /* :121-124 -- same in all three transforms */
        double den = (ub - lb) * 0.5;
        *u = lb + (1.0 + t) * den;
        *jac = den;
        return;
    }
    if (tr == ORC_TR_IF_INF) {
        if (il && iu) {                       /* :112-114 */
            double den = 1.0 / (1.0 - t * t);
            *u = t * den;
            *jac = (1.0 + t * t) * (den * den);
        } else if (il) {                      /* :115-117 */
            double den = 1.0 / (1.0 - flipsign(t, lb));
            *u = ub + t * den;
            *jac = den * den;
        } else {                              /* :118-120 */
            double den = 1.0 / (1.0 - flipsign(t, ub));
            *u = lb + t * den;
            *jac = den * den;
        }
        return;
    }
    if (tr == ORC_TR_COT && !il && iu) {      /* :338-351 */
        double arg = ORC_HALF_PI * (1.0 - t);
        double c = 1.0 / tan(arg);            /* Julia: cot(x) = inv(tan(x)) */
        double csc2 = 1.0 + c * c;
        *u = lb + c;
        *jac = ORC_HALF_PI * csc2;
        return;
    }
    /* tan map (:203-232) -- also what the cot transform uses for the other two cases (:314-337) */
    {
        double arg = ORC_HALF_PI * t;
        double tv = tan(arg);
        double sec2 = 1.0 + tv * tv;
        double base = (il && iu) ? 0.0 : (il ? ub : lb);
        *u = (il && iu) ? tv : base + tv;
        *jac = ORC_HALF_PI * sec2;
    }
}

/* infinity_handling.jl:11-25: u,jac per coordinate; jac = prod left to right; f(u,p)*jac */
double orc_cov_eval(const double *t, int dim, void *vc)
{
    const orc_cov *c = (const orc_cov *)vc;
    double u[ORC_MAXDIM], jac = 1.0;
    for (int i = 0; i < dim; i++) {
        double ji;
        orc_t2ujac(c->tr, t[i], c->lb[i], c->ub[i], &u[i], &ji);
        jac = (i == 0) ? ji : jac * ji;
    }
    return c->f(u, dim, c->user) * jac;
}

/* ------------------------------------------------------------------------------------------
 * Fixed rules -- src/quadrules.jl:1-19 and ext/IntegralsFastGaussQuadratureExt.jl:11-32
 * ------------------------------------------------------------------------------------------ */
double orc_evalrule(orc_integrand f, void *user, int dim, const double *lb, const double *ub,
                    const double *nodes, const double *weights, int64_t N)
{
    double scale[ORC_MAXDIM], shift[ORC_MAXDIM], x[ORC_MAXDIM];
    for (int k = 0; k < dim; k++) { scale[k] = (ub[k] - lb[k]) / 2; shift[k] = (lb[k] + ub[k]) / 2; }
    if (N <= 0) return NAN; /* reference throws ArgumentError("empty quadrature rule") */
    double I = 0.0;
    for (int64_t i = 0; i < N; i++) {
        for (int k = 0; k < dim; k++) x[k] = scale[k] * nodes[i * dim + k] + shift[k];
        double v = weights[i] * f(x, dim, user);
        I = (i == 0) ? v : I + v;
    }
    double ps = scale[0];
    for (int k = 1; k < dim; k++) ps *= scale[k];
    return ps * I;
}

double orc_evalrule_tensor(orc_integrand f, void *user, int dim, const double *lb, const double *ub,
                           const double *x1d, const double *w1d, int n1d)
{
    double scale[ORC_MAXDIM], shift[ORC_MAXDIM], x[ORC_MAXDIM];
    int idx[ORC_MAXDIM];
    for (int k = 0; k < dim; k++) { scale[k] = (ub[k] - lb[k]) / 2; shift[k] = (lb[k] + ub[k]) / 2; idx[k] = 0; }
    if (n1d <= 0) return NAN;
    double I = 0.0; int first = 1;
    for (;;) {
        double w = w1d[idx[0]];
        for (int k = 1; k < dim; k++) w *= w1d[idx[k]];
        for (int k = 0; k < dim; k++) x[k] = scale[k] * x1d[idx[k]] + shift[k];
        double v = w * f(x, dim, user);
        I = first ? v : I + v; first = 0;
        int k = 0;
        while (k < dim && ++idx[k] == n1d) { idx[k] = 0; k++; }
        if (k == dim) break;
    }
    double ps = scale[0];
    for (int k = 1; k < dim; k++) ps *= scale[k];
    return ps * I;
}

double orc_gauss_legendre(orc_integrand f, void *user, double lb, double ub,
                          const double *nodes, const double *weights, int n)
{
    double scale = (ub - lb) / 2, shift = (lb + ub) / 2, x;
    x = scale * nodes[0] + shift;
    double I = weights[0] * f(&x, 1, user);
    for (int i = 1; i < n; i++) {
        x = scale * nodes[i] + shift;
        I += weights[i] * f(&x, 1, user);
    }
    return scale * I;
}

double orc_composite_gauss_legendre(orc_integrand f, void *user, double lb, double ub,
                                    const double *nodes, const double *weights, int n, int subintervals)
{
    double h = (ub - lb) / subintervals;
    double I = orc_gauss_legendre(f, user, lb, lb + h, nodes, weights, n);
    for (int i = 1; i < subintervals; i++) {
        double _lb = lb + i * h, _ub = _lb + h;
        I += orc_gauss_legendre(f, user, _lb, _ub, nodes, weights, n);
    }
    return I;
}

/* ------------------------------------------------------------------------------------------
 * Sampled rules -- src/trapezoidal.jl, src/simpsons.jl, src/sampled.jl   (i is 1-based)
 * ------------------------------------------------------------------------------------------ */
double orc_trap_uniform_w(int64_t n, double h, int64_t i)      /* trapezoidal.jl:16-20 */
{
    return (i == 1 || i == n) ? h / 2 : h;
}
double orc_trap_nonuniform_w(const double *x0, int64_t n, int64_t i) /* trapezoidal.jl:35-40 */
{
    const double *x = x0 - 1; /* 1-based view */
    if (i == 1) return (x[i + 1] - x[i]) / 2;
    if (i == n) return (x[i] - x[i - 1]) / 2;
    return (x[i + 1] - x[i - 1]) / 2;
}
double orc_simpson_uniform_w(int64_t n, double h, int64_t i)   /* simpsons.jl:16-25 */
{
    if (i == 1 || i == n) return 17 * h / 48;
    if (i == 2 || i == n - 1) return 59 * h / 48;
    if (i == 3 || i == n - 2) return 43 * h / 48;
    if (i == 4 || i == n - 3) return 49 * h / 48;
    return h;
}
static inline double cube(double v) { return v * v * v; } /* Julia x^3 -> x*x*x */
double orc_simpson_nonuniform_w(const double *x0, int64_t n, int64_t i) /* simpsons.jl:40-82 */
{
    /* B(k) = x[begin+k]; E(k) = x[end-k] */
#define B(k) x0[(k)]
#define E(k) x0[n - 1 - (k)]
    int64_t j = i - 1;
    if (i == 1)
        return (B(2) - B(0)) / 6 * (2 - (B(2) - B(1)) / (B(1) - B(0)));
    if (n & 1) { /* even number of subintervals */
        if (i == n)
            return (E(0) - E(2)) / 6 * (2 - (E(1) - E(2)) / (E(0) - E(1)));
    } else {
        if (i == n)
            return (E(0) - E(1)) * (2 * (E(0) - E(1)) + 3 * (E(1) - E(2))) / (E(0) - E(2)) / 6;
        if (i == n - 1)
            return (E(1) - E(3)) / 6 * (2 - (E(2) - E(3)) / (E(1) - E(2))) +
                   (E(0) - E(1)) * ((E(0) - E(1)) + 3 * (E(1) - E(2))) / (E(1) - E(2)) / 6;
        if (i == n - 2)
            return cube(E(1) - E(3)) / (E(2) - E(3)) / (E(1) - E(2)) / 6 -
                   cube(E(0) - E(1)) / (E(1) - E(2)) / (E(0) - E(2)) / 6;
    }
    if (j & 1)
        return cube(B(j + 1) - B(j - 1)) / (B(j) - B(j - 1)) / (B(j + 1) - B(j)) / 6;
    return (B(j) - B(j - 2)) / 6 * (2 - (B(j - 1) - B(j - 2)) / (B(j) - B(j - 1))) +
           (B(j + 2) - B(j)) / 6 * (2 - (B(j + 2) - B(j + 1)) / (B(j + 1) - B(j)));
#undef B
#undef E
}

int orc_sampled_weights(int rule, int64_t n, double h, const double *x, double *w)
{
    if (rule != 0 && rule != 1) return -1;
    if (x && rule == 1 && n <= 2) return -2; /* simpsons.jl:46 assertion */
    if (x && rule == 0 && n < 2) return -2;
    for (int64_t i = 1; i <= n; i++) {
        if (!x) w[i - 1] = rule == 0 ? orc_trap_uniform_w(n, h, i) : orc_simpson_uniform_w(n, h, i);
        else    w[i - 1] = rule == 0 ? orc_trap_nonuniform_w(x, n, i) : orc_simpson_nonuniform_w(x, n, i);
    }
    return 0;
}

/* sampled.jl:51-64 (vector), :68-96 (matrix, dim = last), :103-124 (general): always
   out = w1*f1 then out += wi*fi in increasing i, separate multiply and add. */
int orc_sampled_evalrule(int rule, const double *y, int64_t inner, int64_t n, int64_t outer,
                         double h, const double *x, double *out, int nthreads)
{
    if (n == 0) return -3; /* ArgumentError("No points to integrate") */
    double *w = (double *)malloc(sizeof(double) * (size_t)n);
    if (!w) return -4;
    int rc = orc_sampled_weights(rule, n, h, x, w);
    if (rc) { free(w); return rc; }
    if (nthreads < 1) nthreads = 1;
    /* each (o, block of inner) is independent; order within every output is sequential in i */
    const int64_t BLK = 512;
    int64_t nblk = (inner + BLK - 1) / BLK;
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads)
    for (int64_t o = 0; o < outer; o++) {
        for (int64_t b = 0; b < nblk; b++) {
            int64_t j0 = b * BLK, j1 = j0 + BLK < inner ? j0 + BLK : inner;
            const double *yo = y + o * n * inner;
            double *oo = out + o * inner;
            for (int64_t j = j0; j < j1; j++) oo[j] = w[0] * yo[j];
            for (int64_t i = 1; i < n; i++) {
                double wi = w[i];
                const double *yi = yo + i * inner;
                for (int64_t j = j0; j < j1; j++) oo[j] += wi * yi[j];
            }
        }
    }
    free(w);
    return 0;
}

/* Float32 samples on a Float32 grid (test/interface_tests.jl:524-539 "Float32 type preservation"): the reference's
   generic code runs entirely in Float32 -- weights from a Float32 step / grid, products and the sequential sum in
   Float32.  Same formulas as above in float arithmetic. */
static float orc_w_f32(int rule, int64_t n, float h, const float *x0, int64_t i)
{
    if (!x0) {
        if (rule == 0) return (i == 1 || i == n) ? h / 2 : h;
        if (i == 1 || i == n) return 17 * h / 48;
        if (i == 2 || i == n - 1) return 59 * h / 48;
        if (i == 3 || i == n - 2) return 43 * h / 48;
        if (i == 4 || i == n - 3) return 49 * h / 48;
        return h;
    }
    if (rule == 0) {
        if (i == 1) return (x0[1] - x0[0]) / 2;
        if (i == n) return (x0[n - 1] - x0[n - 2]) / 2;
        return (x0[i] - x0[i - 2]) / 2;
    }
#define B(k) x0[(k)]
#define E(k) x0[n - 1 - (k)]
#define CUBE(v) ((v) * (v) * (v))
    int64_t j = i - 1;
    if (i == 1) return (B(2) - B(0)) / 6 * (2 - (B(2) - B(1)) / (B(1) - B(0)));
    if (n & 1) {
        if (i == n) return (E(0) - E(2)) / 6 * (2 - (E(1) - E(2)) / (E(0) - E(1)));
    } else {
        if (i == n) return (E(0) - E(1)) * (2 * (E(0) - E(1)) + 3 * (E(1) - E(2))) / (E(0) - E(2)) / 6;
        if (i == n - 1)
            return (E(1) - E(3)) / 6 * (2 - (E(2) - E(3)) / (E(1) - E(2))) +
                   (E(0) - E(1)) * ((E(0) - E(1)) + 3 * (E(1) - E(2))) / (E(1) - E(2)) / 6;
        if (i == n - 2)
            return CUBE(E(1) - E(3)) / (E(2) - E(3)) / (E(1) - E(2)) / 6 -
                   CUBE(E(0) - E(1)) / (E(1) - E(2)) / (E(0) - E(2)) / 6;
    }
    if (j & 1) return CUBE(B(j + 1) - B(j - 1)) / (B(j) - B(j - 1)) / (B(j + 1) - B(j)) / 6;
    return (B(j) - B(j - 2)) / 6 * (2 - (B(j - 1) - B(j - 2)) / (B(j) - B(j - 1))) +
           (B(j + 2) - B(j)) / 6 * (2 - (B(j + 2) - B(j + 1)) / (B(j + 1) - B(j)));
#undef B
#undef E
#undef CUBE
}

int orc_sampled_evalrule_f32(int rule, const float *y, int64_t inner, int64_t n, int64_t outer,
                             float h, const float *x, float *out, float *w_out)
{
    if (n == 0) return -3;
    if (rule != 0 && rule != 1) return -1;
    if (x && ((rule == 1 && n <= 2) || n < 2)) return -2;
    float *w = (float *)malloc(sizeof(float) * (size_t)n);
    if (!w) return -4;
    for (int64_t i = 1; i <= n; i++) w[i - 1] = orc_w_f32(rule, n, h, x, i);
    for (int64_t o = 0; o < outer; o++) {
        const float *yo = y + o * n * inner;
        float *oo = out + o * inner;
        for (int64_t j = 0; j < inner; j++) {
            volatile float acc = w[0] * yo[j];                 /* volatile: no excess precision, one rounding per operation */
            for (int64_t i = 1; i < n; i++) { volatile float pr = w[i] * yo[i * inner + j]; acc = acc + pr; }
            oo[j] = acc;
        }
    }
    if (w_out) memcpy(w_out, w, sizeof(float) * (size_t)n);
    free(w);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * QuadGK.jl (un-vendored dependency; compat "2.11", Project.toml:65), called at
 * src/Integrals.jl:327-335 with order = 7.  Restated from the package's published algorithm
 * (SURVEY Appendix A.1): Kronrod table = QUADPACK qk15 constants.
 * ------------------------------------------------------------------------------------------ */
static const double GK_X[8] = {
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.0};
static const double GK_W[8] = {
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
static const double GK_WG[4] = {
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

typedef struct { double a, b, I, E; } orc_seg;

static inline double f1d(orc_integrand f, void *user, double x) { return f(&x, 1, user); }

/* QuadGK evalrule (n = 7, odd): symmetric pairs are added before weighting */
static orc_seg gk_evalrule(orc_integrand f, void *user, double a, double b)
{
    /* X(i), W(i), WG(i): 1-based like the Julia source */
#define X(i) GK_X[(i) - 1]
#define W(i) GK_W[(i) - 1]
#define WG(i) GK_WG[(i) - 1]
    double s = 0.5 * (b - a);
    double fg = f1d(f, user, a + (1 + X(2)) * s) + f1d(f, user, a + (1 - X(2)) * s);
    double fk = f1d(f, user, a + (1 + X(1)) * s) + f1d(f, user, a + (1 - X(1)) * s);
    double Ig = fg * WG(1);
    double Ik = fg * W(2) + fk * W(1);
    for (int i = 2; i <= 3; i++) {
        fg = f1d(f, user, a + (1 + X(2 * i)) * s) + f1d(f, user, a + (1 - X(2 * i)) * s);
        fk = f1d(f, user, a + (1 + X(2 * i - 1)) * s) + f1d(f, user, a + (1 - X(2 * i - 1)) * s);
        Ig += fg * WG(i);
        Ik += fg * W(2 * i) + fk * W(2 * i - 1);
    }
    double f0 = f1d(f, user, a + s);
    Ig += f0 * WG(4);
    Ik += f0 * W(8) + (f1d(f, user, a + (1 + X(7)) * s) + f1d(f, user, a + (1 - X(7)) * s)) * W(7);
#undef X
#undef W
#undef WG
    orc_seg r;
    r.a = a; r.b = b; r.I = Ik * s; r.E = fabs(Ik * s - Ig * s);
    return r;
}

void orc_gk15(orc_integrand f, void *user, double a, double b, double *I, double *E)
{
    orc_seg s = gk_evalrule(f, user, a, b);
    *I = s.I; *E = s.E;
}

/* binary max-heap on E with the DataStructures.jl/QuadGK percolate semantics (1-based) */
static void seg_percolate_down(orc_seg *xs, int64_t i, orc_seg x, int64_t len)
{
    int64_t l;
    while ((l = 2 * i) <= len) {
        int64_t r = l + 1;
        int64_t j = (r > len || xs[l].E > xs[r].E) ? l : r;
        if (!(xs[j].E > x.E)) break;
        xs[i] = xs[j];
        i = j;
    }
    xs[i] = x;
}
static void seg_percolate_up(orc_seg *xs, int64_t i, orc_seg x)
{
    int64_t j;
    while ((j = i / 2) >= 1 && x.E > xs[j].E) { xs[i] = xs[j]; i = j; }
    xs[i] = x;
}

static int gk_endpoint_roundoff(double a, double b)
{
    double c = 0.5 * (b - a);
    int at_a = (a == a + (1 + GK_X[0]) * c);
    int at_b = (b == a + (1 - GK_X[0]) * c);
    return at_a || at_b;
}

int orc_quadgk(orc_integrand f, void *user, double a, double b, double rtol, double atol,
               int64_t maxevals, double *Iout, double *Eout, int64_t *nevals_out, int64_t *nsegs_out)
{
    int status = 0;
    orc_seg s0 = gk_evalrule(f, user, a, b);
    double I = s0.I, E = s0.E;
    int64_t numevals = 15, cap = 64, len = 1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    if (isnan(E) || isinf(E)) { *Iout = I; *Eout = E; if (nevals_out) *nevals_out = numevals; if (nsegs_out) *nsegs_out = 1; return 2; }
    if (numevals >= maxevals || E <= atol || E <= rtol * fabs(I)) {
        *Iout = I; *Eout = E;
        if (nevals_out) *nevals_out = numevals;
        if (nsegs_out) *nsegs_out = 1;
        return (E <= atol || E <= rtol * fabs(I)) ? 0 : 1;
    }
    orc_seg *segs = (orc_seg *)malloc(sizeof(orc_seg) * (size_t)(cap + 1)); /* 1-based */
    segs[1] = s0;
    while (E > atol && E > rtol * fabs(I) && numevals < maxevals) {
        /* refine(): pop the biggest-error segment and bisect */
        orc_seg s = segs[1];
        orc_seg y = segs[len--];
        if (len >= 1) seg_percolate_down(segs, 1, y, len);
        double mid = (s.a + s.b) / 2;
        if (gk_endpoint_roundoff(s.a, mid) || gk_endpoint_roundoff(mid, s.b)) {
            len++; seg_percolate_up(segs, len, s);
            status = 3;
            break;
        }
        orc_seg s1 = gk_evalrule(f, user, s.a, mid);
        orc_seg s2 = gk_evalrule(f, user, mid, s.b);
        I = (I - s.I) + s1.I + s2.I;
        E = (E - s.E) + s1.E + s2.E;
        numevals += 30;
        if (isnan(s1.E) || isinf(s1.E) || isnan(s2.E) || isinf(s2.E)) status = 2;
        if (len + 2 > cap) { cap *= 2; segs = (orc_seg *)realloc(segs, sizeof(orc_seg) * (size_t)(cap + 1)); }
        len++; seg_percolate_up(segs, len, s1);
        len++; seg_percolate_up(segs, len, s2);
        if (status == 2) break;
    }
    /* resum(): storage order */
    I = segs[1].I; E = segs[1].E;
    for (int64_t i = 2; i <= len; i++) { I += segs[i].I; E += segs[i].E; }
    free(segs);
    if (status == 0 && !(E <= atol || E <= rtol * fabs(I)) && numevals >= maxevals) status = 1;
    *Iout = I; *Eout = E;
    if (nevals_out) *nevals_out = numevals;
    if (nsegs_out) *nsegs_out = len;
    return status;
}

/* ------------------------------------------------------------------------------------------
 * QuadGK with a Gauss-Kronrod rule of any order n >= 2 (QuadGKJL(order = n), src/algorithms.jl:51-56, forwarded as
 * `order = alg.order` at src/Integrals.jl:331-334).  x, w, wg in QuadGK.kronrod(n)'s layout: x[0..n] the non-positive
 * Kronrod nodes (x[n] = 0), w their weights, wg[0..(n+1)/2) the Gauss weights at x[1], x[3], ...  The rule is QuadGK's
 * evalrule for any order (symmetric pairs first; even orders have no Gauss node at 0); the driver is orc_quadgk's.
 * ------------------------------------------------------------------------------------------ */
static orc_seg gk_evalrule_n(orc_integrand f, void *user, double a, double b, int n, const double *x, const double *w, const double *wg)
{
    const int lenx = n + 1, lenwg = (n + 1) / 2, n1 = 1 - (lenx & 1);     /* n1: 0 even order, 1 odd order */
#define PAIRN(i) (f1d(f, user, a + (1 + x[(i) - 1]) * s) + f1d(f, user, a + (1 - x[(i) - 1]) * s))
    double s = 0.5 * (b - a);
    double fg = PAIRN(2), fk = PAIRN(1);
    double Ig = fg * wg[0];
    double Ik = fg * w[1] + fk * w[0];
    for (int i = 2; i <= lenwg - n1; i++) {
        fg = PAIRN(2 * i); fk = PAIRN(2 * i - 1);
        Ig += fg * wg[i - 1];
        Ik += fg * w[2 * i - 1] + fk * w[2 * i - 2];
    }
    if (n1 == 0) Ik += f1d(f, user, a + s) * w[lenx - 1];
    else {
        double f0 = f1d(f, user, a + s);
        Ig += f0 * wg[lenwg - 1];
        Ik += f0 * w[lenx - 1] + PAIRN(lenx - 1) * w[lenx - 2];
    }
#undef PAIRN
    orc_seg r;
    r.a = a; r.b = b; r.I = Ik * s; r.E = fabs(Ik * s - Ig * s);
    return r;
}

int orc_quadgk_rule(orc_integrand f, void *user, double a, double b, double rtol, double atol, int64_t maxevals,
                    int n, const double *x, const double *w, const double *wg, double *Iout, double *Eout, int64_t *nevals_out)
{
    if (n < 2 || n > 64) return -1;
    const int64_t npts = 2 * (int64_t)n + 1;
    int status = 0;
    orc_seg s0 = gk_evalrule_n(f, user, a, b, n, x, w, wg);
    double I = s0.I, E = s0.E;
    int64_t numevals = npts, cap = 64, len = 1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    orc_seg *segs = (orc_seg *)malloc(sizeof(orc_seg) * (size_t)(cap + 1));
    segs[1] = s0;
    if (isnan(E) || isinf(E)) status = 2;
    else if (!(numevals >= maxevals || E <= atol || E <= rtol * fabs(I))) {
        while (E > atol && E > rtol * fabs(I) && numevals < maxevals) {
            orc_seg s = segs[1];
            orc_seg y = segs[len--];
            if (len >= 1) seg_percolate_down(segs, 1, y, len);
            double mid = (s.a + s.b) / 2;
            double c1 = 0.5 * (mid - s.a), c2 = 0.5 * (s.b - mid);            /* check_endpoint_roundoff with this rule's first node */
            if (s.a == s.a + (1 + x[0]) * c1 || mid == s.a + (1 - x[0]) * c1 || mid == mid + (1 + x[0]) * c2 || s.b == mid + (1 - x[0]) * c2) {
                len++; seg_percolate_up(segs, len, s); status = 3; break;
            }
            orc_seg s1 = gk_evalrule_n(f, user, s.a, mid, n, x, w, wg);
            orc_seg s2 = gk_evalrule_n(f, user, mid, s.b, n, x, w, wg);
            I = (I - s.I) + s1.I + s2.I;
            E = (E - s.E) + s1.E + s2.E;
            numevals += 2 * npts;
            if (isnan(s1.E) || isinf(s1.E) || isnan(s2.E) || isinf(s2.E)) status = 2;
            if (len + 2 > cap) { cap *= 2; segs = (orc_seg *)realloc(segs, sizeof(orc_seg) * (size_t)(cap + 1)); }
            len++; seg_percolate_up(segs, len, s1);
            len++; seg_percolate_up(segs, len, s2);
            if (status == 2) break;
        }
    }
    I = segs[1].I; E = segs[1].E;
    for (int64_t i = 2; i <= len; i++) { I += segs[i].I; E += segs[i].E; }
    free(segs);
    if (status == 0 && !(E <= atol || E <= rtol * fabs(I)) && numevals >= maxevals) status = 1;
    *Iout = I; *Eout = E;
    if (nevals_out) *nevals_out = numevals;
    return status;
}

int orc_quadgk_rule_batch(int fam, const double *lb, const double *ub, int bpp, int tr,
                          const double *params, int nparam, int64_t nprob, double reltol, double abstol, int64_t maxiters,
                          int n, const double *x, const double *w, const double *wg,
                          double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, 1, params + (size_t)k * nparam};
        double l = bpp ? lb[k] : lb[0], u = bpp ? ub[k] : ub[0], tl, tu;
        orc_cov c = {orc_family_integrand, &fu, 1, tr, &l, &u};
        orc_u2t(l, u, &tl, &tu);
        int64_t ne = 0;
        status[k] = orc_quadgk_rule(orc_cov_eval, &c, tl, tu, reltol, abstol, maxiters, n, x, w, wg, &I[k], &E[k], &ne);
        nevals[k] = ne;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * QuadGK with a vector-valued integrand: the reference reaches it through
 * quadgk!(f!, result, lb, ub; norm = alg.norm) (src/Integrals.jl:316-326; QuadGKJL.norm defaults to
 * LinearAlgebra.norm, src/algorithms.jl:10-11,56).  The rule is applied component-wise, the error of a segment
 * is norm(Ik*s - Ig*s), the stopping test E <= max(atol, rtol*norm(I)); everything else is orc_quadgk.
 * normkind: 0 = 2-norm (LinearAlgebra.norm), 1 = maximum norm.
 * ------------------------------------------------------------------------------------------ */
#define ORC_MAXCOMP 16
typedef struct { double a, b, E; double I[ORC_MAXCOMP]; } orc_vseg;

static double vnorm(const double *v, int n, int kind)
{
    double r = 0.0;
    if (kind == 1) { for (int i = 0; i < n; i++) if (fabs(v[i]) > r || isnan(v[i])) r = fabs(v[i]); return r; }
    if (kind == 2) return fabs(v[0]);   /* the first component alone: a ForwardDiff.Dual's norm is the norm of its value */
    /* LinearAlgebra.norm (generic_norm2) rescales by the maximum to avoid overflow; for the magnitudes of integrals
       a plain sqrt(sum of squares) differs from it by rounding only */
    for (int i = 0; i < n; i++) r += v[i] * v[i];
    return sqrt(r);
}

static orc_vseg gk_evalrule_vec(orc_vec_integrand f, void *user, int nc, double a, double b, int normkind)
{
#define X(i) GK_X[(i) - 1]
#define W(i) GK_W[(i) - 1]
#define WG(i) GK_WG[(i) - 1]
    double s = 0.5 * (b - a), x, fa[ORC_MAXCOMP], fb[ORC_MAXCOMP], fg[ORC_MAXCOMP], fk[ORC_MAXCOMP], Ig[ORC_MAXCOMP], Ik[ORC_MAXCOMP];
#define PAIR(dst, xi) do { x = a + (1 + (xi)) * s; f(&x, 1, user, fa); x = a + (1 - (xi)) * s; f(&x, 1, user, fb); \
                           for (int c = 0; c < nc; c++) dst[c] = fa[c] + fb[c]; } while (0)
    PAIR(fg, X(2)); PAIR(fk, X(1));
    for (int c = 0; c < nc; c++) { Ig[c] = fg[c] * WG(1); Ik[c] = fg[c] * W(2) + fk[c] * W(1); }
    for (int i = 2; i <= 3; i++) {
        PAIR(fg, X(2 * i)); PAIR(fk, X(2 * i - 1));
        for (int c = 0; c < nc; c++) { Ig[c] += fg[c] * WG(i); Ik[c] += fg[c] * W(2 * i) + fk[c] * W(2 * i - 1); }
    }
    double f0[ORC_MAXCOMP];
    x = a + s; f(&x, 1, user, f0);
    PAIR(fk, X(7));
    for (int c = 0; c < nc; c++) { Ig[c] += f0[c] * WG(4); Ik[c] += f0[c] * W(8) + fk[c] * W(7); }
#undef PAIR
#undef X
#undef W
#undef WG
    orc_vseg r; double d[ORC_MAXCOMP];
    r.a = a; r.b = b;
    for (int c = 0; c < nc; c++) { r.I[c] = Ik[c] * s; d[c] = Ik[c] * s - Ig[c] * s; }
    r.E = vnorm(d, nc, normkind);
    return r;
}

static void vseg_down(orc_vseg *xs, int64_t i, orc_vseg x, int64_t len)
{
    int64_t l;
    while ((l = 2 * i) <= len) {
        int64_t r = l + 1;
        int64_t j = (r > len || xs[l].E > xs[r].E) ? l : r;
        if (!(xs[j].E > x.E)) break;
        xs[i] = xs[j]; i = j;
    }
    xs[i] = x;
}
static void vseg_up(orc_vseg *xs, int64_t i, orc_vseg x)
{
    int64_t j;
    while ((j = i / 2) >= 1 && x.E > xs[j].E) { xs[i] = xs[j]; i = j; }
    xs[i] = x;
}

int orc_quadgk_vec(orc_vec_integrand f, void *user, int nc, double a, double b, double rtol, double atol,
                   int64_t maxevals, int normkind, double *Iout, double *Eout, int64_t *nevals_out)
{
    if (nc < 1 || nc > ORC_MAXCOMP) return -1;
    int status = 0;
    orc_vseg s0 = gk_evalrule_vec(f, user, nc, a, b, normkind);
    double I[ORC_MAXCOMP], E = s0.E;
    for (int c = 0; c < nc; c++) I[c] = s0.I[c];
    int64_t numevals = 15, cap = 64, len = 1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    orc_vseg *segs = (orc_vseg *)malloc(sizeof(orc_vseg) * (size_t)(cap + 1));
    segs[1] = s0;
    if (isnan(E) || isinf(E)) status = 2;
    else if (!(numevals >= maxevals || E <= atol || E <= rtol * vnorm(I, nc, normkind))) {
        while (E > atol && E > rtol * vnorm(I, nc, normkind) && numevals < maxevals) {
            orc_vseg s = segs[1];
            orc_vseg y = segs[len--];
            if (len >= 1) vseg_down(segs, 1, y, len);
            double mid = (s.a + s.b) / 2;
            if (gk_endpoint_roundoff(s.a, mid) || gk_endpoint_roundoff(mid, s.b)) { len++; vseg_up(segs, len, s); status = 3; break; }
            orc_vseg s1 = gk_evalrule_vec(f, user, nc, s.a, mid, normkind);
            orc_vseg s2 = gk_evalrule_vec(f, user, nc, mid, s.b, normkind);
            for (int c = 0; c < nc; c++) I[c] = (I[c] - s.I[c]) + s1.I[c] + s2.I[c];
            E = (E - s.E) + s1.E + s2.E;
            numevals += 30;
            if (isnan(s1.E) || isinf(s1.E) || isnan(s2.E) || isinf(s2.E)) status = 2;
            if (len + 2 > cap) { cap *= 2; segs = (orc_vseg *)realloc(segs, sizeof(orc_vseg) * (size_t)(cap + 1)); }
            len++; vseg_up(segs, len, s1);
            len++; vseg_up(segs, len, s2);
            if (status == 2) break;
        }
    }
    for (int c = 0; c < nc; c++) I[c] = segs[1].I[c];
    E = segs[1].E;
    for (int64_t i = 2; i <= len; i++) { for (int c = 0; c < nc; c++) I[c] += segs[i].I[c]; E += segs[i].E; }
    free(segs);
    if (status == 0 && !(E <= atol || E <= rtol * vnorm(I, nc, normkind)) && numevals >= maxevals) status = 1;
    for (int c = 0; c < nc; c++) Iout[c] = I[c];
    *Eout = E;
    if (nevals_out) *nevals_out = numevals;
    return status;
}

/* batch driver over the registered family: component c of problem k uses parameter column params[:, c, k]; the
   automatic ChangeOfVariables (common.jl:97-103) is applied to every component */
typedef struct { int fam, nc, nparam, tr; const double *par; double lb, ub; } orc_vecfam_user;
static void orc_vecfam_eval(const double *t, int dim, void *user, double *out)
{
    (void)dim;
    const orc_vecfam_user *u = (const orc_vecfam_user *)user;
    double x, jac;
    orc_t2ujac(u->tr, t[0], u->lb, u->ub, &x, &jac);
    for (int c = 0; c < u->nc; c++) out[c] = orc_family_eval(u->fam, 1, &x, u->par + (size_t)c * u->nparam) * jac;
}
int orc_quadgk_vec_batch(int fam, const double *lb, const double *ub, int bpp, int tr,
                         const double *params, int nparam, int ncomp, int64_t nprob,
                         double reltol, double abstol, int64_t maxiters, int normkind,
                         double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_vecfam_user u = {fam, ncomp, nparam, tr, params + (size_t)k * nparam * ncomp, bpp ? lb[k] : lb[0], bpp ? ub[k] : ub[0]};
        double tl, tu;
        orc_u2t(u.lb, u.ub, &tl, &tu);
        int64_t ne = 0;
        status[k] = orc_quadgk_vec(orc_vecfam_eval, &u, ncomp, tl, tu, reltol, abstol, maxiters, normkind, I + (size_t)k * ncomp, &E[k], &ne);
        nevals[k] = ne;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * HCubature.jl (un-vendored dependency; compat "1.7", Project.toml:60), called at
 * src/Integrals.jl:380-393.  Restated from the published algorithm:
 * Genz-Malik degree-7 rule with embedded degree-5 rule for dim>=2, GK(7,15) for dim==1,
 * binary max-heap of boxes, bisection along the largest fourth difference.
 * ------------------------------------------------------------------------------------------ */
typedef struct { double a[ORC_MAXDIM], b[ORC_MAXDIM], I, E; int kdiv; } orc_box;

static void gm_rule(orc_integrand f, void *user, int n, const double *a, const double *b,
                    double *Iout, double *Eout, int *kdiv_out)
{
    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l4 = l3, l5 = sqrt(9.0 / 19.0);
    const double twon = (double)(1 << n);
    const double w1 = twon * ((12824 - 9120 * n + 400 * n * n) / 19683.0);
    const double w2 = twon * (980 / 6561.0);
    const double w3 = twon * ((1820 - 400 * n) / 19683.0);
    const double w4 = twon * (200 / 19683.0);
    const double w5 = 6859 / 19683.0;
    const double w4p = twon * (25 / 729.0);
    const double w3p = twon * ((265 - 100 * n) / 1458.0);
    const double w2p = twon * (245 / 486.0);
    const double w1p = twon * ((729 - 950 * n + 50 * n * n) / 729.0);

    double c[ORC_MAXDIM], d[ORC_MAXDIM], x[ORC_MAXDIM], divdiff[ORC_MAXDIM];
    double V = 1.0;
    for (int i = 0; i < n; i++) {
        c[i] = 0.5 * (a[i] + b[i]);
        d[i] = 0.5 * fabs(b[i] - a[i]);
        V = (i == 0) ? d[i] : V * d[i];
    }
    for (int i = 0; i < n; i++) x[i] = c[i];
    double f1 = f(x, n, user), f2 = 0.0, f3 = 0.0;
    double twelvef1 = 12 * f1;
    for (int i = 0; i < n; i++) {
        double p2 = d[i] * l2, p3 = d[i] * l3;
        x[i] = c[i] + p2; double fa = f(x, n, user);
        x[i] = c[i] - p2; double f2i = fa + f(x, n, user);
        x[i] = c[i] + p3; fa = f(x, n, user);
        x[i] = c[i] - p3; double f3i = fa + f(x, n, user);
        x[i] = c[i];
        f2 += f2i; f3 += f3i;
        divdiff[i] = fabs(f3i + twelvef1 - 7 * f2i);
    }
    double f4 = 0.0;
    for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++) {
            double pi_ = d[i] * l4, pj = d[j] * l4;
            for (int s = 0; s < 4; s++) {
                x[i] = (s & 2) ? c[i] - pi_ : c[i] + pi_;
                x[j] = (s & 1) ? c[j] - pj : c[j] + pj;
                f4 += f(x, n, user);
            }
            x[i] = c[i]; x[j] = c[j];
        }
    double f5 = 0.0;
    for (int s = 0; s < (1 << n); s++) {
        for (int i = 0; i < n; i++) {
            double p = d[i] * l5;
            x[i] = ((s >> i) & 1) ? c[i] - p : c[i] + p;
        }
        f5 += f(x, n, user);
    }
    double I = V * (w1 * f1 + w2 * f2 + w3 * f3 + w4 * f4 + w5 * f5);
    double Ip = V * (w1p * f1 + w2p * f2 + w3p * f3 + w4p * f4);
    double E = fabs(I - Ip);
    int kdiv = 0;
    double maxdivdiff = 0.0;
    double deltaf = E / (pow(10.0, n) * V);
    for (int i = 0; i < n; i++) {
        double delta = divdiff[i] - maxdivdiff;
        if (delta > deltaf) { kdiv = i; maxdivdiff = divdiff[i]; }
        else if (fabs(delta) <= deltaf && d[i] > d[kdiv]) kdiv = i;
    }
    *Iout = I; *Eout = E; *kdiv_out = kdiv;
}

/* HCubature's 1-D rule: GK(7,15) about the centre, signed half-width */
static void gk_rule_hc(orc_integrand f, void *user, double a, double b, double *Iout, double *Eout, int *kdiv)
{
    double c = (a + b) * 0.5, D = (b - a) * 0.5;
    double fx0 = f1d(f, user, c);
    double I = fx0 * GK_W[7], Ip = fx0 * GK_WG[3];
    for (int i = 1; i <= 7; i++) {
        double Dx = D * GK_X[i - 1];
        double fx = f1d(f, user, c + Dx) + f1d(f, user, c - Dx);
        I += fx * GK_W[i - 1];
        if ((i & 1) == 0) Ip += fx * GK_WG[(i >> 1) - 1];
    }
    I *= D; Ip *= D;
    *Iout = I; *Eout = fabs(I - Ip); *kdiv = 0;
}

void orc_cubrule(orc_integrand f, void *user, int dim, const double *a, const double *b,
                 double *I, double *E, int *kdiv)
{
    if (dim == 1) gk_rule_hc(f, user, a[0], b[0], I, E, kdiv);
    else gm_rule(f, user, dim, a, b, I, E, kdiv);
}

/* DataStructures.jl BinaryMaxHeap bubble semantics (1-based valtree) */
static void box_bubble_up(orc_box *t, int64_t i)
{
    orc_box v = t[i];
    while (i > 1) {
        int64_t p = i >> 1;
        if (!(v.E > t[p].E)) break;
        t[i] = t[p]; i = p;
    }
    t[i] = v;
}
static void box_bubble_down(orc_box *t, int64_t i, int64_t n)
{
    orc_box v = t[i];
    int swapped = 1;
    int64_t last_parent = n >> 1;
    while (swapped && i <= last_parent) {
        int64_t lc = i << 1;
        if (lc < n) {
            int64_t rc = lc + 1;
            if (t[rc].E > t[lc].E) {
                if (t[rc].E > v.E) { t[i] = t[rc]; i = rc; } else swapped = 0;
            } else {
                if (t[lc].E > v.E) { t[i] = t[lc]; i = lc; } else swapped = 0;
            }
        } else {
            if (t[lc].E > v.E) { t[i] = t[lc]; i = lc; } else swapped = 0;
        }
    }
    t[i] = v;
}

int orc_hcubature(orc_integrand f, void *user, int dim, const double *a, const double *b,
                  double rtol, double atol, int64_t maxevals, int initdiv,
                  double *Iout, double *Eout, int64_t *nevals_out, int64_t *nboxes_out)
{
    if (dim < 1 || dim > ORC_MAXDIM || initdiv < 1 || rtol < 0 || atol < 0 || maxevals < 0) return -1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    const int64_t evals_per_box = dim == 1 ? 15 : 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim);
    int64_t numevals = evals_per_box, cap = 256, len = 0;
    double Delta[ORC_MAXDIM];
    orc_box first;
    double prodD = 1.0;
    for (int i = 0; i < dim; i++) {
        Delta[i] = (b[i] - a[i]) / initdiv;
        prodD *= Delta[i];
        first.a[i] = a[i];
        first.b[i] = initdiv == 1 ? b[i] : a[i] + Delta[i];
    }
    orc_cubrule(f, user, dim, first.a, first.b, &first.I, &first.E, &first.kdiv);
    double I = first.I, E = first.E;
    if (prodD == 0.0) { *Iout = I; *Eout = E; if (nevals_out) *nevals_out = numevals; if (nboxes_out) *nboxes_out = 1; return 0; }
    orc_box *t = (orc_box *)malloc(sizeof(orc_box) * (size_t)(cap + 1));
    t[++len] = first;
    if (initdiv > 1) {
        int c[ORC_MAXDIM] = {0};
        int64_t total = 1;
        for (int i = 0; i < dim; i++) total *= initdiv;
        for (int64_t q = 1; q < total; q++) { /* CartesianIndices order, first index fastest; skip the first box */
            int k = 0;
            while (++c[k] == initdiv) { c[k] = 0; k++; }
            orc_box bx;
            for (int i = 0; i < dim; i++) {
                bx.a[i] = a[i] + c[i] * Delta[i];
                bx.b[i] = (c[i] == initdiv - 1) ? b[i] : bx.a[i] + Delta[i];
            }
            orc_cubrule(f, user, dim, bx.a, bx.b, &bx.I, &bx.E, &bx.kdiv);
            I += bx.I; E += bx.E; numevals += evals_per_box;
            if (len + 1 > cap) { cap *= 2; t = (orc_box *)realloc(t, sizeof(orc_box) * (size_t)(cap + 1)); }
            t[++len] = bx; box_bubble_up(t, len);
        }
    }
    int status = 0;
    for (;;) {
        if (E <= fmax(rtol * fabs(I), atol)) break;
        if (numevals >= maxevals) { status = 1; break; }
        if (!isfinite(E)) { status = 2; break; }
        /* pop the box with the largest error */
        orc_box box = t[1];
        if (len == 1) len = 0;
        else { t[1] = t[len--]; box_bubble_down(t, 1, len); }
        int k = box.kdiv;
        double w = (box.b[k] - box.a[k]) * 0.5;
        orc_box b1 = box, b2 = box;
        b1.a[k] = box.a[k] + w;            /* upper half first */
        orc_cubrule(f, user, dim, b1.a, b1.b, &b1.I, &b1.E, &b1.kdiv);
        b2.b[k] = box.b[k] - w;
        orc_cubrule(f, user, dim, b2.a, b2.b, &b2.I, &b2.E, &b2.kdiv);
        numevals += 2 * evals_per_box;
        I += b1.I + b2.I - box.I;
        E += b1.E + b2.E - box.E;
        if (len + 2 > cap) { cap *= 2; t = (orc_box *)realloc(t, sizeof(orc_box) * (size_t)(cap + 1)); }
        t[++len] = b1; box_bubble_up(t, len);
        t[++len] = b2; box_bubble_up(t, len);
    }
    /* roundoff paranoia: re-sum in storage order */
    I = 0.0; E = 0.0;
    for (int64_t i = 1; i <= len; i++) { I += t[i].I; E += t[i].E; }
    free(t);
    *Iout = I; *Eout = E;
    if (nevals_out) *nevals_out = numevals;
    if (nboxes_out) *nboxes_out = len;
    return status;
}

/* ------------------------------------------------------------------------------------------
 * HCubature with a vector-valued integrand and a norm: HCubatureJL.norm
 * (src/algorithms.jl:66-67,108-115) is forwarded as `norm = alg.norm` at src/Integrals.jl:380-393.
 * The rule is applied component-wise on shared points; a box's error is norm(I - I') over the
 * components, the fourth differences that choose the split axis are norms too, the stopping test
 * is E <= max(rtol*norm(I), atol); everything else is orc_hcubature.
 * ------------------------------------------------------------------------------------------ */
typedef struct { double a[ORC_MAXDIM], b[ORC_MAXDIM], I[ORC_MAXCOMP], E; int kdiv; } orc_vbox;

static void vaxpy(double *acc, const double *v, int nc) { for (int c = 0; c < nc; c++) acc[c] += v[c]; }

static void gm_rule_vec(orc_vec_integrand f, void *user, int nc, int n, const double *a, const double *b, int normkind,
                        double *Iout, double *Eout, int *kdiv_out)
{
    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l4 = l3, l5 = sqrt(9.0 / 19.0);
    const double twon = (double)(1 << n);
    const double w1 = twon * ((12824 - 9120 * n + 400 * n * n) / 19683.0);
    const double w2 = twon * (980 / 6561.0);
    const double w3 = twon * ((1820 - 400 * n) / 19683.0);
    const double w4 = twon * (200 / 19683.0);
    const double w5 = 6859 / 19683.0;
    const double w4p = twon * (25 / 729.0);
    const double w3p = twon * ((265 - 100 * n) / 1458.0);
    const double w2p = twon * (245 / 486.0);
    const double w1p = twon * ((729 - 950 * n + 50 * n * n) / 729.0);

    double c[ORC_MAXDIM], d[ORC_MAXDIM], x[ORC_MAXDIM], divdiff[ORC_MAXDIM];
    double f1[ORC_MAXCOMP], f2[ORC_MAXCOMP] = {0}, f3[ORC_MAXCOMP] = {0}, f4[ORC_MAXCOMP] = {0}, f5[ORC_MAXCOMP] = {0};
    double fa[ORC_MAXCOMP], fb[ORC_MAXCOMP], f2i[ORC_MAXCOMP], f3i[ORC_MAXCOMP], dd[ORC_MAXCOMP];
    double V = 1.0;
    for (int i = 0; i < n; i++) {
        c[i] = 0.5 * (a[i] + b[i]);
        d[i] = 0.5 * fabs(b[i] - a[i]);
        V = (i == 0) ? d[i] : V * d[i];
    }
    for (int i = 0; i < n; i++) x[i] = c[i];
    f(x, n, user, f1);
    for (int i = 0; i < n; i++) {
        double p2 = d[i] * l2, p3 = d[i] * l3;
        x[i] = c[i] + p2; f(x, n, user, fa);
        x[i] = c[i] - p2; f(x, n, user, fb);
        for (int k = 0; k < nc; k++) f2i[k] = fa[k] + fb[k];
        x[i] = c[i] + p3; f(x, n, user, fa);
        x[i] = c[i] - p3; f(x, n, user, fb);
        for (int k = 0; k < nc; k++) f3i[k] = fa[k] + fb[k];
        x[i] = c[i];
        vaxpy(f2, f2i, nc); vaxpy(f3, f3i, nc);
        for (int k = 0; k < nc; k++) dd[k] = f3i[k] + 12 * f1[k] - 7 * f2i[k];
        divdiff[i] = vnorm(dd, nc, normkind);
    }
    for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++) {
            double pi_ = d[i] * l4, pj = d[j] * l4;
            for (int s = 0; s < 4; s++) {
                x[i] = (s & 2) ? c[i] - pi_ : c[i] + pi_;
                x[j] = (s & 1) ? c[j] - pj : c[j] + pj;
                f(x, n, user, fa); vaxpy(f4, fa, nc);
            }
            x[i] = c[i]; x[j] = c[j];
        }
    for (int s = 0; s < (1 << n); s++) {
        for (int i = 0; i < n; i++) {
            double p = d[i] * l5;
            x[i] = ((s >> i) & 1) ? c[i] - p : c[i] + p;
        }
        f(x, n, user, fa); vaxpy(f5, fa, nc);
    }
    for (int k = 0; k < nc; k++) {
        double I = V * (w1 * f1[k] + w2 * f2[k] + w3 * f3[k] + w4 * f4[k] + w5 * f5[k]);
        double Ip = V * (w1p * f1[k] + w2p * f2[k] + w3p * f3[k] + w4p * f4[k]);
        Iout[k] = I; dd[k] = I - Ip;
    }
    double E = vnorm(dd, nc, normkind);
    int kdiv = 0;
    double maxdivdiff = 0.0;
    double deltaf = E / (pow(10.0, n) * V);
    for (int i = 0; i < n; i++) {
        double delta = divdiff[i] - maxdivdiff;
        if (delta > deltaf) { kdiv = i; maxdivdiff = divdiff[i]; }
        else if (fabs(delta) <= deltaf && d[i] > d[kdiv]) kdiv = i;
    }
    *Eout = E; *kdiv_out = kdiv;
}

/* HCubature's 1-D rule (GK(7,15) about the centre, signed half-width) on a vector */
static void gk_rule_hc_vec(orc_vec_integrand f, void *user, int nc, double a, double b, int normkind, double *Iout, double *Eout, int *kdiv)
{
    double c = (a + b) * 0.5, D = (b - a) * 0.5, x;
    double fa[ORC_MAXCOMP], fb[ORC_MAXCOMP], I[ORC_MAXCOMP], Ip[ORC_MAXCOMP], dd[ORC_MAXCOMP];
    x = c; f(&x, 1, user, fa);
    for (int k = 0; k < nc; k++) { I[k] = fa[k] * GK_W[7]; Ip[k] = fa[k] * GK_WG[3]; }
    for (int i = 1; i <= 7; i++) {
        double Dx = D * GK_X[i - 1];
        x = c + Dx; f(&x, 1, user, fa);
        x = c - Dx; f(&x, 1, user, fb);
        for (int k = 0; k < nc; k++) {
            double fx = fa[k] + fb[k];
            I[k] += fx * GK_W[i - 1];
            if ((i & 1) == 0) Ip[k] += fx * GK_WG[(i >> 1) - 1];
        }
    }
    for (int k = 0; k < nc; k++) { I[k] *= D; Ip[k] *= D; Iout[k] = I[k]; dd[k] = I[k] - Ip[k]; }
    *Eout = vnorm(dd, nc, normkind); *kdiv = 0;
}

static void cubrule_vec(orc_vec_integrand f, void *user, int nc, int dim, const double *a, const double *b, int normkind,
                        double *I, double *E, int *kdiv)
{
    if (dim == 1) gk_rule_hc_vec(f, user, nc, a[0], b[0], normkind, I, E, kdiv);
    else gm_rule_vec(f, user, nc, dim, a, b, normkind, I, E, kdiv);
}

/* the BinaryMaxHeap of orc_hcubature on vector boxes */
static void vbox_bubble_up(orc_vbox *t, int64_t i)
{
    orc_vbox v = t[i];
    while (i > 1) {
        int64_t p = i >> 1;
        if (!(v.E > t[p].E)) break;
        t[i] = t[p]; i = p;
    }
    t[i] = v;
}
static void vbox_bubble_down(orc_vbox *t, int64_t i, int64_t n)
{
    orc_vbox v = t[i];
    int swapped = 1;
    int64_t last_parent = n >> 1;
    while (swapped && i <= last_parent) {
        int64_t lc = i << 1;
        if (lc < n) {
            int64_t rc = lc + 1;
            if (t[rc].E > t[lc].E) {
                if (t[rc].E > v.E) { t[i] = t[rc]; i = rc; } else swapped = 0;
            } else {
                if (t[lc].E > v.E) { t[i] = t[lc]; i = lc; } else swapped = 0;
            }
        } else {
            if (t[lc].E > v.E) { t[i] = t[lc]; i = lc; } else swapped = 0;
        }
    }
    t[i] = v;
}

int orc_hcubature_vec(orc_vec_integrand f, void *user, int nc, int dim, const double *a, const double *b,
                      double rtol, double atol, int64_t maxevals, int initdiv, int normkind,
                      double *Iout, double *Eout, int64_t *nevals_out)
{
    if (nc < 1 || nc > ORC_MAXCOMP || dim < 1 || dim > ORC_MAXDIM || initdiv < 1 || rtol < 0 || atol < 0 || maxevals < 0) return -1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    const int64_t evals_per_box = dim == 1 ? 15 : 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim);
    int64_t numevals = evals_per_box, cap = 256, len = 0;
    double Delta[ORC_MAXDIM], I[ORC_MAXCOMP];
    orc_vbox first;
    double prodD = 1.0;
    for (int i = 0; i < dim; i++) {
        Delta[i] = (b[i] - a[i]) / initdiv;
        prodD *= Delta[i];
        first.a[i] = a[i];
        first.b[i] = initdiv == 1 ? b[i] : a[i] + Delta[i];
    }
    cubrule_vec(f, user, nc, dim, first.a, first.b, normkind, first.I, &first.E, &first.kdiv);
    double E = first.E;
    for (int c = 0; c < nc; c++) I[c] = first.I[c];
    if (prodD == 0.0) { for (int c = 0; c < nc; c++) Iout[c] = I[c]; *Eout = E; if (nevals_out) *nevals_out = numevals; return 0; }
    orc_vbox *t = (orc_vbox *)malloc(sizeof(orc_vbox) * (size_t)(cap + 1));
    t[++len] = first;
    if (initdiv > 1) {
        int ci[ORC_MAXDIM] = {0};
        int64_t total = 1;
        for (int i = 0; i < dim; i++) total *= initdiv;
        for (int64_t q = 1; q < total; q++) {
            int k = 0;
            while (++ci[k] == initdiv) { ci[k] = 0; k++; }
            orc_vbox bx;
            for (int i = 0; i < dim; i++) {
                bx.a[i] = a[i] + ci[i] * Delta[i];
                bx.b[i] = (ci[i] == initdiv - 1) ? b[i] : bx.a[i] + Delta[i];
            }
            cubrule_vec(f, user, nc, dim, bx.a, bx.b, normkind, bx.I, &bx.E, &bx.kdiv);
            vaxpy(I, bx.I, nc); E += bx.E; numevals += evals_per_box;
            if (len + 1 > cap) { cap *= 2; t = (orc_vbox *)realloc(t, sizeof(orc_vbox) * (size_t)(cap + 1)); }
            t[++len] = bx; vbox_bubble_up(t, len);
        }
    }
    int status = 0;
    for (;;) {
        if (E <= fmax(rtol * vnorm(I, nc, normkind), atol)) break;
        if (numevals >= maxevals) { status = 1; break; }
        if (!isfinite(E)) { status = 2; break; }
        orc_vbox box = t[1];
        if (len == 1) len = 0;
        else { t[1] = t[len--]; vbox_bubble_down(t, 1, len); }
        int k = box.kdiv;
        double w = (box.b[k] - box.a[k]) * 0.5;
        orc_vbox b1 = box, b2 = box;
        b1.a[k] = box.a[k] + w;            /* upper half first */
        cubrule_vec(f, user, nc, dim, b1.a, b1.b, normkind, b1.I, &b1.E, &b1.kdiv);
        b2.b[k] = box.b[k] - w;
        cubrule_vec(f, user, nc, dim, b2.a, b2.b, normkind, b2.I, &b2.E, &b2.kdiv);
        numevals += 2 * evals_per_box;
        for (int c = 0; c < nc; c++) I[c] += b1.I[c] + b2.I[c] - box.I[c];
        E += b1.E + b2.E - box.E;
        if (len + 2 > cap) { cap *= 2; t = (orc_vbox *)realloc(t, sizeof(orc_vbox) * (size_t)(cap + 1)); }
        t[++len] = b1; vbox_bubble_up(t, len);
        t[++len] = b2; vbox_bubble_up(t, len);
    }
    for (int c = 0; c < nc; c++) I[c] = 0.0;
    E = 0.0;
    for (int64_t i = 1; i <= len; i++) { vaxpy(I, t[i].I, nc); E += t[i].E; }
    free(t);
    for (int c = 0; c < nc; c++) Iout[c] = I[c];
    *Eout = E;
    if (nevals_out) *nevals_out = numevals;
    return status;
}

/* batch driver over the registered family: component c of problem k uses parameter column params[:, c, k]; the
   automatic ChangeOfVariables (common.jl:97-103) is shared by the components */
typedef struct { int fam, nc, nparam, tr; const double *par; const double *lb, *ub; } orc_vecfam_nd_user;
static void orc_vecfam_nd_eval(const double *t, int dim, void *user, double *out)
{
    const orc_vecfam_nd_user *u = (const orc_vecfam_nd_user *)user;
    double x[ORC_MAXDIM], jac = 1.0;
    for (int i = 0; i < dim; i++) {
        double ji;
        orc_t2ujac(u->tr, t[i], u->lb[i], u->ub[i], &x[i], &ji);
        jac = (i == 0) ? ji : jac * ji;
    }
    for (int c = 0; c < u->nc; c++) out[c] = orc_family_eval(u->fam, dim, x, u->par + (size_t)c * u->nparam) * jac;
}
int orc_hcubature_vec_batch(int fam, int dim, const double *lb, const double *ub, int bpp, int tr,
                            const double *params, int nparam, int ncomp, int64_t nprob,
                            double reltol, double abstol, int64_t maxiters, int initdiv, int normkind,
                            double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (orc_family_nparam(fam, dim) != nparam || ncomp < 1 || ncomp > ORC_MAXCOMP) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        const double *l = bpp ? lb + k * dim : lb, *u = bpp ? ub + k * dim : ub;
        orc_vecfam_nd_user fu = {fam, ncomp, nparam, tr, params + (size_t)k * nparam * ncomp, l, u};
        double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
        for (int i = 0; i < dim; i++) orc_u2t(l[i], u[i], &tl[i], &tu[i]);
        int64_t ne = 0;
        int st = orc_hcubature_vec(orc_vecfam_nd_eval, &fu, ncomp, dim, tl, tu, reltol, abstol, maxiters, initdiv, normkind,
                                   I + (size_t)k * ncomp, &E[k], &ne);
        if (nevals) nevals[k] = ne;
        if (status) status[k] = st;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Parameter sensitivities: I and dI/dp_s for chosen parameters s, as ONE vector-valued integral of
 * [f, df/dp_s1, ...] on shared nodes and subdivision -- the reference's ForwardDiff path ("a vector-valued integral of the primal and dual
 * simultaneously", ext/IntegralsForwardDiffExt.jl:55-134; with QuadGKJL / HCubatureJL the duals flow through the quadrature and its norm
 * sees the value only: normkind 2) and its Zygote path (a vector-valued integral of df/dp, ext/IntegralsZygoteExt.jl:144-178: normkind 0).
 * Checker of ib200_quadgk_sens_batch / ib200_hcubature_sens_batch.  params: nparam x nprob; I: (1 + nsens) x nprob.
 * ------------------------------------------------------------------------------------------ */
#define ORC_MAXPARAM 32
typedef struct { int fam, dim, nparam, tr, nsens; const int *idx; const double *par; const double *lb, *ub; } orc_sens_user;
static void orc_sens_eval(const double *t, int dim, void *user, double *out)
{
    const orc_sens_user *u = (const orc_sens_user *)user;
    double x[ORC_MAXDIM], jac = 1.0, g[1 + ORC_MAXPARAM];
    for (int i = 0; i < dim; i++) {
        double ji;
        orc_t2ujac(u->tr, t[i], u->lb[i], u->ub[i], &x[i], &ji);
        jac = (i == 0) ? ji : jac * ji;
    }
    orc_family_grad(u->fam, dim, x, u->par, g);
    out[0] = g[0] * jac;
    for (int s = 0; s < u->nsens; s++) out[1 + s] = g[1 + u->idx[s]] * jac;
}
static int sens_args_ok(int fam, int dim, int nparam, const int *idx, int nsens)
{
    double x[ORC_MAXDIM] = {0}, par[ORC_MAXPARAM] = {0}, g[1 + ORC_MAXPARAM];
    for (int i = 0; i < ORC_MAXPARAM; i++) par[i] = 1.0;
    if (orc_family_nparam(fam, dim) != nparam || nparam > ORC_MAXPARAM || nsens < 0 || 1 + nsens > ORC_MAXCOMP) return 0;
    for (int s = 0; s < nsens; s++) if (idx[s] < 0 || idx[s] >= nparam) return 0;
    return orc_family_grad(fam, dim, x, par, g) == 0;
}
int orc_quadgk_sens_batch(int fam, const double *lb, const double *ub, int bpp, int tr,
                          const double *params, int nparam, int64_t nprob, const int *idx, int nsens,
                          double reltol, double abstol, int64_t maxiters, int normkind,
                          double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (!sens_args_ok(fam, 1, nparam, idx, nsens)) return -1;
    if (nthreads < 1) nthreads = 1;
    const int nc = 1 + nsens;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        const double l = bpp ? lb[k] : lb[0], uu = bpp ? ub[k] : ub[0];
        orc_sens_user u = {fam, 1, nparam, tr, nsens, idx, params + (size_t)k * nparam, &l, &uu};
        double tl, tu;
        orc_u2t(l, uu, &tl, &tu);
        int64_t ne = 0;
        status[k] = orc_quadgk_vec(orc_sens_eval, &u, nc, tl, tu, reltol, abstol, maxiters, normkind, I + (size_t)k * nc, &E[k], &ne);
        nevals[k] = ne;
    }
    return 0;
}
int orc_hcubature_sens_batch(int fam, int dim, const double *lb, const double *ub, int bpp, int tr,
                             const double *params, int nparam, int64_t nprob, const int *idx, int nsens,
                             double reltol, double abstol, int64_t maxiters, int initdiv, int normkind,
                             double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (!sens_args_ok(fam, dim, nparam, idx, nsens)) return -1;
    if (nthreads < 1) nthreads = 1;
    const int nc = 1 + nsens;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        const double *l = bpp ? lb + k * dim : lb, *uu = bpp ? ub + k * dim : ub;
        orc_sens_user u = {fam, dim, nparam, tr, nsens, idx, params + (size_t)k * nparam, l, uu};
        double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
        for (int i = 0; i < dim; i++) orc_u2t(l[i], uu[i], &tl[i], &tu[i]);
        int64_t ne = 0;
        int st = orc_hcubature_vec(orc_sens_eval, &u, nc, dim, tl, tu, reltol, abstol, maxiters, initdiv, normkind,
                                   I + (size_t)k * nc, &E[k], &ne);
        if (nevals) nevals[k] = ne;
        if (status) status[k] = st;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Round-synchronous variant of the HCubature loop (NOT in the reference): the checker for
 * ib200_hcubature_single (csrc/hcub_single.cu), which replaces the strictly sequential "pop the
 * worst box" order of HCubature.hcubature (call site src/Integrals.jl:386-393) by rounds --
 * every box whose error is >= half the current upper bound of the largest box error is bisected
 * in the same round; the threshold is lowered further to the exponent-histogram bound below which the
 * reference would provably not refine either.  Same rule (gm_rule), same split axis, same stopping test, same final
 * re-sum; only the ORDER in which boxes are refined differs, so (I, E) agree with orc_hcubature
 * within the requested tolerance, and with the CUDA path to summation-order rounding.
 * m0: the domain is pre-bisected m0 times (axis = level mod dim), as the multi-GPU driver does
 * before dealing sub-boxes to the ranks.  max_boxes: budget of rule applications.
 * ------------------------------------------------------------------------------------------ */
double orc_round_fraction = 0.5;
/* level-2 sub-problems of orc_hcubature_two_level: boxes a sub-problem's store may hold (0 = unbounded) -- the arena of
   csrc/hcubature_rb.cuh (option single_sub_boxes), with converged-box retirement when it fills; orc_last_retired = boxes retired by the last solve */
int64_t orc_sub_arena = 0;
int64_t orc_last_retired = 0;
double orc_retire_fraction = 0.25;
/* the rounds themselves, from a store of `len` unevaluated boxes (all of them in the todo list) */
static int hcub_rounds_core2(orc_integrand f, void *user, int dim, orc_box *t, int64_t cap, int64_t len,
                             double rtol, double atol, int64_t max_boxes, int first_box_only, int64_t switch_boxes, double switch_factor,
                             int64_t arena, int64_t *nretired_out,
                             double *Iout, double *Eout, int64_t *napplied_out, int64_t *nboxes_out, int64_t *rounds_out,
                             orc_box **store_out);
static int hcub_rounds_core(orc_integrand f, void *user, int dim, orc_box *t, int64_t cap, int64_t len,
                            double rtol, double atol, int64_t max_boxes, int first_box_only,
                            double *Iout, double *Eout, int64_t *napplied_out, int64_t *nboxes_out, int64_t *rounds_out)
{
    return hcub_rounds_core2(f, user, dim, t, cap, len, rtol, atol, max_boxes, first_box_only, 0, 0.0, 0, NULL, Iout, Eout, napplied_out, nboxes_out,
                             rounds_out, NULL);
}
/* switch_boxes > 0: the two-level solve (orc_hcubature_two_level) -- stop with status 3 ("hand over") after the stopping test of a round
   once the store holds >= switch_boxes boxes AND E <= switch_factor * tol (or the store holds >= 64 * switch_boxes boxes), and return the
   store (unsplit, every box evaluated) through store_out instead of freeing it.
   arena > 0: the store may hold at most `arena` boxes (a level-2 sub-problem's arena, csrc/hcubature_rb.cuh).  When the boxes selected by
   a round would not fit, the converged boxes are RETIRED first -- every box whose error lies below theta_r leaves the store, its integral
   and error stay in the sums -- and the round is decided again; theta_r is the upper edge of the last occupied exponent bin up to which the
   bins' counts x upper edges sum to <= orc_retire_fraction of the tolerance still unspent (tol - retired error), and never above the split
   threshold.  The second selection is capped by the room left (whole bins from the top, then the first boxes of the next bin in store
   order).  If less than 1/64 of the arena is free after a retirement, or nothing can be selected, the solve ends with status 1 (arena full). */
static int hcub_rounds_core2(orc_integrand f, void *user, int dim, orc_box *t, int64_t cap, int64_t len,
                             double rtol, double atol, int64_t max_boxes, int first_box_only, int64_t switch_boxes, double switch_factor,
                             int64_t arena, int64_t *nretired_out,
                             double *Iout, double *Eout, int64_t *napplied_out, int64_t *nboxes_out, int64_t *rounds_out,
                             orc_box **store_out)
{
    double Iret = 0.0, Eret = 0.0;
    int64_t nretired = 0;
    int64_t ntodo = 0, applied = 0, rounds = 0;
    int64_t *todo = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    char *sel = (char *)malloc((size_t)cap);
    for (int64_t q = 0; q < len; q++) todo[ntodo++] = q;
    double I = 0.0, E = 0.0, theta = 0.0;
    int status = 0;
    for (;;) {
        double mx = 0.0;
        for (int64_t j = 0; j < ntodo; j++) {
            orc_box *bx = &t[todo[j]];
            orc_cubrule(f, user, dim, bx->a, bx->b, &bx->I, &bx->E, &bx->kdiv);   /* Genz-Malik, or GK(7,15) in 1-D */
            I += bx->I; E += bx->E; if (bx->E > mx) mx = bx->E;
        }
        applied += ntodo; rounds++; ntodo = 0;
        if (first_box_only) break;                       /* degenerate domain: HCubature returns after the first box */
        if (E <= fmax(rtol * fabs(I), atol)) break;
        if (!isfinite(E)) { status = 2; break; }
        if (applied >= max_boxes) { status = 1; break; }
        if (switch_boxes > 0 && len >= switch_boxes &&
            (E <= switch_factor * fmax(rtol * fabs(I), atol) || len >= 64 * switch_boxes)) { status = 3; break; }
        const double theta_in = theta;
        int64_t nsplit = 0;
        int overflow = 0;
        for (int attempt = 0; ; attempt++) {
        theta = 0.5 * fmax(mx, theta_in);
        const double tol = fmax(rtol * fabs(I), atol);
        double cnt[2048];
        {   /* threshold from the exponent histogram (see the header comment of csrc/hcub_single.cu) */
            for (int b2 = 0; b2 < 2048; b2++) cnt[b2] = 0.0;
            for (int64_t i = 0; i < len; i++) { uint64_t bits; memcpy(&bits, &t[i].E, 8); cnt[(bits >> 52) & 0x7ff] += 1.0; }
            int bh = -1; double low = 0.0;
            for (int b2 = 0; b2 < 2047; b2++) {
                if (cnt[b2] != 0.0) { low += cnt[b2] * (b2 == 0 ? 0.0 : scalbn(1.0, b2 - 1023)); if (!(low <= tol - Eret)) break; }
                bh = b2;
            }
            const double remaining = (double)max_boxes - (double)applied;
            int bb = 2047; double above = 0.0;
            for (int b2 = 2046; b2 >= 0; b2--) { above += cnt[b2]; if (2.0 * above > remaining) break; bb = b2; }
            /* (3) selectivity: at most ORC_ROUND_FRACTION of the stored boxes per round (the largest errors first) */
            int bf = 2047; double top = 0.0;
            for (int b2 = 2046; b2 >= 0; b2--) { top += cnt[b2]; if (top > orc_round_fraction * (double)len) break; bf = b2; }
            if (bf > bb) bb = bf;
            const int bs = (bh + 1 > bb) ? bh + 1 : bb;
            const double theta_h = bs <= 0 ? 0.0 : (bs >= 2047 ? 1.7976931348623157e308 : scalbn(1.0, bs - 1023));
            theta = fmin(theta, theta_h);
        }
        /* (4) second attempt (after a retirement): the selection must fit the room left in the arena.  Bins are taken whole from the top
           while they fit; of the bin that does not fit any more, the first `quota` boxes in store order are taken */
        double th_hi = theta, th_lo = theta;
        int64_t quota = 0;
        if (attempt > 0) {
            const int64_t room = arena - len;
            if (room < arena / 64) { overflow = 1; break; }          /* not worth a round: the arena stays full */
            double up = 0.0;
            for (int b2 = 2046; b2 >= 0; b2--) {
                if (cnt[b2] == 0.0) continue;
                if (up + cnt[b2] > (double)room) {
                    th_hi = fmax(theta, b2 + 1 >= 2047 ? 1.7976931348623157e308 : scalbn(1.0, b2 + 1 - 1023));
                    th_lo = fmax(theta, b2 == 0 ? 0.0 : scalbn(1.0, b2 - 1023));
                    quota = room - (int64_t)up;
                    break;
                }
                up += cnt[b2];
            }
        }
        nsplit = 0;
        {
            int64_t nlo = 0;
            for (int64_t i = 0; i < len; i++) {
                int take = t[i].E >= th_hi;
                if (!take && t[i].E >= th_lo) { take = nlo < quota; nlo++; }
                sel[i] = (char)take;
                nsplit += take;
            }
        }
        if (attempt > 0 && nsplit == 0) { overflow = 1; break; }
        if (arena <= 0 || len + nsplit <= arena) break;
        if (attempt > 0) { overflow = 1; break; }
        {   /* retirement (see above); the histogram without the selected boxes, as the kernel's selection pass leaves it */
            for (int64_t i = 0; i < len; i++) if (sel[i]) { uint64_t bits; memcpy(&bits, &t[i].E, 8); cnt[(bits >> 52) & 0x7ff] -= 1.0; }
            const double R = orc_retire_fraction * (tol - Eret);
            int br = -1; double acc = 0.0;
            for (int b2 = 0; b2 <= 2045; b2++) {
                if (cnt[b2] == 0.0) continue;
                acc += cnt[b2] * scalbn(1.0, b2 + 1 - 1023);
                if (!(acc <= R)) break;
                br = b2;
            }
            const double theta_r = fmin(br < 0 ? 0.0 : scalbn(1.0, br + 1 - 1023), theta);
            int64_t kept = 0;
            for (int64_t i = 0; i < len; i++) {
                if (t[i].E >= theta_r) { if (kept != i) t[kept] = t[i]; kept++; }
                else { Iret += t[i].I; Eret += t[i].E; }
            }
            nretired += len - kept;
            len = kept;
        }
        }
        if (overflow) { status = 1; break; }
        if (len + nsplit > cap) {
            cap = 2 * (len + nsplit);
            t = (orc_box *)realloc(t, sizeof(orc_box) * (size_t)cap);
            todo = (int64_t *)realloc(todo, sizeof(int64_t) * (size_t)cap);
            sel = (char *)realloc(sel, (size_t)cap);
        }
        const int64_t n0 = len;
        for (int64_t i = 0; i < n0; i++) {
            if (!sel[i]) continue;
            orc_box box = t[i];
            I -= box.I; E -= box.E;
            int k = box.kdiv;
            double w = (box.b[k] - box.a[k]) * 0.5;
            orc_box b1 = box, b2 = box;
            b1.a[k] = box.a[k] + w; b1.I = 0.0; b1.E = 0.0;
            b2.b[k] = box.b[k] - w; b2.I = 0.0; b2.E = 0.0;
            t[i] = b1; t[len] = b2;
            todo[ntodo++] = i; todo[ntodo++] = len; len++;
        }
    }
    I = 0.0; E = 0.0;
    for (int64_t i = 0; i < len; i++) { I += t[i].I; E += t[i].E; }
    I += Iret; E += Eret;
    if (nretired_out) *nretired_out = nretired;
    free(todo); free(sel);
    if (store_out) *store_out = t; else free(t);
    *Iout = I; *Eout = E;
    if (napplied_out) *napplied_out = applied;
    if (nboxes_out) *nboxes_out = len;
    if (rounds_out) *rounds_out = rounds;
    return status;
}

/* ------------------------------------------------------------------------------------------
 * Two-level variant (NOT in the reference): the checker of ib200_hcubature_single for solves whose box store would outgrow the
 * device.  Level 1 = the rounds above until the store holds switch_boxes boxes and the error is within switch_factor of the
 * tolerance.  Level 2 = every stored box becomes an independent sub-problem, refined by the same rounds until ITS error is
 * below ITS share of the tolerance, eps_i = (1 - 2^-6) tol * E_i^alpha / sum_j E_j^alpha (tol = max(rtol |I|, atol) with the
 * level-1 estimate I); a sub-problem's boxes are summed and dropped when it is done, so memory is bounded by the level-1 store plus
 * one sub-problem per worker, and the sub-problems need no communication (multi-GPU: dealt to the ranks, one sum at the end).
 * sum eps_i < tol, so when every sub-problem converges E <= tol: HCubature's stopping test holds for the whole integral.
 * Same rule, split axis and per-box error as HCubature; what differs from the reference is, again, only the ORDER of refinement.
 * sub_max_boxes: budget of rule applications per sub-problem (the analogue of a full arena: status 1).
 * ------------------------------------------------------------------------------------------ */
int orc_hcubature_two_level(orc_integrand f, void *user, int dim, const double *a, const double *b,
                            double rtol, double atol, int64_t max_boxes, int64_t switch_boxes, double switch_factor,
                            int64_t sub_max_boxes, double alpha,
                            double *Iout, double *Eout, int64_t *napplied_out, int64_t *nboxes_out, int64_t *rounds_out, int64_t *nsub_out)
{
    if (dim < 1 || dim > ORC_MAXDIM || rtol < 0 || atol < 0 || max_boxes < 1 || switch_boxes < 1) return -1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    int64_t cap = 1024, applied = 0, len = 0, rounds = 0;
    orc_box *t = (orc_box *)malloc(sizeof(orc_box) * (size_t)cap), *store = NULL;
    for (int i = 0; i < dim; i++) { t[0].a[i] = a[i]; t[0].b[i] = b[i]; }
    t[0].I = 0.0; t[0].E = 0.0; t[0].kdiv = 0;
    double I = 0.0, E = 0.0;
    int st = hcub_rounds_core2(f, user, dim, t, cap, 1, rtol, atol, max_boxes, 0, switch_boxes, switch_factor, 0, NULL, &I, &E, &applied, &len, &rounds, &store);
    orc_last_retired = 0;
    if (nsub_out) *nsub_out = 0;
    if (st != 3) {
        free(store);
        *Iout = I; *Eout = E;
        if (napplied_out) *napplied_out = applied;
        if (nboxes_out) *nboxes_out = len;
        if (rounds_out) *rounds_out = rounds;
        return st;
    }
    const double tol = fmax(rtol * fabs(I), atol) * (1.0 - 0.015625);
    double wsum = 0.0;
    for (int64_t i = 0; i < len; i++) wsum += pow(store[i].E, alpha);
    double Is = 0.0, Es = 0.0;
    int status = 0;
    int64_t nsub = 0, boxes = 0;
    for (int64_t i = 0; i < len; i++) {
        const double eps = tol * (pow(store[i].E, alpha) / wsum);
        if (store[i].E <= eps) { Is += store[i].I; Es += store[i].E; boxes++; continue; }
        orc_box *tt = (orc_box *)malloc(sizeof(orc_box) * 1024);
        tt[0] = store[i];
        /* the box is already evaluated: its first round is its own bisection (err >= theta with theta = its error) -- restated
           by starting the core on the two halves */
        {
            orc_box box = store[i];
            int k = box.kdiv;
            double w = (box.b[k] - box.a[k]) * 0.5;
            orc_box b1 = box, b2 = box;
            b1.a[k] = box.a[k] + w; b1.I = 0.0; b1.E = 0.0;
            b2.b[k] = box.b[k] - w; b2.I = 0.0; b2.E = 0.0;
            tt[0] = b1; tt[1] = b2;
        }
        double Ii = 0.0, Ei = 0.0; int64_t ap = 0, nb = 0, nr = 0;
        int64_t nret = 0;
        int s2 = hcub_rounds_core2(f, user, dim, tt, 1024, 2, 0.0, eps, sub_max_boxes, 0, 0, 0.0, orc_sub_arena, &nret, &Ii, &Ei, &ap, &nb, &nr, NULL);
        orc_last_retired += nret;
        if (s2 != 0) status = s2 == 2 ? 2 : (status == 2 ? 2 : 1);
        Is += Ii; Es += Ei; applied += ap; boxes += nb; nsub++;
    }
    free(store);
    *Iout = Is; *Eout = Es;
    if (napplied_out) *napplied_out = applied;
    if (nboxes_out) *nboxes_out = boxes;
    if (rounds_out) *rounds_out = rounds;
    if (nsub_out) *nsub_out = nsub;
    return status;
}

int orc_hcubature_two_level_family(int fam, int dim, const double *lb, const double *ub, int tr, const double *params,
                                   double reltol, double abstol, int64_t max_boxes, int64_t switch_boxes, double switch_factor,
                                   int64_t sub_max_boxes, double alpha,
                                   double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds, int64_t *nsub)
{
    orc_family_user u = {fam, dim, params};
    orc_cov c = {orc_family_integrand, &u, dim, tr, lb, ub};
    double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
    for (int i = 0; i < dim; i++) orc_u2t(lb[i], ub[i], &tl[i], &tu[i]);
    return orc_hcubature_two_level(orc_cov_eval, &c, dim, tl, tu, reltol, abstol, max_boxes, switch_boxes, switch_factor, sub_max_boxes, alpha,
                                   I, E, napplied, nboxes, rounds, nsub);
}

int orc_hcubature_rounds(orc_integrand f, void *user, int dim, const double *a, const double *b,
                         double rtol, double atol, int64_t max_boxes, int m0,
                         double *Iout, double *Eout, int64_t *napplied_out, int64_t *nboxes_out, int64_t *rounds_out)
{
    if (dim < 1 || dim > ORC_MAXDIM || rtol < 0 || atol < 0 || max_boxes < 1 || m0 < 0 || m0 > 30) return -1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    int64_t cap = (1LL << m0) + 1024, len = 0;
    orc_box *t = (orc_box *)malloc(sizeof(orc_box) * (size_t)cap);
    for (int64_t q = 0; q < (1LL << m0); q++) {
        orc_box bx;
        for (int i = 0; i < dim; i++) { bx.a[i] = a[i]; bx.b[i] = b[i]; }
        for (int lev = 0; lev < m0; lev++) {
            int k = lev % dim;
            double w = (bx.b[k] - bx.a[k]) * 0.5;
            if ((q >> lev) & 1) bx.a[k] = bx.a[k] + w; else bx.b[k] = bx.b[k] - w;
        }
        bx.I = 0.0; bx.E = 0.0; bx.kdiv = 0;
        t[len++] = bx;
    }
    return hcub_rounds_core(f, user, dim, t, cap, len, rtol, atol, max_boxes, 0, Iout, Eout, napplied_out, nboxes_out, rounds_out);
}

/* The same rounds as HCubature.hcubature's drop-in for a BATCH (checker of ib200_hcubature_batch, mode IB200_HCUB_ROUNDS,
 * csrc/hcubature_rb.cuh): the initial store is HCubature's initdiv grid (CartesianIndices order, first index fastest, the
 * box bounds of orc_hcubature), a degenerate domain returns after the first box, and maxevals is honoured as a budget of
 * ceil(maxevals / points per rule) rule applications.  nevals = rule applications * points per rule. */
int orc_hcubature_rounds_init(orc_integrand f, void *user, int dim, const double *a, const double *b,
                              double rtol, double atol, int64_t maxevals, int initdiv,
                              double *Iout, double *Eout, int64_t *nevals_out, int64_t *nboxes_out, int64_t *rounds_out)
{
    if (dim < 1 || dim > ORC_MAXDIM || initdiv < 1 || rtol < 0 || atol < 0 || maxevals < 0) return -1;
    if (rtol == 0 && atol == 0) rtol = sqrt(2.220446049250313e-16);
    const int64_t np = dim == 1 ? 15 : 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim);
    const int64_t max_apps = maxevals < (1LL << 60) ? (maxevals + np - 1) / np : (1LL << 60);
    double Delta[ORC_MAXDIM], prodD = 1.0;
    int64_t total = 1;
    for (int i = 0; i < dim; i++) { Delta[i] = (b[i] - a[i]) / initdiv; prodD *= Delta[i]; total *= initdiv; }
    const int degenerate = (prodD == 0.0);
    if (degenerate) total = 1;
    int64_t cap = total + 1024, len = 0;
    orc_box *t = (orc_box *)malloc(sizeof(orc_box) * (size_t)cap);
    int c[ORC_MAXDIM] = {0};
    for (int64_t q = 0; q < total; q++) {
        orc_box bx;
        for (int i = 0; i < dim; i++) {
            bx.a[i] = a[i] + c[i] * Delta[i];
            bx.b[i] = (c[i] == initdiv - 1) ? b[i] : bx.a[i] + Delta[i];
        }
        bx.I = 0.0; bx.E = 0.0; bx.kdiv = 0;
        t[len++] = bx;
        int k = 0;
        while (k < dim && ++c[k] == initdiv) { c[k] = 0; k++; }
    }
    int64_t applied = 0;
    int st = hcub_rounds_core(f, user, dim, t, cap, len, rtol, atol, max_apps, degenerate, Iout, Eout, &applied, nboxes_out, rounds_out);
    if (nevals_out) *nevals_out = applied * np;
    return st;
}

int orc_solve_hcubature_rounds(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                               double reltol, double abstol, int64_t max_boxes, int m0,
                               double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds)
{
    orc_cov c = {f, user, dim, tr, lb, ub};
    double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
    for (int i = 0; i < dim; i++) orc_u2t(lb[i], ub[i], &tl[i], &tu[i]);
    return orc_hcubature_rounds(orc_cov_eval, &c, dim, tl, tu, reltol, abstol, max_boxes, m0, I, E, napplied, nboxes, rounds);
}

int orc_hcubature_rounds_family(int fam, int dim, const double *lb, const double *ub, int tr, const double *params,
                                double reltol, double abstol, int64_t max_boxes, int m0,
                                double *I, double *E, int64_t *napplied, int64_t *nboxes, int64_t *rounds)
{
    orc_family_user u = {fam, dim, params};
    return orc_solve_hcubature_rounds(orc_family_integrand, &u, dim, lb, ub, tr, reltol, abstol, max_boxes, m0,
                                      I, E, napplied, nboxes, rounds);
}

/* ------------------------------------------------------------------------------------------
 * Integrals.jl-level solves: init wraps every tuple-domain alg in
 * ChangeOfVariables(transformation_if_inf, alg) (src/common.jl:97-103); __solve maps the
 * domain with u2t and the integrand with t2ujac (src/Integrals.jl:192-224,
 * src/infinity_handling.jl:165-170) and re-dispatches to the back-end.
 * ------------------------------------------------------------------------------------------ */
int orc_solve_quadgk(orc_integrand f, void *user, double lb, double ub, int tr,
                     double reltol, double abstol, int64_t maxiters,
                     double *I, double *E, int64_t *nevals)
{
    orc_cov c = {f, user, 1, tr, &lb, &ub};
    double tl, tu;
    orc_u2t(lb, ub, &tl, &tu);
    return orc_quadgk(orc_cov_eval, &c, tl, tu, reltol, abstol, maxiters, I, E, nevals, NULL);
}

int orc_solve_hcubature(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                        double reltol, double abstol, int64_t maxiters, int initdiv,
                        double *I, double *E, int64_t *nevals)
{
    orc_cov c = {f, user, dim, tr, lb, ub};
    double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
    for (int i = 0; i < dim; i++) orc_u2t(lb[i], ub[i], &tl[i], &tu[i]);
    return orc_hcubature(orc_cov_eval, &c, dim, tl, tu, reltol, abstol, maxiters, initdiv, I, E, nevals, NULL);
}

double orc_solve_quadrule(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                          const double *nodes, const double *weights, int64_t N)
{
    orc_cov c = {f, user, dim, tr, lb, ub};
    double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
    for (int i = 0; i < dim; i++) orc_u2t(lb[i], ub[i], &tl[i], &tu[i]);
    return orc_evalrule(orc_cov_eval, &c, dim, tl, tu, nodes, weights, N);
}

double orc_solve_quadrule_tensor(orc_integrand f, void *user, int dim, const double *lb, const double *ub, int tr,
                                 const double *x1d, const double *w1d, int n1d)
{
    orc_cov c = {f, user, dim, tr, lb, ub};
    double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
    for (int i = 0; i < dim; i++) orc_u2t(lb[i], ub[i], &tl[i], &tu[i]);
    return orc_evalrule_tensor(orc_cov_eval, &c, dim, tl, tu, x1d, w1d, n1d);
}

double orc_solve_gauss_legendre(orc_integrand f, void *user, double lb, double ub, int tr,
                                const double *nodes, const double *weights, int n, int subintervals)
{
    orc_cov c = {f, user, 1, tr, &lb, &ub};
    double tl, tu;
    orc_u2t(lb, ub, &tl, &tu);
    /* ext/IntegralsFastGaussQuadratureExt.jl:47-71: composite iff the alg was built with subintervals>1 */
    if (subintervals > 1) return orc_composite_gauss_legendre(orc_cov_eval, &c, tl, tu, nodes, weights, n, subintervals);
    return orc_gauss_legendre(orc_cov_eval, &c, tl, tu, nodes, weights, n);
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers (CPU baseline): a threaded loop of independent solve() calls, the stand-in
 * for `Threads.@threads for k: solve(remake(prob, p = ps[:,k]), alg)`.
 * ------------------------------------------------------------------------------------------ */
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_quadgk_batch(int fam, const double *lb, const double *ub, int bpp, int tr,
                     const double *params, int nparam, int64_t nprob,
                     double reltol, double abstol, int64_t maxiters,
                     double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (orc_family_nparam(fam, 1) != nparam) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, 1, params + k * nparam};
        double l = bpp ? lb[k] : lb[0], u = bpp ? ub[k] : ub[0];
        int64_t ne = 0;
        int st = orc_solve_quadgk(orc_family_integrand, &fu, l, u, tr, reltol, abstol, maxiters, &I[k], &E[k], &ne);
        if (nevals) nevals[k] = ne;
        if (status) status[k] = st;
    }
    return 0;
}

int orc_hcubature_batch(int fam, int dim, const double *lb, const double *ub, int bpp, int tr,
                        const double *params, int nparam, int64_t nprob,
                        double reltol, double abstol, int64_t maxiters, int initdiv,
                        double *I, double *E, int64_t *nevals, int32_t *status, int nthreads)
{
    if (orc_family_nparam(fam, dim) != nparam) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, dim, params + k * nparam};
        const double *l = bpp ? lb + k * dim : lb, *u = bpp ? ub + k * dim : ub;
        int64_t ne = 0;
        int st = orc_solve_hcubature(orc_family_integrand, &fu, dim, l, u, tr, reltol, abstol, maxiters, initdiv, &I[k], &E[k], &ne);
        if (nevals) nevals[k] = ne;
        if (status) status[k] = st;
    }
    return 0;
}

int orc_hcubature_rounds_batch(int fam, int dim, const double *lb, const double *ub, int bpp, int tr,
                               const double *params, int nparam, int64_t nprob,
                               double reltol, double abstol, int64_t maxiters, int initdiv,
                               double *I, double *E, int64_t *nevals, int32_t *status, int64_t *rounds, int nthreads)
{
    if (orc_family_nparam(fam, dim) != nparam) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, dim, params + k * nparam};
        const double *l = bpp ? lb + k * dim : lb, *u = bpp ? ub + k * dim : ub;
        orc_cov c = {orc_family_integrand, &fu, dim, tr, l, u};
        double tl[ORC_MAXDIM], tu[ORC_MAXDIM];
        for (int i = 0; i < dim; i++) orc_u2t(l[i], u[i], &tl[i], &tu[i]);
        int64_t ne = 0, nr = 0;
        int st = orc_hcubature_rounds_init(orc_cov_eval, &c, dim, tl, tu, reltol, abstol, maxiters, initdiv, &I[k], &E[k], &ne, NULL, &nr);
        if (nevals) nevals[k] = ne;
        if (status) status[k] = st;
        if (rounds) rounds[k] = nr;
    }
    return 0;
}

int orc_fixed_tensor_batch(int fam, int dim, const double *lb, const double *ub, int bpp, int tr,
                           const double *params, int nparam, int64_t nprob,
                           const double *x1d, const double *w1d, int n1d, double *I, int nthreads)
{
    if (orc_family_nparam(fam, dim) != nparam) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, dim, params + k * nparam};
        const double *l = bpp ? lb + k * dim : lb, *u = bpp ? ub + k * dim : ub;
        I[k] = orc_solve_quadrule_tensor(orc_family_integrand, &fu, dim, l, u, tr, x1d, w1d, n1d);
    }
    return 0;
}

int orc_fixed_nodes_batch(int fam, int dim, const double *lb, const double *ub, int bpp, int tr,
                          const double *params, int nparam, int64_t nprob,
                          const double *nodes, const double *weights, int64_t N, double *I, int nthreads)
{
    if (orc_family_nparam(fam, dim) != nparam) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, dim, params + k * nparam};
        const double *l = bpp ? lb + k * dim : lb, *u = bpp ? ub + k * dim : ub;
        I[k] = orc_solve_quadrule(orc_family_integrand, &fu, dim, l, u, tr, nodes, weights, N);
    }
    return 0;
}

int orc_gauss_legendre_batch(int fam, const double *lb, const double *ub, int bpp, int tr,
                             const double *params, int nparam, int64_t nprob,
                             const double *nodes, const double *weights, int n, int subintervals,
                             double *I, int nthreads)
{
    if (orc_family_nparam(fam, 1) != nparam) return -1;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
    for (int64_t k = 0; k < nprob; k++) {
        orc_family_user fu = {fam, 1, params + k * nparam};
        double l = bpp ? lb[k] : lb[0], u = bpp ? ub[k] : ub[0];
        I[k] = orc_solve_gauss_legendre(orc_family_integrand, &fu, l, u, tr, nodes, weights, n, subintervals);
    }
    return 0;
}
