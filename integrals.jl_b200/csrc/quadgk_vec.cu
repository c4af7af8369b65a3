// quadgk_vec.cu -- batched adaptive Gauss-Kronrod (7,15) quadrature of VECTOR-valued integrands, one warp per problem.
//
// SURVEY 8(f) rank 1.  Replaces QuadGK.quadgk!(f!, result, lb, ub; rtol, atol, maxevals, order = 7, norm) as called from
// __solvebp_call(::QuadGKJL) for array-valued integrands (src/Integrals.jl:316-326; QuadGKJL.norm, src/algorithms.jl:10-11,56),
// behind the automatic ChangeOfVariables (src/common.jl:97-103).  Component c of problem k is the registered family
// evaluated with parameter column params[:, c, k]; the ncomp components share the nodes and the subdivision, the error
// of a segment is norm(K - G) over the components, and the refinement stops when E <= max(atol, rtol * norm(I)).
// Oracle: orc_quadgk_vec.
//
// Mapping: as quadgk.cu (lanes 0-14 / 16-30 evaluate the 15 Kronrod nodes of the two half-segments; segment pool
// {E, a, b, I[ncomp]} in shared memory spilling to a per-warp global arena; REDUX arg-max).  A lane maps its node to u
// once and evaluates every component there; the Kronrod and Gauss sums of each component are formed with xor-shuffles.
#include "common.cuh"
#include "family.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

__constant__ double v_gk_x[8] = {
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.0};
__constant__ double v_gk_w[8] = {
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double v_gk_wg[4] = {
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

// segment s: E, a, b at e[s], a[s], b[s]; component c of its integral at I[c * cap + s] (shared part) / gI[c * capG + g]
template <int NC>
struct VecPool {
    double *sE, *sA, *sB, *sI;   // shared, capacity capS
    double *gE, *gA, *gB, *gI;   // global spill, capacity capG
    int capS, capG;
    __device__ __forceinline__ double getE(int s) const { return s < capS ? sE[s] : gE[s - capS]; }
    __device__ __forceinline__ void load(int s, int nc, double &a, double &b, double &E, double (&I)[NC]) const {
        if (s < capS) { a = sA[s]; b = sB[s]; E = sE[s];
#pragma unroll
            for (int c = 0; c < NC; c++) if (c < nc) I[c] = sI[c * capS + s]; }
        else { const int g = s - capS; a = gA[g]; b = gB[g]; E = gE[g];
#pragma unroll
            for (int c = 0; c < NC; c++) if (c < nc) I[c] = gI[(size_t)c * capG + g]; }
    }
    __device__ __forceinline__ void store(int s, int nc, double a, double b, double E, const double (&I)[NC]) const {
        if (s < capS) { sA[s] = a; sB[s] = b; sE[s] = E;
#pragma unroll
            for (int c = 0; c < NC; c++) if (c < nc) sI[c * capS + s] = I[c]; }
        else { const int g = s - capS; gA[g] = a; gB[g] = b; gE[g] = E;
#pragma unroll
            for (int c = 0; c < NC; c++) if (c < nc) gI[(size_t)c * capG + g] = I[c]; }
    }
};

__device__ __forceinline__ bool v_endpoint_roundoff(double a, double b) {
    const double c = 0.5 * (b - a);
    return (a == a + (1 + v_gk_x[0]) * c) || (b == a + (1 - v_gk_x[0]) * c);
}

template <int NC>
__device__ __forceinline__ double vnorm(const double (&v)[NC], int nc, int kind) {
    double r = 0.0;
    if (kind == 2) return fabs(v[0]);          // the first component alone (sensitivities: a ForwardDiff.Dual's norm is its value's)
    if (kind == 1) {
#pragma unroll
        for (int c = 0; c < NC; c++) if (c < nc) { const double t = fabs(v[c]); if (t > r || isnan(t)) r = t; }
        return r;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) if (c < nc) r = r + v[c] * v[c];
    return sqrt(r);
}

template <int FAM, bool ALLFIN, int NC>
__global__ void __launch_bounds__(kBlock, 1)
quadgk_vec_kernel(const double *__restrict__ params, int nparam, int ncomp, int64_t nprob,
                  const double *__restrict__ lb, const double *__restrict__ ub, int bpp, int tr,
                  double rtol, double atol, int64_t maxevals, int normkind, const ib::SensIdx sens, int capS, int capG, double *__restrict__ arena,
                  unsigned long long *__restrict__ counter,
                  double *__restrict__ out_I, double *__restrict__ out_E, int64_t *__restrict__ out_nevals,
                  int32_t *__restrict__ out_status) {
    using Fm = ib::Family<FAM, 1>;
    constexpr int NQ = Fm::NQ;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + wib;
    VecPool<NC> pool;
    pool.capS = capS; pool.capG = capG;
    double *wsm = sm + (size_t)wib * ((3 + NC) * capS + NC * NQ);
    pool.sE = wsm; pool.sA = pool.sE + capS; pool.sB = pool.sA + capS; pool.sI = pool.sB + capS;
    double *qs = pool.sI + (size_t)NC * capS;                        // prepared parameters of the components: [NC][NQ]
    double *ga = arena + (size_t)gwarp * (3 + NC) * capG;
    pool.gE = ga; pool.gA = pool.gE + capG; pool.gB = pool.gA + capG; pool.gI = pool.gB + capG;
    const int cap = capS + capG;

    const int j = lane & 15;
    const double fac = j < 7 ? 1 + v_gk_x[j] : (j < 14 ? 1 - v_gk_x[j - 7] : 1.0);
    const double wk = j < 7 ? v_gk_w[j] : (j == 14 ? v_gk_w[7] : 0.0);
    const double wg = (j < 7 && (j & 1)) ? v_gk_wg[j >> 1] : (j == 14 ? v_gk_wg[3] : 0.0);
    const bool eval_lane = j < 15;

    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;

    for (;;) {
        unsigned long long pidx = 0;
        if (lane == 0) pidx = atomicAdd(counter, 1ULL);
        pidx = __shfl_sync(0xffffffffu, pidx, 0);
        if ((int64_t)pidx >= nprob) break;
        const int64_t prob = (int64_t)pidx;

        // the change of variables is shared by the components; every component has its own prepared parameters
        const double l = bpp ? lb[prob] : lb[0], u_ = bpp ? ub[prob] : ub[0];
        ib::AxisMap ax;
        ib::make_axis(l, u_, ax);
        const double den = (u_ - l) * 0.5;
        const double tlo = ALLFIN ? -1.0 : ax.tlo, thi = ALLFIN ? 1.0 : ax.thi;
        __syncwarp();
        if (sens.n > 0) {          // sensitivities: ONE parameter column per problem, kept raw (ib::Grad differentiates with respect to it)
            for (int t = lane; t < nparam; t += 32) qs[t] = params[(size_t)prob * nparam + t];
        } else {
            for (int c = lane; c < ncomp; c += 32) {
                double q[NQ];
                Fm::prepare(params + ((size_t)prob * ncomp + c) * nparam, q);
#pragma unroll
                for (int t = 0; t < NQ; t++) qs[c * NQ + t] = q[t];
            }
        }
        __syncwarp();

        auto rule2 = [&](double a0, double b0, double a1, double b1, double (&I0)[NC], double &E0, double (&I1)[NC], double &E1) {
            const bool hi = lane >= 16;
            const double a = hi ? a1 : a0, b = hi ? b1 : b0;
            const double s = 0.5 * (b - a);
            double g[NC], d[NC], Is[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) g[c] = 0.0;
            if (eval_lane) {
                const double t = a + fac * s;
                double x[1], jac;
                if (ALLFIN) { x[0] = l + (1.0 + t) * den; jac = den; }
                else ib::t2ujac(tr, t, ax, x[0], jac);
                bool done = false;
                if constexpr (ib::Grad<FAM, 1>::ok) {
                    if (sens.n > 0) {
                        double go[1 + ib::fam_nparam(FAM, 1)];
                        ib::Grad<FAM, 1>::run(x, qs, go);
                        g[0] = go[0] * jac;
#pragma unroll
                        for (int c = 1; c < NC; c++) if (c < ncomp) g[c] = go[1 + sens.idx[c - 1]] * jac;
                        done = true;
                    }
                }
                if (!done) {
#pragma unroll
                    for (int c = 0; c < NC; c++) if (c < ncomp) {
                        double q[NQ];
#pragma unroll
                        for (int t2 = 0; t2 < NQ; t2++) q[t2] = qs[c * NQ + t2];
                        g[c] = Fm::eval(x, q) * jac;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if (c < ncomp) {
                    const double partner = __shfl_down_sync(0xffffffffu, g[c], 7);
                    const double v = j < 7 ? g[c] + partner : g[c];
                    double ck = v * wk, cg = v * wg;
#pragma unroll
                    for (int off = 8; off > 0; off >>= 1) {
                        ck += __shfl_xor_sync(0xffffffffu, ck, off);
                        cg += __shfl_xor_sync(0xffffffffu, cg, off);
                    }
                    Is[c] = ck * s; d[c] = ck * s - cg * s;
                } else { Is[c] = 0.0; d[c] = 0.0; }
            }
            const double Es = vnorm<NC>(d, ncomp, normkind);
#pragma unroll
            for (int c = 0; c < NC; c++) { I0[c] = __shfl_sync(0xffffffffu, Is[c], 0); I1[c] = __shfl_sync(0xffffffffu, Is[c], 16); }
            E0 = __shfl_sync(0xffffffffu, Es, 0); E1 = __shfl_sync(0xffffffffu, Es, 16);
        };

        double I[NC], Iu[NC], E, Eu;
        rule2(tlo, thi, tlo, thi, I, E, Iu, Eu);
        int64_t nev = 15;
        int nseg = 1;
        int status = IB200_ST_CONVERGED;
        if (lane == 0) pool.store(0, ncomp, tlo, thi, E, I);
        __syncwarp();
        if (!isfinite(E)) status = IB200_ST_NONFINITE;
        else if (!(E <= atol || E <= rtol * vnorm<NC>(I, ncomp, normkind))) {
            if (nev >= maxevals) status = IB200_ST_MAXEVALS;
            while (E > atol && E > rtol * vnorm<NC>(I, ncomp, normkind) && nev < maxevals) {
                unsigned long long best = 0ULL; int bslot = 0x7fffffff;
                for (int s = lane; s < nseg; s += 32) {
                    const unsigned long long k = (unsigned long long)__double_as_longlong(pool.getE(s));
                    if (k > best || (k == best && s < bslot)) { best = k; bslot = s; }
                }
                if (nseg <= lane) { best = 0ULL; bslot = 0x7fffffff; }
                const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
                const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                const bool win = (hi == mhi && lo == mlo && bslot != 0x7fffffff);
                const int slot = (int)__reduce_min_sync(0xffffffffu, win ? (unsigned)bslot : 0x7fffffffu);
                double sa, sb, sEr, sI[NC];
                pool.load(slot, ncomp, sa, sb, sEr, sI);
                const double mid = (sa + sb) / 2;
                if (v_endpoint_roundoff(sa, mid) || v_endpoint_roundoff(mid, sb)) { status = IB200_ST_ROUNDOFF; break; }
                if (nseg + 1 > cap) { status = IB200_ST_MAXEVALS; break; }
                double I1[NC], I2[NC], E1, E2;
                rule2(sa, mid, mid, sb, I1, E1, I2, E2);
#pragma unroll
                for (int c = 0; c < NC; c++) I[c] = (I[c] - sI[c]) + I1[c] + I2[c];
                E = (E - sEr) + E1 + E2;
                nev += 30;
                __syncwarp();
                if (lane == 0) { pool.store(slot, ncomp, sa, mid, E1, I1); pool.store(nseg, ncomp, mid, sb, E2, I2); }
                nseg++;
                __syncwarp();
                if (!isfinite(E1) || !isfinite(E2)) { status = IB200_ST_NONFINITE; break; }
            }
            // re-sum over all segments (QuadGK resum), fixed order: lane-strided then xor tree
            double pE = 0.0, pI[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) pI[c] = 0.0;
            for (int s = lane; s < nseg; s += 32) {
                double a_, b_, e_, i_[NC];
                pool.load(s, ncomp, a_, b_, e_, i_);
                pE += e_;
#pragma unroll
                for (int c = 0; c < NC; c++) if (c < ncomp) pI[c] += i_[c];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pE += __shfl_xor_sync(0xffffffffu, pE, off);
#pragma unroll
                for (int c = 0; c < NC; c++) pI[c] += __shfl_xor_sync(0xffffffffu, pI[c], off);
            }
#pragma unroll
            for (int c = 0; c < NC; c++) I[c] = pI[c];
            E = pE;
            if (status == IB200_ST_CONVERGED && !(E <= atol || E <= rtol * vnorm<NC>(I, ncomp, normkind)) && nev >= maxevals) status = IB200_ST_MAXEVALS;
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < NC; c++) if (c < ncomp) out_I[prob * ncomp + c] = I[c];
            if (out_E) out_E[prob] = E;
            if (out_nevals) out_nevals[prob] = nev;
            if (out_status) out_status[prob] = status;
        }
        __syncwarp();
    }
}

}  // namespace

static int32_t quadgk_vec_impl(ib200_ctx *ctx, int32_t family, int32_t transform,
                               const double *lb, const double *ub, int32_t bpp,
                               const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                               double reltol, double abstol, int64_t maxevals, int32_t normkind, const ib::SensIdx sens,
                               double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status);

extern "C" int32_t ib200_quadgk_vec_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                          const double *lb, const double *ub, int32_t bpp,
                                          const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                                          double reltol, double abstol, int64_t maxevals, int32_t normkind,
                                          double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_quadgk_vec_batch: ctx is NULL");
    IB_REQUIRE(ctx, normkind == 0 || normkind == 1, "ib200_quadgk_vec_batch: norm must be 0 (2-norm) or 1 (maximum norm)");
    ib::SensIdx none; none.n = 0;
    return quadgk_vec_impl(ctx, family, transform, lb, ub, bpp, params, nparam, ncomp, nprob, reltol, abstol, maxevals, normkind, none,
                           out_I, out_E, out_nevals, out_status);
}

// I and dI/dp as one vector-valued integral [f, df/dp_s...] (see ib::Grad, family.cuh)
extern "C" int32_t ib200_quadgk_sens_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                           const double *lb, const double *ub, int32_t bpp,
                                           const double *params, int32_t nparam, int64_t nprob,
                                           const int32_t *sens_idx, int32_t nsens,
                                           double reltol, double abstol, int64_t maxevals, int32_t normkind,
                                           double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_quadgk_sens_batch: ctx is NULL");
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, "ib200_quadgk_sens_batch: unknown integrand family");
    IB_REQUIRE(ctx, nsens >= 1 && nsens <= 7 && sens_idx != nullptr, "ib200_quadgk_sens_batch: 1 <= nsens <= 7 parameters to differentiate with respect to");
    IB_REQUIRE(ctx, normkind >= 0 && normkind <= 2, "ib200_quadgk_sens_batch: norm must be 0 (2-norm), 1 (maximum norm) or 2 (value only)");
    IB_REQUIRE(ctx, family != IB200_FAM_GENZ_DISC, "ib200_quadgk_sens_batch: the discontinuous family has no parameter sensitivities (its support moves with u)");
    ib::SensIdx si; si.n = nsens;
    for (int s = 0; s < 7; s++) {
        si.idx[s] = s < nsens ? sens_idx[s] : 0;
        IB_REQUIRE(ctx, si.idx[s] >= 0 && si.idx[s] < nparam, "ib200_quadgk_sens_batch: parameter index out of range");
    }
    return quadgk_vec_impl(ctx, family, transform, lb, ub, bpp, params, nparam, 1 + nsens, nprob, reltol, abstol, maxevals, normkind, si,
                           out_I, out_E, out_nevals, out_status);
}

static int32_t quadgk_vec_impl(ib200_ctx *ctx, int32_t family, int32_t transform,
                               const double *lb, const double *ub, int32_t bpp,
                               const double *params, int32_t nparam, int32_t ncomp, int64_t nprob,
                               double reltol, double abstol, int64_t maxevals, int32_t normkind, const ib::SensIdx sens,
                               double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_quadgk_vec_batch: ctx is NULL");
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, "ib200_quadgk_vec_batch: unknown integrand family");
    IB_REQUIRE(ctx, transform >= 0 && transform <= 2, "ib200_quadgk_vec_batch: unknown transform id");
    IB_REQUIRE(ctx, ib::fam_dim_ok(family, 1) && nparam == ib::fam_nparam(family, 1),
               "QuadGKB200 only accepts one-dimensional quadrature problems.");          // Integrals.jl:263-266
    IB_REQUIRE(ctx, ncomp >= 1 && ncomp <= 8, "ib200_quadgk_vec_batch: 1 <= ncomp <= 8 components");
    IB_REQUIRE(ctx, !(reltol < 0) && !(abstol < 0), "ib200_quadgk_vec_batch: negative tolerance");
    IB_REQUIRE(ctx, nprob >= 0, "ib200_quadgk_vec_batch: negative nprob");
    if (nprob == 0) return IB200_OK;
    IB_REQUIRE(ctx, lb && ub && params && out_I, "ib200_quadgk_vec_batch: NULL data pointer");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t nb = bpp ? (size_t)nprob : 1;
    const bool allfin = ib200_all_finite(ctx, lb, ub, nb);
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

    const int NC = ncomp <= 4 ? 4 : 8;
    const int NQmax = 4;                                        // largest Family<FAM,1>::NQ (gausspoly)
    size_t smem_budget = (size_t)(200 * 1024);
    if (smem_budget > ctx->smem_optin) smem_budget = ctx->smem_optin;
    int capS = (int)((smem_budget / (kWarps * sizeof(double)) - (size_t)NC * NQmax) / (3 + NC));
    capS &= ~7;
    if (capS > 1024) capS = 1024;
    if (capS < 32) capS = 32;
    int64_t need = maxevals >= (int64_t)1 << 40 ? (int64_t)1 << 40 : 1 + (maxevals > 15 ? (maxevals - 15 + 29) / 30 : 0);
    int64_t capG = need - capS;
    if (capG < 8) capG = 8;
    if (capG > 8192) capG = 8192;
    int64_t blocks = (nprob + kWarps - 1) / kWarps;
    if (blocks > (int64_t)ctx->sm_count) blocks = ctx->sm_count;
    double *arena = (double *)ctx->get(SL_WORK0, (size_t)blocks * kWarps * (3 + NC) * (size_t)capG * sizeof(double));
    unsigned long long *counter = (unsigned long long *)ctx->get(SL_WORK1, sizeof(unsigned long long));
    if (!arena || !counter) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_quadgk_vec_batch: scratch allocation failed");
    IB_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));

    bool launched = false;
    cudaError_t attr_err = cudaSuccess;
    ib200_time_begin(ctx);
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, 1)) {
            auto go = [&](auto kern, int nc_t) {
                const size_t smem = (size_t)kWarps * ((size_t)(3 + nc_t) * capS + (size_t)nc_t * ib::Family<FAM, 1>::NQ) * sizeof(double);
                attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(dpar.d, nparam, ncomp, nprob, dlb.d, dub.d, bpp, transform,
                                                                     reltol, abstol, maxevals, normkind, sens, capS, (int)capG, arena, counter,
                                                                     oI.d, oE.d, oN.d, oS.d);
                launched = true;
            };
            if (NC == 4) { if (allfin) go(quadgk_vec_kernel<FAM, true, 4>, 4); else go(quadgk_vec_kernel<FAM, false, 4>, 4); }
            else         { if (allfin) go(quadgk_vec_kernel<FAM, true, 8>, 8); else go(quadgk_vec_kernel<FAM, false, 8>, 8); }
        }
    });
    if (attr_err != cudaSuccess) { cudaGetLastError(); return ib200_fail(ctx, IB200_ERR_CUDA, "cudaFuncSetAttribute failed"); }
    if (!launched) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_quadgk_vec_batch: family not compiled in for dim 1");
    IB_LAUNCH_CHECK(ctx);
    ib200_time_end(ctx);
    if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
    if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}
