// quadgk.cu -- batched adaptive Gauss-Kronrod (7,15) quadrature, one warp per problem.
//
// Replaces QuadGK.quadgk(f, lb, ub; rtol, atol, maxevals, order = 7) as called from
// __solvebp_call(::QuadGKJL) (src/Integrals.jl:252-339, scalar out-of-place branch :327-335),
// including the automatic ChangeOfVariables in front of it (src/common.jl:97-103): the adaptive loop
// runs on the transformed interval (tlo, thi) and the integrand is g(t) = f(u(t), p) * du/dt.
//
// Algorithm (QuadGK.jl, restated in oracle/oracle.c orc_quadgk): evaluate the 15-point rule on the
// whole interval; while E > atol && E > rtol*|I| && evals < maxevals: take the segment with the
// largest error, bisect it, evaluate both halves (30 evaluations), update I and E by difference;
// finally re-sum I and E over all segments.
//
// B200 mapping
//   * a persistent grid of warps pulls problems from a global atomic counter (work stealing);
//   * the two half-segments are evaluated together: lanes 0-14 the 15 Kronrod nodes of the left
//     half, lanes 16-30 those of the right half (30 of 32 lanes busy); symmetric pairs are added
//     first (as QuadGK does) and the weighted Kronrod / Gauss sums are formed with xor-shuffles;
//   * the reference's binary heap is replaced by a warp-distributed segment pool: an unsorted
//     structure-of-arrays {E,a,b,I} in shared memory (spilling to a per-warp global arena); the
//     worst segment is found by a strided scan + three REDUX ops on the (non-negative) error bits.
//     Selection order is the reference's (largest error first; ties -> lowest slot).
#include "common.cuh"
#include "family.cuh"
#include "hcubature_pair.cuh"

namespace hp {
template <> int32_t launch<1>(ib200_ctx *, int, bool, Params &);
}

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

__constant__ double c_gk_x[8] = {
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.0};
__constant__ double c_gk_w[8] = {
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_gk_wg[4] = {
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

struct SegPool {
    double *sE, *sA, *sB, *sI;   // shared, capacity capS
    double *gE, *gA, *gB, *gI;   // global spill, capacity capG
    int capS, capG;
    __device__ __forceinline__ double getE(int s) const { return s < capS ? sE[s] : gE[s - capS]; }
    __device__ __forceinline__ void load(int s, double &a, double &b, double &I, double &E) const {
        if (s < capS) { a = sA[s]; b = sB[s]; I = sI[s]; E = sE[s]; }
        else { const int g = s - capS; a = gA[g]; b = gB[g]; I = gI[g]; E = gE[g]; }
    }
    __device__ __forceinline__ void store(int s, double a, double b, double I, double E) const {
        if (s < capS) { sA[s] = a; sB[s] = b; sI[s] = I; sE[s] = E; }
        else { const int g = s - capS; gA[g] = a; gB[g] = b; gI[g] = I; gE[g] = E; }
    }
};

// QuadGK's check_endpoint_roundoff: does the rule evaluate at an endpoint of (a, b)?
__device__ __forceinline__ bool endpoint_roundoff(double a, double b) {
    const double c = 0.5 * (b - a);
    return (a == a + (1 + c_gk_x[0]) * c) || (b == a + (1 - c_gk_x[0]) * c);
}

template <int FAM, bool ALLFIN>
__global__ void __launch_bounds__(kBlock, 4)
quadgk_kernel(const double *__restrict__ params, int nparam, int64_t nprob,
              const double *__restrict__ lb, const double *__restrict__ ub, int bpp, int tr,
              double rtol, double atol, int64_t maxevals, int capS, int capG, double *__restrict__ arena,
              unsigned long long *__restrict__ counter,
              double *__restrict__ out_I, double *__restrict__ out_E, int64_t *__restrict__ out_nevals,
              int32_t *__restrict__ out_status) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + wib;
    SegPool pool;
    pool.capS = capS; pool.capG = capG;
    pool.sE = sm + (size_t)wib * 4 * capS; pool.sA = pool.sE + capS; pool.sB = pool.sA + capS; pool.sI = pool.sB + capS;
    pool.gE = arena + (size_t)gwarp * 4 * capG; pool.gA = pool.gE + capG; pool.gB = pool.gA + capG; pool.gI = pool.gB + capG;
    const int cap = capS + capG;

    // per-lane node constants: lanes (l & 15) = 0..6 -> a + (1 + x_j) s ; 7..13 -> a + (1 - x_j) s ; 14 -> a + s
    const int j = lane & 15;
    const double fac = j < 7 ? 1 + c_gk_x[j] : (j < 14 ? 1 - c_gk_x[j - 7] : 1.0);
    const double wk = j < 7 ? c_gk_w[j] : (j == 14 ? c_gk_w[7] : 0.0);
    const double wg = (j < 7 && (j & 1)) ? c_gk_wg[j >> 1] : (j == 14 ? c_gk_wg[3] : 0.0);
    const bool eval_lane = j < 15;

    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps(Float64))

    for (;;) {
        unsigned long long pidx = 0;
        if (lane == 0) pidx = atomicAdd(counter, 1ULL);
        pidx = __shfl_sync(0xffffffffu, pidx, 0);
        if ((int64_t)pidx >= nprob) break;
        const int64_t prob = (int64_t)pidx;

        ib::Problem<FAM, 1, ALLFIN> P;
        double tlo = -1.0, thi = 1.0;
        {
            const double l = bpp ? lb[prob] : lb[0], u = bpp ? ub[prob] : ub[0];
            P.load(params + prob * nparam, &l, &u, tr);
            if (!ALLFIN) { tlo = P.ax[0].tlo; thi = P.ax[0].thi; }
        }

        // evaluate the rule on the two segments (a0,b0) [lanes 0-15] and (a1,b1) [lanes 16-31]
        auto rule2 = [&](double a0, double b0, double a1, double b1, double &I0, double &E0, double &I1, double &E1) {
            const bool hi = lane >= 16;
            const double a = hi ? a1 : a0, b = hi ? b1 : b0;
            const double s = 0.5 * (b - a);
            double g = 0.0;
            if (eval_lane) { double t[1] = {a + fac * s}; g = P(t); }
            const double partner = __shfl_down_sync(0xffffffffu, g, 7);
            const double v = j < 7 ? g + partner : g;
            double ck = v * wk, cg = v * wg;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                ck += __shfl_xor_sync(0xffffffffu, ck, off);
                cg += __shfl_xor_sync(0xffffffffu, cg, off);
            }
            const double Is = ck * s, Es = fabs(ck * s - cg * s);
            I0 = __shfl_sync(0xffffffffu, Is, 0);  E0 = __shfl_sync(0xffffffffu, Es, 0);
            I1 = __shfl_sync(0xffffffffu, Is, 16); E1 = __shfl_sync(0xffffffffu, Es, 16);
        };

        double I, E, Iu, Eu;
        rule2(tlo, thi, tlo, thi, I, E, Iu, Eu);
        int64_t nev = 15;
        int nseg = 1;
        int status = IB200_ST_CONVERGED;
        if (lane == 0) pool.store(0, tlo, thi, I, E);
        __syncwarp();
        if (!isfinite(E)) status = IB200_ST_NONFINITE;
        else if (!(E <= atol || E <= rtol * fabs(I))) {
            if (nev >= maxevals) status = IB200_ST_MAXEVALS;
            while (E > atol && E > rtol * fabs(I) && nev < maxevals) {
                // --- select the segment with the largest error (ties -> lowest slot)
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
                double sa, sb, sI, sEr;
                pool.load(slot, sa, sb, sI, sEr);
                const double mid = (sa + sb) / 2;
                if (endpoint_roundoff(sa, mid) || endpoint_roundoff(mid, sb)) { status = IB200_ST_ROUNDOFF; break; }
                if (nseg + 1 > cap) { status = IB200_ST_MAXEVALS; break; }
                double I1, E1, I2, E2;
                rule2(sa, mid, mid, sb, I1, E1, I2, E2);
                I = (I - sI) + I1 + I2;
                E = (E - sEr) + E1 + E2;
                nev += 30;
                __syncwarp();
                if (lane == 0) { pool.store(slot, sa, mid, I1, E1); pool.store(nseg, mid, sb, I2, E2); }
                nseg++;
                __syncwarp();
                if (!isfinite(E1) || !isfinite(E2)) { status = IB200_ST_NONFINITE; break; }
            }
            // --- re-sum over all segments (QuadGK resum), fixed order: lane-strided then xor tree
            double pI = 0.0, pE = 0.0;
            for (int s = lane; s < nseg; s += 32) { double a_, b_, i_, e_; pool.load(s, a_, b_, i_, e_); pI += i_; pE += e_; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pI += __shfl_xor_sync(0xffffffffu, pI, off);
                pE += __shfl_xor_sync(0xffffffffu, pE, off);
            }
            I = pI; E = pE;
            if (status == IB200_ST_CONVERGED && !(E <= atol || E <= rtol * fabs(I)) && nev >= maxevals) status = IB200_ST_MAXEVALS;
        }
        if (lane == 0) {
            out_I[prob] = I;
            if (out_E) out_E[prob] = E;
            if (out_nevals) out_nevals[prob] = nev;
            if (out_status) out_status[prob] = status;
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// Thread-per-problem variant (large batches).  ncu on the warp-per-problem kernel above (profiles/r01_quadgk_c1_*): 84 % of the
// issue slots busy with 360 warp instructions per refinement step, of which the 30 integrand evaluations are the smaller part --
// shuffle reductions, the pool scan and the REDUX arg-max are paid per step by 32 lanes for ONE problem.  Here a LANE owns a
// problem: it evaluates both 15-point rules of a bisection itself (sums in QuadGK's own order, so a segment's (I, E) are the oracle's
// doubles), keeps its segments in a lane-interleaved pool (error keys of the first 64 segments in shared memory, so the arg-max scan
// is a conflict-free LDS stream; records and further keys in HBM, coalesced because the lanes of a warp walk the slots together)
// and pulls the next problem from the global counter as soon as it is done, independently of its neighbours.
// Same algorithm and selection order as above: largest error first, ties -> lowest slot, I and E updated by difference, final re-sum.
// ------------------------------------------------------------------------------------------------
constexpr int kTpBlock = 128;
constexpr int kTpKeys = 64;          // error keys per lane in shared memory

template <int FAM, bool ALLFIN>
__device__ __forceinline__ void gk15_thread(const ib::Problem<FAM, 1, ALLFIN> &P, double a, double b, double &Iout, double &Eout) {
    const double s = 0.5 * (b - a);
    double fp[7], fm[7], f0;
    if constexpr (FAM == IB200_FAM_SINXP && ALLFIN) {
        // sin(x p) on a finite interval: the 15 sines in lock step, five at a time (family.cuh sin_bfN) -- libm's sin carries a
        // slow-path branch per call that keeps ptxas from interleaving the independent polynomial chains
        double arg[15], val[15];
#pragma unroll
        for (int j = 0; j < 15; j++) {
            const double fac = j < 7 ? 1 + c_gk_x[j] : (j < 14 ? 1 - c_gk_x[j - 7] : 1.0);
            arg[j] = (P.lb[0] + (1.0 + (a + fac * s)) * P.den[0]) * P.q[0];
        }
        bool big = false;
#pragma unroll
        for (int g0 = 0; g0 < 15; g0 += 5) {
            double x5[5], o5[5];
#pragma unroll
            for (int j = 0; j < 5; j++) x5[j] = arg[g0 + j];
            big = ib::sin_bfN_fast<5>(x5, o5) || big;
#pragma unroll
            for (int j = 0; j < 5; j++) val[g0 + j] = o5[j] * P.jacc;
        }
        if (big) {                    // an argument of 2^30 or more: libm, in one out-of-line call (keeps 15 inlined slow paths out of the refinement loop)
            double tx[15], to[15];
#pragma unroll
            for (int j = 0; j < 15; j++) tx[j] = arg[j];
            ib::sin_libm_n(tx, to, 15);
#pragma unroll
            for (int j = 0; j < 15; j++) val[j] = to[j] * P.jacc;
        }
#pragma unroll
        for (int j = 0; j < 7; j++) { fp[j] = val[j]; fm[j] = val[7 + j]; }
        f0 = val[14];
    } else {
        auto g = [&](double fac) { double t[1] = {a + fac * s}; return P(t); };
#pragma unroll
        for (int j = 0; j < 7; j++) { fp[j] = g(1 + c_gk_x[j]); fm[j] = g(1 - c_gk_x[j]); }
        f0 = g(1.0);
    }
    // QuadGK's evalrule (oracle gk_evalrule): symmetric pairs first, Gauss nodes are the odd ones
    double fg = fp[1] + fm[1], fk = fp[0] + fm[0];
    double Ig = fg * c_gk_wg[0];
    double Ik = fg * c_gk_w[1] + fk * c_gk_w[0];
#pragma unroll
    for (int i = 2; i <= 3; i++) {
        fg = fp[2 * i - 1] + fm[2 * i - 1];
        fk = fp[2 * i - 2] + fm[2 * i - 2];
        Ig += fg * c_gk_wg[i - 1];
        Ik += fg * c_gk_w[2 * i - 1] + fk * c_gk_w[2 * i - 2];
    }
    Ig += f0 * c_gk_wg[3];
    Ik += f0 * c_gk_w[7] + (fp[6] + fm[6]) * c_gk_w[6];
    Iout = Ik * s;
    Eout = fabs(Ik * s - Ig * s);
}

template <int FAM, bool ALLFIN>
__global__ void __launch_bounds__(kTpBlock, 3)
quadgk_tp_kernel(const double *__restrict__ params, int nparam, int64_t nprob,
                 const double *__restrict__ lb, const double *__restrict__ ub, int bpp, int tr,
                 double rtol, double atol, int64_t maxevals, int cap, double *__restrict__ arena,
                 unsigned long long *__restrict__ counter,
                 double *__restrict__ out_I, double *__restrict__ out_E, int64_t *__restrict__ out_nevals,
                 int32_t *__restrict__ out_status) {
    extern __shared__ double sm[];                                  // keys[kTpKeys][kTpBlock]
    const int tid = threadIdx.x;
    const unsigned full = 0xffffffffu;
    double *keyS = sm + tid;
    // lane-interleaved pools of this block: [slot][thread]
    double *gA = arena + (size_t)blockIdx.x * 4 * cap * kTpBlock + tid;
    double *gB = gA + (size_t)cap * kTpBlock, *gI = gB + (size_t)cap * kTpBlock, *gE = gI + (size_t)cap * kTpBlock;
    auto key_ld = [&](int s) { return s < kTpKeys ? keyS[s * kTpBlock] : gE[(size_t)s * kTpBlock]; };
    auto key_st = [&](int s, double e) { if (s < kTpKeys) keyS[s * kTpBlock] = e; else gE[(size_t)s * kTpBlock] = e; };

    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps(Float64))

    ib::Problem<FAM, 1, ALLFIN> P;
    bool have = false, exhausted = false;
    int64_t prob = -1, nev = 0;
    int nseg = 0, status = IB200_ST_CONVERGED;
    double I = 0.0, E = 0.0;

    for (;;) {
        if (!have && !exhausted) {
            const unsigned long long pidx = atomicAdd(counter, 1ULL);
            if ((int64_t)pidx >= nprob) exhausted = true;
            else {
                prob = (int64_t)pidx;
                const double l = bpp ? lb[prob] : lb[0], u = bpp ? ub[prob] : ub[0];
                P.load(params + prob * nparam, &l, &u, tr);
                double tlo = -1.0, thi = 1.0;
                if (!ALLFIN) { tlo = P.ax[0].tlo; thi = P.ax[0].thi; }
                gk15_thread<FAM, ALLFIN>(P, tlo, thi, I, E);
                gA[0] = tlo; gB[0] = thi; gI[0] = I; key_st(0, E);
                nev = 15; nseg = 1; status = IB200_ST_CONVERGED; have = true;
            }
        }
        if (__all_sync(full, !have)) break;                          // every lane of the warp is out of work
        bool refine = false;
        if (have) {
            if (!isfinite(E)) status = IB200_ST_NONFINITE;
            else if (!(E <= atol || E <= rtol * fabs(I))) {
                if (nev >= maxevals || nseg + 1 > cap) status = IB200_ST_MAXEVALS;
                else refine = status == IB200_ST_CONVERGED;
            }
        }
        if (refine) {
            // --- the segment with the largest error (ties -> lowest slot)
            double best = -1.0; int slot = 0;
            for (int s = 0; s < nseg; s++) { const double e = key_ld(s); if (e > best) { best = e; slot = s; } }
            const double sa = gA[(size_t)slot * kTpBlock], sb = gB[(size_t)slot * kTpBlock], sI = gI[(size_t)slot * kTpBlock];
            const double mid = (sa + sb) / 2;
            if (endpoint_roundoff(sa, mid) || endpoint_roundoff(mid, sb)) { status = IB200_ST_ROUNDOFF; refine = false; }
            else {
                double I1, E1, I2, E2;
                gk15_thread<FAM, ALLFIN>(P, sa, mid, I1, E1);
                gk15_thread<FAM, ALLFIN>(P, mid, sb, I2, E2);
                I = (I - sI) + I1 + I2;
                E = (E - best) + E1 + E2;
                nev += 30;
                gB[(size_t)slot * kTpBlock] = mid; gI[(size_t)slot * kTpBlock] = I1; key_st(slot, E1);
                gA[(size_t)nseg * kTpBlock] = mid; gB[(size_t)nseg * kTpBlock] = sb; gI[(size_t)nseg * kTpBlock] = I2; key_st(nseg, E2);
                nseg++;
                if (!isfinite(E1) || !isfinite(E2)) status = IB200_ST_NONFINITE;
            }
        }
        if (have && !refine) {
            // --- done: re-sum over all segments in slot order (QuadGK's closing sum), results out, this lane is free again
            if (nseg > 1) {
                double pI = 0.0, pE = 0.0;
                for (int s = 0; s < nseg; s++) { pI += gI[(size_t)s * kTpBlock]; pE += key_ld(s); }
                I = pI; E = pE;
                if (status == IB200_ST_CONVERGED && !(E <= atol || E <= rtol * fabs(I)) && nev >= maxevals) status = IB200_ST_MAXEVALS;
            }
            out_I[prob] = I;
            if (out_E) out_E[prob] = E;
            if (out_nevals) out_nevals[prob] = nev;
            if (out_status) out_status[prob] = status;
            have = false;
        }
    }
}

}  // namespace

// order == 0: the built-in (7,15) pair; otherwise the rule (x, w, wg) of that order in QuadGK.kronrod's layout (host pointers)
static int32_t quadgk_impl(ib200_ctx *ctx, int32_t family, int32_t transform,
                           const double *lb, const double *ub, int32_t bpp,
                           const double *params, int32_t nparam, int64_t nprob,
                           double reltol, double abstol, int64_t maxevals,
                           int32_t order, const double *gx, const double *gw, const double *gwg,
                           double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_quadgk_batch: ctx is NULL");
    IB_REQUIRE(ctx, family >= 0 && family < IB200_FAM_COUNT, "ib200_quadgk_batch: unknown integrand family");
    IB_REQUIRE(ctx, transform >= 0 && transform <= 2, "ib200_quadgk_batch: unknown transform id");
    IB_REQUIRE(ctx, ib::fam_dim_ok(family, 1) && nparam == ib::fam_nparam(family, 1),
               "QuadGKB200 only accepts one-dimensional quadrature problems.");          // Integrals.jl:263-266
    IB_REQUIRE(ctx, !(reltol < 0) && !(abstol < 0), "ib200_quadgk_batch: negative tolerance");
    IB_REQUIRE(ctx, nprob >= 0, "ib200_quadgk_batch: negative nprob");
    if (nprob == 0) return IB200_OK;
    IB_REQUIRE(ctx, lb && ub && params && out_I, "ib200_quadgk_batch: NULL data pointer");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));

    const size_t nb = bpp ? (size_t)nprob : 1;
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

    // large batches: pair mode (thread per segment, two lanes per problem, 16 problems per warp; hcubature_pair.cuh with
    // D = 1).  ncu on the warp-per-problem kernel below (profiles/r01_quadgk_c1_*): 84 % of the issue slots busy with
    // 360 warp instructions per refinement step, of which the 30 integrand evaluations are a small part.
    // Measured (C1 sweep, kernel time): 65,536 problems 1.29 ms (warp) vs 1.49 ms (pair); 2^20 problems 18.9 vs 14.6 ms:
    // a step is only 30 evaluations, so the pair kernel's per-step pool traffic weighs more than in the cubature case.
    const bool use_pair = order != 0 || ctx->opt_quadgk_mode == 2 || (ctx->opt_quadgk_mode == 0 && nprob >= ((int64_t)1 << 18));
    if (use_pair) {
        const int64_t npts = order ? 2 * (int64_t)order + 1 : 15;
        int64_t need_seg = maxevals >= ((int64_t)1 << 40) ? 4096 : 2 + (maxevals > npts ? (maxevals - npts + 2 * npts - 1) / (2 * npts) : 0);
        if (ctx->opt_hcub_capacity > 0) need_seg = ctx->opt_hcub_capacity;
        if (need_seg > 65536 - 1024) need_seg = 65536 - 1024;
        unsigned long long *counter = (unsigned long long *)ctx->get(SL_WORK1, sizeof(unsigned long long));
        if (!counter) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_quadgk_batch: scratch allocation failed");
        IB_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
        hp::Params q;
        q.params = dpar.d; q.nparam = nparam; q.nprob = nprob; q.lb = dlb.d; q.ub = dub.d; q.bpp = bpp; q.tr = transform;
        q.rtol = reltol; q.atol = abstol; q.maxevals = maxevals; q.initdiv = 1;
        q.cap = (int)((need_seg + 1023) / 1024 * 1024);
        q.keys = q.l1 = q.rec = nullptr; q.l1off = nullptr; q.counter = counter;
        q.order = nullptr; q.norder = nullptr;
        q.gk_order = order; q.gk_x = q.gk_w = q.gk_wg = nullptr;
        DevIn<double> dgx, dgw, dgwg;
        if (order) {
            if ((rc = dgx.init(ctx, gx, (size_t)order + 1, SL_IN3)) || (rc = dgw.init(ctx, gw, (size_t)order + 1, SL_IN4)) ||
                (rc = dgwg.init(ctx, gwg, (size_t)(order + 1) / 2, SL_IN5))) return rc;
            q.gk_x = dgx.d; q.gk_w = dgw.d; q.gk_wg = dgwg.d;
        }
        q.out_I = oI.d; q.out_E = oE.d; q.out_nevals = oN.d; q.out_status = oS.d;
        ctx->err.clear();
        ib200_time_begin(ctx);
        if ((rc = hp::launch<1>(ctx, family, allfin, q))) return rc;
        ib200_time_end(ctx);
        if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
        if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return IB200_OK;
    }

    // thread per problem (quadgk_tp_kernel): batches large enough to give every lane of the GPU a problem of its own
    const bool use_tp = ctx->opt_quadgk_mode == 3 || (ctx->opt_quadgk_mode == 0 && nprob >= 8192);
    if (use_tp) {
        int64_t need_seg = maxevals >= ((int64_t)1 << 40) ? 2048 : 2 + (maxevals > 15 ? (maxevals - 15 + 29) / 30 : 0);
        if (ctx->opt_hcub_capacity > 0) need_seg = ctx->opt_hcub_capacity;
        if (need_seg < kTpKeys) need_seg = kTpKeys;
        if (need_seg > 8192) need_seg = 8192;
        const int cap = (int)need_seg;
        const size_t smem = (size_t)kTpKeys * kTpBlock * sizeof(double);
        int64_t blocks = (nprob + kTpBlock - 1) / kTpBlock;
        const int64_t maxb = (int64_t)ctx->sm_count * 3;
        if (blocks > maxb) blocks = maxb;
        double *arena = (double *)ctx->get(SL_WORK0, (size_t)blocks * 4 * (size_t)cap * kTpBlock * sizeof(double));
        unsigned long long *counter = (unsigned long long *)ctx->get(SL_WORK1, sizeof(unsigned long long));
        if (!arena || !counter) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_quadgk_batch: scratch allocation failed");
        IB_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
        bool launched = false;
        cudaError_t attr_err = cudaSuccess;
        ib200_time_begin(ctx);
        ib200_dispatch_family(family, [&](auto famc) {
            constexpr int FAM = decltype(famc)::value;
            if constexpr (ib::fam_dim_ok(FAM, 1)) {
                auto go = [&](auto kern) {
                    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kern<<<(unsigned)blocks, kTpBlock, smem, ctx->stream>>>(dpar.d, nparam, nprob, dlb.d, dub.d, bpp, transform,
                                                                          reltol, abstol, maxevals, cap, arena, counter, oI.d, oE.d, oN.d, oS.d);
                    launched = true;
                };
                if (allfin) go(quadgk_tp_kernel<FAM, true>); else go(quadgk_tp_kernel<FAM, false>);
            }
        });
        if (attr_err != cudaSuccess) { cudaGetLastError(); return ib200_fail(ctx, IB200_ERR_CUDA, "cudaFuncSetAttribute failed"); }
        if (!launched) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_quadgk_batch: family not compiled in for dim 1");
        IB_LAUNCH_CHECK(ctx);
        ib200_time_end(ctx);
        if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
        if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return IB200_OK;
    }

    // shared pool: 5 CTAs/SM target -> ~44 KB dynamic smem per CTA
    int ctas_per_sm = 4;
    if (ctx->opt_adaptive_warps_per_sm > 0) ctas_per_sm = (int)((ctx->opt_adaptive_warps_per_sm + kWarps - 1) / kWarps);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    size_t smem_budget = (size_t)(224 * 1024) / ctas_per_sm - 1024;   // 1 KB/CTA is reserved by the runtime
    if (smem_budget > ctx->smem_optin) smem_budget = ctx->smem_optin;
    int capS = (int)(smem_budget / (kWarps * 4 * sizeof(double)));
    capS &= ~7;
    if (capS < 32) capS = 32;
    const size_t smem = (size_t)kWarps * 4 * capS * sizeof(double);
    // global spill: enough for maxevals, bounded at 8192 segments per warp
    int64_t need = maxevals >= (int64_t)1 << 40 ? (int64_t)1 << 40 : 1 + (maxevals > 15 ? (maxevals - 15 + 29) / 30 : 0);
    int64_t capG = need - capS;
    if (capG < 8) capG = 8;
    if (capG > 8192) capG = 8192;
    int64_t blocks = (nprob + kWarps - 1) / kWarps;
    const int64_t maxb = (int64_t)ctx->sm_count * ctas_per_sm;
    if (blocks > maxb) blocks = maxb;
    double *arena = (double *)ctx->get(SL_WORK0, (size_t)blocks * kWarps * 4 * (size_t)capG * sizeof(double));
    unsigned long long *counter = (unsigned long long *)ctx->get(SL_WORK1, sizeof(unsigned long long));
    if (!arena || !counter) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_quadgk_batch: scratch allocation failed");
    IB_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));

    bool launched = false;
    cudaError_t attr_err = cudaSuccess;
    ib200_time_begin(ctx);
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, 1)) {
            auto go = [&](auto kern) {
                attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(dpar.d, nparam, nprob, dlb.d, dub.d, bpp, transform,
                                                                     reltol, abstol, maxevals, capS, (int)capG, arena, counter,
                                                                     oI.d, oE.d, oN.d, oS.d);
                launched = true;
            };
            if (allfin) go(quadgk_kernel<FAM, true>); else go(quadgk_kernel<FAM, false>);
        }
    });
    if (attr_err != cudaSuccess) { cudaGetLastError(); return ib200_fail(ctx, IB200_ERR_CUDA, "cudaFuncSetAttribute failed"); }
    if (!launched) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_quadgk_batch: family not compiled in for dim 1");
    IB_LAUNCH_CHECK(ctx);
    ib200_time_end(ctx);
    if ((rc = oI.finish(ctx)) || (rc = oE.finish(ctx)) || (rc = oN.finish(ctx)) || (rc = oS.finish(ctx))) return rc;
    if (oI.is_host() || oE.is_host() || oN.is_host() || oS.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}

extern "C" int32_t ib200_quadgk_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                      const double *lb, const double *ub, int32_t bpp,
                                      const double *params, int32_t nparam, int64_t nprob,
                                      double reltol, double abstol, int64_t maxevals,
                                      double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    return quadgk_impl(ctx, family, transform, lb, ub, bpp, params, nparam, nprob, reltol, abstol, maxevals, 0, nullptr, nullptr, nullptr,
                       out_I, out_E, out_nevals, out_status);
}

extern "C" int32_t ib200_quadgk_rule_batch(ib200_ctx *ctx, int32_t family, int32_t transform,
                                           const double *lb, const double *ub, int32_t bpp,
                                           const double *params, int32_t nparam, int64_t nprob,
                                           double reltol, double abstol, int64_t maxevals,
                                           int32_t order, const double *x, const double *w, const double *wg,
                                           double *out_I, double *out_E, int64_t *out_nevals, int32_t *out_status) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_quadgk_rule_batch: ctx is NULL");
    IB_REQUIRE(ctx, order >= 2 && order <= hp::kGkMaxOrder, "ib200_quadgk_rule_batch: 2 <= order <= 30");
    IB_REQUIRE(ctx, x && w && wg, "ib200_quadgk_rule_batch: NULL rule table");
    return quadgk_impl(ctx, family, transform, lb, ub, bpp, params, nparam, nprob, reltol, abstol, maxevals, order, x, w, wg,
                       out_I, out_E, out_nevals, out_status);
}
