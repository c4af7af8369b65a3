// hcubature_impl.cuh -- batched adaptive cubature (Genz-Malik degree 7/5, or GK(7,15) in 1-D),
// one warp per problem.  Included by hcubature_d<D>.cu, one translation unit per dimension.
//
// Replaces HCubature.hcubature / hquadrature(f, lb, ub; rtol, atol, maxevals, initdiv) as called from
// __solvebp_call(::HCubatureJL) (src/Integrals.jl:347-397, call sites :380-393) including the
// automatic ChangeOfVariables in front of it (src/common.jl:97-103).  Algorithm (HCubature.jl,
// restated in oracle/oracle.c orc_hcubature / gm_rule): rule on the initial box(es); until
// E <= max(rtol*|I|, atol) or evals >= maxevals or E non-finite: take the box with the largest
// error, halve it along the axis of largest fourth difference, apply the rule to both halves,
// update I and E by difference; finally re-sum I and E over all boxes.
//
// B200 mapping (everything per warp; a persistent grid pulls problems from an atomic counter)
//   * Rule evaluation in two phases.  The Genz-Malik points of a box only use 7 abscissae per axis
//     (centre, +-l2, +-l3, +-l5 half-widths), and every registered integrand has the separable form
//     f(x) = finish(init (+|*) term_1(x_1) ... term_d(x_d)) (family.cuh).  Phase 1 tabulates the
//     2 boxes x d axes x 7 abscissae per-axis terms (variable transform included) in shared memory;
//     phase 2 spreads the 2*NP points over the lanes, each folding d table entries and applying
//     finish (the transcendental) once.  The integrand is still evaluated at every point of the rule.
//   * Weighted sums of both rules are combined with xor-shuffles so that lanes 0-15 end up with
//     child 0 and lanes 16-31 with child 1; each half-warp then forms (I, E) and picks the split
//     axis of its child from the fourth differences.
//   * The reference's binary heap is replaced by a two-level tournament: error keys of all boxes in
//     global memory in groups of 64, the maximum of every group in shared memory.  Selecting the
//     worst box = strided scan of the shared level + REDUX arg-max, one coalesced 512-byte key-group
//     load + REDUX arg-max.  Child 0 overwrites the parent's slot, child 1 is appended.  Selection
//     order is the reference's (largest error first; ties -> lowest slot).
#pragma once
#include "common.cuh"
#include "family.cuh"

namespace hc {

constexpr int kWarpsPerBlock = 4;
constexpr int kBlock = 32 * kWarpsPerBlock;
constexpr int kGroup = 64;   // boxes per tournament group

struct Params {
    const double *params; int nparam; int64_t nprob;
    const double *lb, *ub; int bpp, tr;
    double rtol, atol; int64_t maxevals; int initdiv;
    int cap;                    // boxes per warp (multiple of kGroup)
    double *keys;               // [warps][cap]        error of every box
    double *ivals;              // [warps][cap]        integral of every box
    double *recs;               // [warps][cap][RECW]  a[D], b[D], kdiv
    unsigned long long *counter;
    double *out_I, *out_E; int64_t *out_nevals; int32_t *out_status;
};

__host__ __device__ constexpr int gm_npoints(int d) { return d == 1 ? 15 : 1 + 4 * d + 2 * d * (d - 1) + (1 << d); }
__host__ __device__ constexpr int rec_width(int d) { return 2 * d + 1; }

__constant__ double c_x[8] = {
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.0};
__constant__ double c_w[8] = {
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_wg[4] = {
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

// warp arg-max of non-negative double keys (as ordered 64-bit patterns); ties -> smallest idx.
// Lanes without a candidate pass idx = 0x7fffffff.  Returns the winning idx (uniform).
__device__ __forceinline__ int warp_argmax(unsigned long long key, int idx) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const bool valid = idx != 0x7fffffff;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, valid ? hi : 0u);
    const bool c1 = valid && hi == mhi;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
    const bool win = c1 && lo == mlo;
    return (int)__reduce_min_sync(0xffffffffu, win ? (unsigned)idx : 0x7fffffffu);
}
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long key) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ unsigned long long dkey(double e) { return (unsigned long long)__double_as_longlong(e); }

// per-warp shared-memory footprint in doubles (besides the tournament level)
template <int D> __host__ __device__ constexpr int warp_stage_doubles() {
    // tabT[2][D][7] + tabJ[2][D][7] + fst[2][1+4D] + hs[2][D] + box[2][2D] + par[3D+1] + lb[D] + ub[D] + den[D]
    return 14 * D + 14 * D + 2 * (1 + 4 * D) + 2 * D + 4 * D + (3 * D + 1) + 3 * D;
}

template <int FAM, int D, bool ALLFIN>
__global__ void __launch_bounds__(kBlock)
hcub_kernel(const Params p) {
    using F = ib::Family<FAM, D>;
    constexpr int NP = gm_npoints(D);
    constexpr int RECW = rec_width(D);
    constexpr int NAX = 1 + 4 * D;             // centre + lambda2/lambda3 axis points (D >= 2)
    constexpr int NPAR = ib::fam_nparam(FAM, D);
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
    const int ngroups = p.cap / kGroup;
    // shared layout: [warps][ngroups] group maxima ; [warps] stage ; weights[10] ; point table
    unsigned long long *L1 = reinterpret_cast<unsigned long long *>(sm) + (size_t)wib * ngroups;
    double *stage = sm + (size_t)kWarpsPerBlock * ngroups + (size_t)wib * warp_stage_doubles<D>();
    double *tabT = stage;                      // [2][D][7]
    double *tabJ = tabT + 14 * D;              // [2][D][7]
    double *fst = tabJ + 14 * D;               // [2][NAX]
    double *hs = fst + 2 * NAX;                // [2][D]
    double *box = hs + 2 * D;                  // [2][2D]  a[D], b[D] of the two boxes under evaluation
    double *par = box + 4 * D;                 // [NPAR] (<= 3D+1)
    double *lbs = par + (3 * D + 1);           // [D]
    double *ubs = lbs + D;                     // [D]
    double *dens = ubs + D;                    // [D]  (ALLFIN) (ub-lb)*0.5
    double *wtab = sm + (size_t)kWarpsPerBlock * ngroups + (size_t)kWarpsPerBlock * warp_stage_doubles<D>();  // [10]
    unsigned *pt = reinterpret_cast<unsigned *>(wtab + 10);
    double *keys = p.keys + (size_t)gwarp * p.cap;
    double *ivals = p.ivals + (size_t)gwarp * p.cap;
    double *recs = p.recs + (size_t)gwarp * p.cap * RECW;

    const double l2 = sqrt(9.0 / 70.0), l3 = sqrt(9.0 / 10.0), l5 = sqrt(9.0 / 19.0);
    if (D >= 2) {
        // Genz-Malik weights (HCubature.jl genz-malik.jl; oracle gm_rule): wtab[cls], wtab[5 + cls]
        if (threadIdx.x == 0) {
            const double twon = (double)(1 << D);
            wtab[0] = twon * ((12824 - 9120 * D + 400 * D * D) / 19683.0);
            wtab[1] = twon * (980 / 6561.0);
            wtab[2] = twon * ((1820 - 400 * D) / 19683.0);
            wtab[3] = twon * (200 / 19683.0);
            wtab[4] = 6859 / 19683.0;
            wtab[5] = twon * ((729 - 950 * D + 50 * D * D) / 729.0);
            wtab[6] = twon * (245 / 486.0);
            wtab[7] = twon * ((265 - 100 * D) / 1458.0);
            wtab[8] = twon * (25 / 729.0);
            wtab[9] = 0.0;
        }
        // point table: 3 bits per axis = abscissa index (0 centre, 1/2 +-l2, 3/4 +-l3, 5/6 +-l5); class in bits 24-26
        for (int q = threadIdx.x; q < NP; q += kBlock) {
            unsigned e = 0, cls = 0;
            if (q == 0) { cls = 0; }
            else if (q < 1 + 2 * D) { const int r = q - 1; cls = 1; e = (1u + (r & 1)) << (3 * (r >> 1)); }
            else if (q < 1 + 4 * D) { const int r = q - 1 - 2 * D; cls = 2; e = (3u + (r & 1)) << (3 * (r >> 1)); }
            else if (q < 1 + 4 * D + 2 * D * (D - 1)) {
                int r = q - 1 - 4 * D; const int s = r & 3; r >>= 2;
                int i = 0, j = 1;
                for (int cnt = 0; cnt < r; cnt++) { if (++j == D) { ++i; j = i + 1; } }
                cls = 3;
                e = ((3u + ((s >> 1) & 1)) << (3 * i)) | ((3u + (s & 1)) << (3 * j));
            } else {
                const unsigned r = q - (1 + 4 * D + 2 * D * (D - 1)); cls = 4;
                for (int k = 0; k < D; k++) e |= (5u + ((r >> k) & 1u)) << (3 * k);
            }
            pt[q] = e | (cls << 24);
        }
    }
    __syncthreads();

    double rtol = p.rtol;
    const double atol = p.atol;
    if (rtol == 0.0 && atol == 0.0) rtol = 1.4901161193847656e-08;   // sqrt(eps), HCubature default

    for (;;) {
        unsigned long long pidx = 0;
        if (lane == 0) pidx = atomicAdd(p.counter, 1ULL);
        pidx = __shfl_sync(0xffffffffu, pidx, 0);
        if ((int64_t)pidx >= p.nprob) break;
        const int64_t prob = (int64_t)pidx;

        // ---- problem data -> shared
        __syncwarp();
        if (lane < NPAR) par[lane] = p.params[prob * p.nparam + lane];
        if (lane < D) {
            const double l = p.bpp ? p.lb[prob * D + lane] : p.lb[lane];
            const double u = p.bpp ? p.ub[prob * D + lane] : p.ub[lane];
            lbs[lane] = l; ubs[lane] = u; dens[lane] = (u - l) * 0.5;
        }
        __syncwarp();
        double jacc = 1.0;                      // ALLFIN: constant Jacobian prod_k den_k
        if (ALLFIN) {
#pragma unroll
            for (int k = 0; k < D; k++) jacc = (k == 0) ? dens[0] : jacc * dens[k];
        }
        const double finit = F::sep_init(par);

        // t -> (u, jac) on axis k
        auto map_axis = [&](int k, double t, double &u, double &jac) {
            if (ALLFIN) { u = lbs[k] + (1.0 + t) * dens[k]; jac = 1.0; }
            else { ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m); ib::t2ujac(p.tr, t, m, u, jac); }
        };

        // rule on the two boxes staged in box[0], box[1]; results uniform in all lanes
        auto rule2 = [&](bool second, double &I0, double &E0, int &k0, double &I1, double &E1, int &k1) {
            if constexpr (D == 1) {
                // HCubature's 1-D rule: GK(7,15) about the centre with signed half-width
                const int ch = lane >> 4, j = lane & 15;
                const double a = box[2 * ch], b = box[2 * ch + 1];
                const double c = (a + b) * 0.5, Dl = (b - a) * 0.5;
                double g = 0.0;
                if (j < 15 && (second || ch == 0)) {
                    const double dx = j < 14 ? Dl * c_x[j < 7 ? j : j - 7] : 0.0;
                    const double t = j < 7 ? c + dx : (j < 14 ? c - dx : c);
                    double u, jac;
                    map_axis(0, t, u, jac);
                    const double T = F::sep_term(0, u, par);
                    g = F::sep_finish(F::kProd ? finit * T : finit + T, par) * (ALLFIN ? jacc : jac);
                }
                const double partner = __shfl_down_sync(0xffffffffu, g, 7);
                const double v = j < 7 ? g + partner : g;
                double ck = v * (j < 7 ? c_w[j] : (j == 14 ? c_w[7] : 0.0));
                double cg = v * ((j < 7 && (j & 1)) ? c_wg[j >> 1] : (j == 14 ? c_wg[3] : 0.0));
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    ck += __shfl_xor_sync(0xffffffffu, ck, off);
                    cg += __shfl_xor_sync(0xffffffffu, cg, off);
                }
                const double Ik = ck * Dl, Ig = cg * Dl, Er = fabs(Ik - Ig);
                I0 = __shfl_sync(0xffffffffu, Ik, 0);  E0 = __shfl_sync(0xffffffffu, Er, 0);
                I1 = __shfl_sync(0xffffffffu, Ik, 16); E1 = __shfl_sync(0xffffffffu, Er, 16);
                k0 = 0; k1 = 0;
            } else {
                // phase 1: per-axis term tables  idx = (ch*D + k)*7 + o
                const int ntask = second ? 14 * D : 7 * D;
                for (int idx = lane; idx < ntask; idx += 32) {
                    const int ch = idx >= 7 * D;
                    const int r = idx - ch * 7 * D;
                    const int k = r / 7, o = r - 7 * k;
                    const double a = box[ch * 2 * D + k], b = box[ch * 2 * D + D + k];
                    const double c = 0.5 * (a + b), h = 0.5 * fabs(b - a);
                    const double lam = o == 0 ? 0.0 : (o <= 2 ? l2 : (o <= 4 ? l3 : l5));
                    const double dl = h * lam;
                    const double t = o == 0 ? c : ((o & 1) ? c + dl : c - dl);
                    double u, jac;
                    map_axis(k, t, u, jac);
                    tabT[idx] = F::sep_term(k, u, par);
                    if (!ALLFIN) tabJ[idx] = jac;
                    if (o == 0) hs[ch * D + k] = h;
                }
                __syncwarp();
                // phase 2: the points
                double sI0 = 0.0, sP0 = 0.0, sI1 = 0.0, sP1 = 0.0;
                const int total = second ? 2 * NP : NP;
                for (int w = lane; w < total; w += 32) {
                    const int ch = w >= NP;
                    const int q = w - ch * NP;
                    const unsigned e = pt[q];
                    const double *tT = tabT + ch * 7 * D;
                    const double *tJ = tabJ + ch * 7 * D;
                    double acc = finit, jac = 1.0;
#pragma unroll
                    for (int k = 0; k < D; k++) {
                        const int sel = (e >> (3 * k)) & 7u;
                        const double T = tT[7 * k + sel];
                        acc = F::kProd ? acc * T : acc + T;
                        if (!ALLFIN) { const double J = tJ[7 * k + sel]; jac = (k == 0) ? J : jac * J; }
                    }
                    const double f = F::sep_finish(acc, par) * (ALLFIN ? jacc : jac);
                    if (q < NAX) fst[ch * NAX + q] = f;
                    const unsigned cls = e >> 24;
                    const double wI = wtab[cls], wP = wtab[5 + cls];
                    if (ch) { sI1 = fma(wI, f, sI1); sP1 = fma(wP, f, sP1); }
                    else    { sI0 = fma(wI, f, sI0); sP0 = fma(wP, f, sP0); }
                }
                // combine: lanes 0-15 keep child 0, lanes 16-31 keep child 1
                const bool hi = lane >= 16;
                double kI = hi ? sI1 : sI0, kP = hi ? sP1 : sP0;
                kI += __shfl_xor_sync(0xffffffffu, hi ? sI0 : sI1, 16);
                kP += __shfl_xor_sync(0xffffffffu, hi ? sP0 : sP1, 16);
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    kI += __shfl_xor_sync(0xffffffffu, kI, off);
                    kP += __shfl_xor_sync(0xffffffffu, kP, off);
                }
                __syncwarp();
                // per half-warp: volume, I, E, split axis (largest fourth difference, HCubature's tie fuzz)
                const int ch = hi ? 1 : 0;
                const double *hh = hs + ch * D;
                const double *ff = fst + ch * NAX;
                double V = hh[0];
#pragma unroll
                for (int k = 1; k < D; k++) V *= hh[k];
                const double Iv = V * kI, Ip = V * kP;
                const double Er = fabs(Iv - Ip);
                double p10 = 1.0;
#pragma unroll
                for (int k = 0; k < D; k++) p10 *= 10.0;
                const double deltaf = Er / (p10 * V);
                const double twelvef1 = 12 * ff[0];
                int kd = 0; double maxdd = 0.0, hkd = hh[0];
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const double f2k = ff[1 + 2 * k] + ff[2 + 2 * k];
                    const double f3k = ff[1 + 2 * D + 2 * k] + ff[2 + 2 * D + 2 * k];
                    const double dd = fabs(f3k + twelvef1 - 7 * f2k);
                    const double delta = dd - maxdd;
                    if (delta > deltaf) { kd = k; maxdd = dd; hkd = hh[k]; }
                    else if (fabs(delta) <= deltaf && hh[k] > hkd) { kd = k; hkd = hh[k]; }
                }
                I0 = __shfl_sync(0xffffffffu, Iv, 0);  E0 = __shfl_sync(0xffffffffu, Er, 0);  k0 = __shfl_sync(0xffffffffu, kd, 0);
                I1 = __shfl_sync(0xffffffffu, Iv, 16); E1 = __shfl_sync(0xffffffffu, Er, 16); k1 = __shfl_sync(0xffffffffu, kd, 16);
                __syncwarp();
            }
        };

        // ---- pool helpers (uniform control flow)
        int nbox = 0;
        auto write_box = [&](int slot, int ch, double I, double E, int kd) {
            if (lane < RECW) recs[(size_t)slot * RECW + lane] = lane < 2 * D ? box[ch * 2 * D + lane] : (double)kd;
            if (lane == 0) { keys[slot] = E; ivals[slot] = I; }
        };
        auto append = [&](int ch, double I, double E, int kd) {
            const int slot = nbox;
            write_box(slot, ch, I, E, kd);
            if (lane == 0) {
                const int g = slot / kGroup;
                const unsigned long long k = dkey(E);
                L1[g] = (slot % kGroup == 0) ? k : (L1[g] > k ? L1[g] : k);
            }
            nbox++;
            __syncwarp();
        };

        // ---- initial box(es): HCubature's initdiv grid, CartesianIndices order (first index fastest)
        double I = 0.0, E = 0.0;
        int64_t nev = 0;
        int status = IB200_ST_CONVERGED;
        bool degenerate = false;
        {
            const int initdiv = p.initdiv;
            int64_t total = 1;
            double prodD = 1.0;
#pragma unroll
            for (int k = 0; k < D; k++) {
                total *= initdiv;
                double tl, th;
                if (ALLFIN) { tl = -1.0; th = 1.0; } else { ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m); tl = m.tlo; th = m.thi; }
                prodD *= (th - tl) / initdiv;
            }
            degenerate = (prodD == 0.0);
            if (degenerate) total = 1;              // HCubature returns after the first box
            for (int64_t q = 0; q < total; q += 2) {
                const bool second = q + 1 < total;
                if (lane < 2 * D) {                 // lane -> (box, axis)
                    const int ch = lane >= D, k = lane - ch * D;
                    int64_t r = q + ch;
                    for (int kk = 0; kk < k; kk++) r /= initdiv;
                    const int c = (int)(r % initdiv);
                    double tl, th;
                    if (ALLFIN) { tl = -1.0; th = 1.0; } else { ib::AxisMap m; ib::make_axis(lbs[k], ubs[k], m); tl = m.tlo; th = m.thi; }
                    const double Dl = (th - tl) / initdiv;
                    const double a = tl + c * Dl;
                    const double b = (c == initdiv - 1) ? th : a + Dl;
                    box[ch * 2 * D + k] = a; box[ch * 2 * D + D + k] = b;
                }
                __syncwarp();
                double I0, E0, I1, E1; int k0, k1;
                rule2(second, I0, E0, k0, I1, E1, k1);
                if (q == 0) { I = I0; E = E0; } else { I += I0; E += E0; }
                nev += NP;
                append(0, I0, E0, k0);
                if (second) { I += I1; E += E1; nev += NP; append(1, I1, E1, k1); }
            }
        }

        // ---- adaptive loop
        if (!degenerate) {
            for (;;) {
                if (E <= fmax(rtol * fabs(I), atol)) break;
                if (nev >= p.maxevals) { status = IB200_ST_MAXEVALS; break; }
                if (!isfinite(E)) { status = IB200_ST_NONFINITE; break; }
                if (nbox + 1 > p.cap) { status = IB200_ST_MAXEVALS; break; }
                // select: group maxima (shared), then the group's 64 keys (global)
                const int n1 = (nbox + kGroup - 1) / kGroup;
                unsigned long long bk = 0ULL; int bi = 0x7fffffff;
                for (int g = lane; g < n1; g += 32) {
                    const unsigned long long k = L1[g];
                    if (bi == 0x7fffffff || k > bk) { bk = k; bi = g; }
                }
                const int g = warp_argmax(bk, bi);
                const int s_lo = g * kGroup + lane, s_hi = s_lo + 32;
                const unsigned long long klo = s_lo < nbox ? dkey(keys[s_lo]) : 0ULL;
                const unsigned long long khi = s_hi < nbox ? dkey(keys[s_hi]) : 0ULL;
                unsigned long long kk = klo; int ki = s_lo < nbox ? s_lo : 0x7fffffff;
                if (s_hi < nbox && khi > kk) { kk = khi; ki = s_hi; }
                const int slot = warp_argmax(kk, ki);
                // parent: error = its key, integral from ivals, geometry from the record
                const double Ebox = __longlong_as_double((long long)warp_max_key(ki == slot ? kk : 0ULL));
                const double Ibox = ivals[slot];
                double rv = 0.0;
                if (lane < RECW) rv = recs[(size_t)slot * RECW + lane];
                const int kd = (int)__shfl_sync(0xffffffffu, rv, 2 * D);
                // children: child 0 = upper half (a[kd] += w), child 1 = lower half (b[kd] -= w)
                {
                    const double akd = __shfl_sync(0xffffffffu, rv, kd), bkd = __shfl_sync(0xffffffffu, rv, D + kd);
                    const double w = (bkd - akd) * 0.5;
                    if (lane < 2 * D) {
                        box[lane] = (lane == kd) ? akd + w : rv;
                        box[2 * D + lane] = (lane == D + kd) ? bkd - w : rv;
                    }
                }
                __syncwarp();
                double I1, E1, I2, E2; int k1, k2;
                rule2(true, I1, E1, k1, I2, E2, k2);
                nev += 2 * NP;
                I += I1 + I2 - Ibox;
                E += E1 + E2 - Ebox;
                // child 0 replaces the parent's slot: recompute the group maximum from the keys in registers
                write_box(slot, 0, I1, E1, k1);
                {
                    const unsigned long long e1 = dkey(E1);
                    const unsigned long long nlo = (s_lo == slot) ? e1 : klo, nhi = (s_hi == slot) ? e1 : khi;
                    const unsigned long long gm = warp_max_key(nlo > nhi ? nlo : nhi);
                    if (lane == 0) L1[g] = gm;
                }
                __syncwarp();
                append(1, I2, E2, k2);
            }
        }

        // ---- re-sum over all boxes (HCubature's "roundoff paranoia"): lane-strided, then xor tree
        {
            double pI = 0.0, pE = 0.0;
            for (int s = lane; s < nbox; s += 32) { pI += ivals[s]; pE += keys[s]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pI += __shfl_xor_sync(0xffffffffu, pI, off);
                pE += __shfl_xor_sync(0xffffffffu, pE, off);
            }
            I = pI; E = pE;
        }
        if (lane == 0) {
            p.out_I[prob] = I;
            if (p.out_E) p.out_E[prob] = E;
            if (p.out_nevals) p.out_nevals[prob] = nev;
            if (p.out_status) p.out_status[prob] = status;
        }
        __syncwarp();
    }
}

template <int D>
inline size_t smem_bytes(int cap) {
    return ((size_t)kWarpsPerBlock * (cap / kGroup) + (size_t)kWarpsPerBlock * warp_stage_doubles<D>() + 10) * sizeof(double) +
           (size_t)gm_npoints(D) * sizeof(unsigned) + 16;
}

// launcher for one dimension; instantiated in hcubature_d<D>.cu.  Sizes the persistent grid from the
// kernel's occupancy, carves the per-warp box pools out of the context's scratch and launches.
template <int D>
int32_t launch(ib200_ctx *ctx, int family, bool allfin, Params &p);

template <int D>
int32_t launch_impl(ib200_ctx *ctx, int family, bool allfin, Params &p) {
    constexpr int RECW = rec_width(D);
    const size_t smem = smem_bytes<D>(p.cap);
    if (smem > ctx->smem_optin) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_hcubature_batch: box capacity too large for shared memory");
    int32_t rc = IB200_ERR_UNSUPPORTED;
    ib200_dispatch_family(family, [&](auto famc) {
        constexpr int FAM = decltype(famc)::value;
        if constexpr (ib::fam_dim_ok(FAM, D)) {
            auto go = [&](auto kern) -> int32_t {
                IB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int occ = 0;
                IB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem));
                if (occ < 1) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_hcubature_batch: kernel does not fit on an SM");
                if (ctx->opt_adaptive_warps_per_sm > 0) {
                    const int want = (int)((ctx->opt_adaptive_warps_per_sm + kWarpsPerBlock - 1) / kWarpsPerBlock);
                    if (want < occ) occ = want;
                }
                int64_t blocks = (p.nprob + kWarpsPerBlock - 1) / kWarpsPerBlock;
                if (blocks > (int64_t)ctx->sm_count * occ) blocks = (int64_t)ctx->sm_count * occ;
                const size_t warps = (size_t)blocks * kWarpsPerBlock;
                p.keys = (double *)ctx->get(SL_WORK0, warps * p.cap * sizeof(double));
                p.ivals = (double *)ctx->get(SL_WORK3, warps * p.cap * sizeof(double));
                p.recs = (double *)ctx->get(SL_WORK2, warps * p.cap * RECW * sizeof(double));
                if (!p.keys || !p.ivals || !p.recs)
                    return ib200_fail(ctx, IB200_ERR_OOM, "ib200_hcubature_batch: box pool allocation failed (lower maxiters or the hcub_capacity option)");
                kern<<<(unsigned)blocks, kBlock, smem, ctx->stream>>>(p);
                IB_LAUNCH_CHECK(ctx);
                return IB200_OK;
            };
            rc = allfin ? go(hcub_kernel<FAM, D, true>) : go(hcub_kernel<FAM, D, false>);
        }
    });
    if (rc == IB200_ERR_UNSUPPORTED && ctx->err.empty())
        return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_hcubature_batch: family not compiled in for this dimension");
    return rc;
}

}  // namespace hc
