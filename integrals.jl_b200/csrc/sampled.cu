// sampled.cu -- trapezoid / Simpson integration of sampled data as an HBM streaming reduction.
//
// Replaces evalrule(data, weights, dim) (src/sampled.jl:51-124) and the lazy weight types of
// src/trapezoidal.jl:16-45 and src/simpsons.jl:16-87.  y is column-major [inner, n, outer]; the
// integration axis is the middle one, out is [inner, outer]:
//     out[j, o] = sum_i w[i] * y[j, i, o]          (separate multiply and add, like the reference)
//
// Kernels
//   weights_kernel          w[i] from (n,h) or from x[] -- n doubles, negligible traffic
//   strided_kernel<VEC>     inner >= 2: every thread owns VEC adjacent outputs and streams a slab of
//                           columns with 128-bit loads (the contiguous axis is `inner`);
//                           partial[o][slab][j], then
//   contiguous_kernel       inner == 1: a CTA streams a slab of one row, block tree reduction
//   finish_kernel           fixed-order sum over slabs -> deterministic, no atomics
// Algorithmic bytes per solve: 8*inner*n*outer (y) + 8*inner*outer (out) [+ 8n for x].
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// T = element type of the grid: the weights are formed in T arithmetic, as the reference's weight types do
// (n, x): the WHOLE grid; this launch writes the weights of the points off .. off + cnt - 1 to w[0 .. cnt) (a slab of the
// integration axis; off = 0, cnt = n for a whole solve)
template <typename T>
__global__ void weights_kernel(int rule, int64_t n, T h, const T *__restrict__ x, T *__restrict__ w, int64_t off, int64_t cnt) {
    const int64_t l0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l0 >= cnt) return;
    const int64_t i0 = off + l0;                                        // 0-based; formulas below use 1-based i
    const int64_t i = i0 + 1;
    T r;
    if (!x) {
        if (rule == IB200_RULE_TRAPEZOID) {                     // trapezoidal.jl:16-20
            r = (i == 1 || i == n) ? h / 2 : h;
        } else {                                                // simpsons.jl:16-25 (first match wins)
            if (i == 1 || i == n) r = 17 * h / 48;
            else if (i == 2 || i == n - 1) r = 59 * h / 48;
            else if (i == 3 || i == n - 2) r = 43 * h / 48;
            else if (i == 4 || i == n - 3) r = 49 * h / 48;
            else r = h;
        }
    } else if (rule == IB200_RULE_TRAPEZOID) {                  // trapezoidal.jl:35-40
        if (i == 1) r = (x[i0 + 1] - x[i0]) / 2;
        else if (i == n) r = (x[i0] - x[i0 - 1]) / 2;
        else r = (x[i0 + 1] - x[i0 - 1]) / 2;
    } else {                                                    // simpsons.jl:40-82
#define B(k) x[(k)]
#define E(k) x[n - 1 - (k)]
        const int64_t j = i - 1;
        auto cube = [](T v) { return v * v * v; };
        if (i == 1) {
            r = (B(2) - B(0)) / 6 * (2 - (B(2) - B(1)) / (B(1) - B(0)));
        } else if ((n & 1) && i == n) {
            r = (E(0) - E(2)) / 6 * (2 - (E(1) - E(2)) / (E(0) - E(1)));
        } else if (!(n & 1) && i == n) {
            r = (E(0) - E(1)) * (2 * (E(0) - E(1)) + 3 * (E(1) - E(2))) / (E(0) - E(2)) / 6;
        } else if (!(n & 1) && i == n - 1) {
            r = (E(1) - E(3)) / 6 * (2 - (E(2) - E(3)) / (E(1) - E(2))) +
                (E(0) - E(1)) * ((E(0) - E(1)) + 3 * (E(1) - E(2))) / (E(1) - E(2)) / 6;
        } else if (!(n & 1) && i == n - 2) {
            r = cube(E(1) - E(3)) / (E(2) - E(3)) / (E(1) - E(2)) / 6 -
                cube(E(0) - E(1)) / (E(1) - E(2)) / (E(0) - E(2)) / 6;
        } else if (j & 1) {
            r = cube(B(j + 1) - B(j - 1)) / (B(j) - B(j - 1)) / (B(j + 1) - B(j)) / 6;
        } else {
            r = (B(j) - B(j - 2)) / 6 * (2 - (B(j - 1) - B(j - 2)) / (B(j) - B(j - 1))) +
                (B(j + 2) - B(j)) / 6 * (2 - (B(j + 2) - B(j + 1)) / (B(j + 1) - B(j)));
        }
#undef B
#undef E
    }
    w[l0] = r;
}

__device__ __forceinline__ double2 ldg_stream2(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream1(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_stream2(const float *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// inner >= 2.  Block = TX x TY threads (TX*TY == kThreads): tx walks the contiguous output axis in
// VEC-wide pieces, ty strides the columns of this slab.  grid = (row tiles, slabs, outer).
// T = double: the BASELINE path.  T = float (Float32 problems, test/interface_tests.jl:524-539): weights and products in
// float like the reference, accumulation in double (the reference's sequential Float32 sum is less accurate, not more).
template <typename T, int VEC, int UNROLL>
__global__ void __launch_bounds__(kThreads)
strided_kernel(const T *__restrict__ y, const T *__restrict__ w, double *__restrict__ partial,
               int64_t inner, int64_t n, int tx_count, int64_t cols_per_slab) {
    using T2 = typename Vec2<T>::type;
    extern __shared__ double sm[];
    const int ty_count = kThreads / tx_count;
    const int tx = threadIdx.x % tx_count, ty = threadIdx.x / tx_count;
    const int64_t j = ((int64_t)blockIdx.x * tx_count + tx) * VEC;
    const int64_t slab = blockIdx.y, o = blockIdx.z, nslab = gridDim.y;
    const int64_t c0 = slab * cols_per_slab;
    const int64_t c1 = (c0 + cols_per_slab < n) ? c0 + cols_per_slab : n;
    const T *yo = y + o * n * inner;
    double a0 = 0.0, a1 = 0.0;
    const bool act = j < inner;
    const bool act1 = (VEC == 2) && (j + 1 < inner);
    if (act) {
        int64_t i = c0 + ty;
        const int64_t step = ty_count;
        for (; i + (UNROLL - 1) * step < c1; i += UNROLL * step) {
            T2 v[UNROLL]; T ww[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const T *p = yo + (i + u * step) * inner + j;
                if (VEC == 2) v[u] = ldg_stream2(p); else { v[u].x = ldg_stream1(p); v[u].y = 0; }
                ww[u] = __ldg(w + i + u * step);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                a0 = a0 + ww[u] * v[u].x;
                if (VEC == 2) a1 = a1 + ww[u] * v[u].y;
            }
        }
        for (; i < c1; i += step) {
            const T *p = yo + i * inner + j;
            const T wi = __ldg(w + i);
            if (VEC == 2) { const T2 v = ldg_stream2(p); a0 = a0 + wi * v.x; a1 = a1 + wi * v.y; }
            else a0 = a0 + wi * ldg_stream1(p);
        }
    }
    if (ty_count > 1) {   // fixed-order combine across ty
        sm[(ty * tx_count + tx) * 2 + 0] = a0;
        sm[(ty * tx_count + tx) * 2 + 1] = a1;
        __syncthreads();
        if (ty == 0) {
            for (int t = 1; t < ty_count; t++) { a0 += sm[(t * tx_count + tx) * 2]; a1 += sm[(t * tx_count + tx) * 2 + 1]; }
        }
    }
    if (ty == 0 && act) {
        double *po = partial + (o * nslab + slab) * inner + j;
        po[0] = a0;
        if (act1) po[1] = a1;
    }
}

// inner == 1: grid = (slabs, rows); a CTA streams its slab of one row.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(kThreads)
contiguous_kernel(const T *__restrict__ y, const T *__restrict__ w, double *__restrict__ partial,
                  int64_t n, int64_t cols_per_slab, int vec_ok) {
    using T2 = typename Vec2<T>::type;
    __shared__ double red[kThreads / 32];
    const int64_t slab = blockIdx.x, o = blockIdx.y, nslab = gridDim.x;
    const int64_t c0 = slab * cols_per_slab;
    const int64_t c1 = (c0 + cols_per_slab < n) ? c0 + cols_per_slab : n;
    const T *yo = y + o * n;
    double a0 = 0.0, a1 = 0.0;
    if (vec_ok) {   // c0 even and row base aligned to the 2-element vector
        int64_t i = c0 + 2 * threadIdx.x;
        const int64_t step = 2 * kThreads;
        for (; i + (UNROLL - 1) * step + 1 < c1; i += UNROLL * step) {
            T2 v[UNROLL], ww[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                v[u] = ldg_stream2(yo + i + u * step);
                ww[u] = *reinterpret_cast<const T2 *>(w + i + u * step);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) { a0 = a0 + ww[u].x * v[u].x; a1 = a1 + ww[u].y * v[u].y; }
        }
        for (; i < c1; i += step) {
            a0 = a0 + w[i] * ldg_stream1(yo + i);
            if (i + 1 < c1) a1 = a1 + w[i + 1] * ldg_stream1(yo + i + 1);
        }
    } else {
        for (int64_t i = c0 + threadIdx.x; i < c1; i += kThreads) a0 = a0 + w[i] * ldg_stream1(yo + i);
    }
    double a = a0 + a1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = red[0];
#pragma unroll
        for (int k = 1; k < kThreads / 32; k++) s += red[k];
        partial[o * nslab + slab] = s;
    }
}

// out[o, j] = sum over slabs in increasing order
template <typename T>
__global__ void finish_kernel(const double *__restrict__ partial, T *__restrict__ out,
                              int64_t inner, int64_t nslab, int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t o = idx / inner, j = idx - o * inner;
    const double *p = partial + o * nslab * inner + j;
    double s = p[0];
    for (int64_t k = 1; k < nslab; k++) s += p[k * inner];
    out[idx] = (T)s;
}

}  // namespace

// all-reduce (sum) of a device buffer over the ranks of the context's communicator (hcub_single.cu owns the NCCL binding)
int32_t ib200_allreduce_sum_f64(ib200_ctx *ctx, double *dbuf, size_t count);

// y holds the points off .. off + n - 1 of a grid of n_total points (a slab of the integration axis; the whole grid when
// off == 0 and n == n_total); x, if given, is the WHOLE grid.
template <typename T>
static int32_t sampled_impl(ib200_ctx *ctx, int32_t rule, const T *y, int64_t inner, int64_t n,
                            int64_t outer, T h, const T *x, T *out, int64_t n_total, int64_t off, bool reduce_ranks) {
    constexpr bool kF64 = sizeof(T) == 8;
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_sampled: ctx is NULL");
    IB_REQUIRE(ctx, rule == IB200_RULE_TRAPEZOID || rule == IB200_RULE_SIMPSON, "ib200_sampled: unknown rule");
    IB_REQUIRE(ctx, inner >= 0 && outer >= 0 && n >= 0, "ib200_sampled: negative extent");
    IB_REQUIRE(ctx, n_total != 0, "No points to integrate");                            // sampled.jl:55,72,106
    IB_REQUIRE(ctx, off >= 0 && off + n <= n_total, "ib200_sampled_slab: the slab does not lie inside the grid");
    if (x) {
        if (rule == IB200_RULE_SIMPSON)
            IB_REQUIRE(ctx, n_total > 2, "The length of the grid must exceed 2 for simpsons rule.");  // simpsons.jl:46
        else
            IB_REQUIRE(ctx, n_total >= 2, "ib200_sampled: a non-uniform trapezoid grid needs at least 2 points");
    }
    if (inner * outer == 0) return IB200_OK;
    IB_REQUIRE(ctx, (y != nullptr || n == 0) && out != nullptr, "ib200_sampled: y/out is NULL");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));

    DevIn<T> dy, dx; DevOut<T> dout;
    int32_t rc;
    if ((rc = dy.init(ctx, y, (size_t)(inner * n * outer), SL_IN0))) return rc;
    if ((rc = dx.init(ctx, x, x ? (size_t)n_total : 0, SL_IN1))) return rc;
    if ((rc = dout.init(ctx, out, (size_t)(inner * outer), SL_OUT0))) return rc;
    T *w = (T *)ctx->get(SL_WORK0, (size_t)(n > 0 ? n : 1) * sizeof(T));
    if (!w) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_sampled: weight scratch allocation failed");

    ib200_time_begin(ctx);
    if (n == 0) {                                               // an empty slab contributes nothing
        IB_CUDA(ctx, cudaMemsetAsync(dout.d, 0, (size_t)(inner * outer) * sizeof(T), ctx->stream));
    } else {
    weights_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(rule, n_total, h, dx.d, w, off, n);
    IB_LAUNCH_CHECK(ctx);

    const int64_t target_ctas = (int64_t)ctx->sm_count * 8;   // 8 x 256-thread CTAs resident per SM
    if (inner >= 2) {
        const bool vec2 = (inner % 2 == 0) && (((uintptr_t)dy.d) % (2 * sizeof(T)) == 0);
        const int VEC = vec2 ? 2 : 1;
        const int64_t pieces = (inner + VEC - 1) / VEC;
        int tx_count = 1;
        while (tx_count < kThreads && tx_count < pieces) tx_count <<= 1;
        const int ty_count = kThreads / tx_count;
        const int64_t row_tiles = (pieces + tx_count - 1) / tx_count;
        int64_t nslab = ctx->opt_sampled_slabs > 0 ? ctx->opt_sampled_slabs
                                                   : (4 * target_ctas + row_tiles * outer - 1) / (row_tiles * outer);
        const int64_t min_cols = (int64_t)ty_count * 16;       // keep slabs long enough to amortise the tail
        if (nslab > (n + min_cols - 1) / min_cols) nslab = (n + min_cols - 1) / min_cols;
        if (nslab < 1) nslab = 1;
        if (nslab > 65535) nslab = 65535;
        const int64_t cols_per_slab = (n + nslab - 1) / nslab;
        nslab = (n + cols_per_slab - 1) / cols_per_slab;
        const bool direct = kF64 && nslab == 1;                 // one slab of doubles: the kernel writes the result itself
        double *partial = direct ? reinterpret_cast<double *>(dout.d)
                                 : (double *)ctx->get(SL_WORK1, (size_t)(outer * nslab * inner) * sizeof(double));
        if (!partial) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_sampled: partial-sum scratch allocation failed");
        const size_t smem = ty_count > 1 ? (size_t)kThreads * 2 * sizeof(double) : 0;
        for (int64_t o0 = 0; o0 < outer; o0 += 65535) {          // gridDim.z <= 65535: the outer extent goes in chunks (any shape, like the reference)
            const int64_t oc = outer - o0 < 65535 ? outer - o0 : 65535;
            dim3 grid((unsigned)row_tiles, (unsigned)nslab, (unsigned)oc);
            const T *yc = dy.d + o0 * n * inner;
            double *pc = partial + o0 * nslab * inner;
            if (vec2) strided_kernel<T, 2, 8><<<grid, kThreads, smem, ctx->stream>>>(yc, w, pc, inner, n, tx_count, cols_per_slab);
            else      strided_kernel<T, 1, 8><<<grid, kThreads, smem, ctx->stream>>>(yc, w, pc, inner, n, tx_count, cols_per_slab);
            IB_LAUNCH_CHECK(ctx);
        }
        if (!direct) {
            const int64_t total = inner * outer;
            finish_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(partial, dout.d, inner, nslab, total);
            IB_LAUNCH_CHECK(ctx);
        }
    } else {
        int64_t nslab = ctx->opt_sampled_slabs > 0 ? ctx->opt_sampled_slabs : (2 * target_ctas + outer - 1) / outer;
        const int64_t min_cols = 2 * kThreads * 8;
        if (nslab > (n + min_cols - 1) / min_cols) nslab = (n + min_cols - 1) / min_cols;
        if (nslab < 1) nslab = 1;
        int64_t cols_per_slab = (n + nslab - 1) / nslab;
        cols_per_slab = (cols_per_slab + 1) & ~(int64_t)1;      // even, so slabs stay 16-byte aligned
        nslab = (n + cols_per_slab - 1) / cols_per_slab;
        const int vec_ok = (((uintptr_t)dy.d) % (2 * sizeof(T)) == 0) && (n % 2 == 0 || outer == 1) && (((uintptr_t)w) % 16 == 0);
        const bool direct = kF64 && nslab == 1;
        double *partial = direct ? reinterpret_cast<double *>(dout.d) : (double *)ctx->get(SL_WORK1, (size_t)(outer * nslab) * sizeof(double));
        if (!partial) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_sampled: partial-sum scratch allocation failed");
        for (int64_t o0 = 0; o0 < outer; o0 += 65535) {          // gridDim.y <= 65535: rows in chunks
            const int64_t oc = outer - o0 < 65535 ? outer - o0 : 65535;
            dim3 grid((unsigned)nslab, (unsigned)oc);
            contiguous_kernel<T, 4><<<grid, kThreads, 0, ctx->stream>>>(dy.d + o0 * n, w, partial + o0 * nslab, n, cols_per_slab,
                                                                       vec_ok && ((o0 * n) % 2 == 0));
            IB_LAUNCH_CHECK(ctx);
        }
        if (!direct) {
            finish_kernel<T><<<(unsigned)((outer + 255) / 256), 256, 0, ctx->stream>>>(partial, dout.d, 1, nslab, outer);
            IB_LAUNCH_CHECK(ctx);
        }
    }
    }   // n > 0
    if (reduce_ranks && ctx->comm) {
        if constexpr (kF64) {
            if ((rc = ib200_allreduce_sum_f64(ctx, reinterpret_cast<double *>(dout.d), (size_t)(inner * outer)))) return rc;
        }
    }
    ib200_time_end(ctx);
    if ((rc = dout.finish(ctx))) return rc;
    if (dout.is_host()) IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}

extern "C" int32_t ib200_sampled(ib200_ctx *ctx, int32_t rule, const double *y, int64_t inner, int64_t n,
                                 int64_t outer, double h, const double *x, double *out) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_sampled: ctx is NULL");
    IB_REQUIRE(ctx, n != 0, "No points to integrate");
    return sampled_impl<double>(ctx, rule, y, inner, n, outer, h, x, out, n, 0, false);
}

extern "C" int32_t ib200_sampled_f32(ib200_ctx *ctx, int32_t rule, const float *y, int64_t inner, int64_t n,
                                     int64_t outer, float h, const float *x, float *out) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_sampled: ctx is NULL");
    IB_REQUIRE(ctx, n != 0, "No points to integrate");
    return sampled_impl<float>(ctx, rule, y, inner, n, outer, h, x, out, n, 0, false);
}

extern "C" int32_t ib200_sampled_slab(ib200_ctx *ctx, int32_t rule, const double *y_slab, int64_t inner, int64_t n_slab,
                                      int64_t outer, int64_t n_total, int64_t offset, double h, const double *x, double *out) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_sampled_slab: ctx is NULL");
    return sampled_impl<double>(ctx, rule, y_slab, inner, n_slab, outer, h, x, out, n_total, offset, true);
}
