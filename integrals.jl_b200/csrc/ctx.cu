// ctx.cu -- context management and the FP64 peak microbenchmark of libintegrals_b200.so
#include "common.cuh"
#include "family.cuh"

thread_local std::string g_ib200_err;

extern "C" int32_t ib200_version(void) { return IB200_VERSION; }

extern "C" int32_t ib200_create(int32_t device, ib200_ctx **out) {
    if (!out) return ib200_fail(nullptr, IB200_ERR_INVALID, "ib200_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return ib200_fail(nullptr, IB200_ERR_CUDA,
                          std::string("ib200_create: no CUDA device available (") + cudaGetErrorString(e) +
                              "); libintegrals_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev)
        return ib200_fail(nullptr, IB200_ERR_INVALID, "ib200_create: device ordinal out of range");
    IB_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    IB_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return ib200_fail(nullptr, IB200_ERR_UNSUPPORTED,
                          "ib200_create: this library is built for sm_100a (B200) only; found compute capability " +
                              std::to_string(prop.major) + "." + std::to_string(prop.minor));
    ib200_ctx *ctx = new ib200_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    // a BLOCKING stream: work the caller enqueued on the legacy default stream (torch's default, a plain CuArray
    // computation) is ordered before the engine's kernels without an explicit synchronisation; callers on other
    // streams use ib200_set_stream
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamDefault) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return ib200_fail(nullptr, IB200_ERR_CUDA, "ib200_create: stream/event creation failed");
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return IB200_OK;
}

extern "C" void ib200_destroy(ib200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ib200_comm_destroy(ctx);
    if (ctx->batch && ctx->batch_free) { ctx->batch_free(ctx->batch); ctx->batch = nullptr; }
    for (auto &b : ctx->scratch) if (b.p) cudaFree(b.p);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char *ib200_last_error(ib200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_ib200_err.c_str(); }

extern "C" int32_t ib200_set_stream(ib200_ctx *ctx, void *s) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_set_stream: ctx is NULL");
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return IB200_OK;
}

extern "C" int32_t ib200_synchronize(ib200_ctx *ctx) {
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_synchronize: ctx is NULL");
    IB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IB200_OK;
}

extern "C" int64_t ib200_kernel_launches(ib200_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" double ib200_last_kernel_ms(ib200_ctx *ctx) {
    if (!ctx || !ctx->timing_open) return 0.0;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) { cudaGetLastError(); return 0.0; }
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) { cudaGetLastError(); return 0.0; }
    return (double)ms;
}

extern "C" int32_t ib200_family_nparam(int32_t family, int32_t dim) {
    if (family < 0 || family >= IB200_FAM_COUNT || !ib::fam_dim_ok(family, dim)) return -1;
    return ib::fam_nparam(family, dim);
}

extern "C" int32_t ib200_set_option(ib200_ctx *ctx, const char *key, int64_t value) {
    IB_REQUIRE(ctx, ctx != nullptr && key != nullptr, "ib200_set_option: NULL argument");
    std::string k(key);
    if (k == "adaptive_warps_per_sm") ctx->opt_adaptive_warps_per_sm = value;
    else if (k == "hcub_capacity") ctx->opt_hcub_capacity = value;
    else if (k == "sampled_slabs") ctx->opt_sampled_slabs = value;
    else if (k == "hcub_mode") ctx->opt_hcub_mode = value;
    else if (k == "hcub_lpt") ctx->opt_hcub_lpt = value;
    else if (k == "hcub_compact") ctx->opt_hcub_compact = value;
    else if (k == "hcub_compact_level") {
        if (value < 1 || value > 31) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_set_option: hcub_compact_level must be 1..31");
        ctx->opt_hcub_compact_level = value;
    }
    else if (k == "quadgk_mode") ctx->opt_quadgk_mode = value;
    else if (k == "single_round_percent") {                 // hcub_single.cu: 1..100, share of the boxes one round may bisect
        if (value < 1 || value > 100) return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_set_option: single_round_percent must be 1..100");
        ctx->opt_single_round_fraction = (double)value / 100.0;
    }
    else if (k == "single_switch_boxes") ctx->opt_single_switch_boxes = value < 0 ? 0 : value;
    else if (k == "single_switch_factor") ctx->opt_single_switch_factor = value < 0 ? 0 : value;
    else if (k == "single_arena_mib") ctx->opt_single_arena_mib = value < 0 ? 0 : value;
    else if (k == "single_sub_boxes") ctx->opt_single_sub_boxes = value < 256 ? 256 : (value > (1 << 22) ? (1 << 22) : value);
    else return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_set_option: unknown key " + k);
    return IB200_OK;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak: 8 independent DFMA chains per thread, 1024 threads x 2 CTAs per SM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chain live
}

extern "C" int32_t ib200_measure_fp64_peak(ib200_ctx *ctx, double *tflops, double *ms_out) {
    IB_REQUIRE(ctx, ctx != nullptr && tflops != nullptr, "ib200_measure_fp64_peak: NULL argument");
    IB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = 1024, blocks = ctx->sm_count * 2, iters = 4096;
    double *d = (double *)ctx->get(SL_WORK0, (size_t)threads * blocks * sizeof(double));
    if (!d) return ib200_fail(ctx, IB200_ERR_OOM, "scratch allocation failed");
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        IB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
        IB_LAUNCH_CHECK(ctx);
        IB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        IB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        IB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return IB200_OK;
}
