// common.cuh -- context, error plumbing and host/device pointer staging for libintegrals_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include "../../include/integrals_b200.h"

#define IB200_VERSION 100

enum ScratchSlot { SL_IN0 = 0, SL_IN1, SL_IN2, SL_IN3, SL_IN4, SL_IN5, SL_OUT0, SL_OUT1, SL_OUT2, SL_OUT3,
                   SL_WORK0, SL_WORK1, SL_WORK2, SL_WORK3, SL_WORK4, SL_WORK5,
                   SL_STORE0, SL_STORE1,     // hcub_single.cu box store: exported to the peer ranks (CUDA IPC), never shared with other calls
                   SL_ARENA0, SL_ARENA1, SL_ARENA2,   // hcubature_rb.cuh per-warp arenas of the level-2 sub-problems of ib200_hcubature_single
                   SL_COUNT };

struct ib200_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing_open = false;
    double last_ms = 0.0;
    int64_t launches = 0;
    std::string err;
    void *comm = nullptr;                    // ncclComm_t of the multi-GPU single-domain solve (hcub_single.cu)
    int comm_rank = 0, comm_world = 1;
    void *single_peers = nullptr;            // hcub_single.cu: CUDA IPC mappings of the peers' box stores
    void *batch = nullptr;                   // batch_bridge.cu: state of the open caller-evaluated solve
    void (*batch_free)(void *) = nullptr;
    struct Buf { void *p = nullptr; size_t cap = 0; } scratch[SL_COUNT];
    // options
    int64_t opt_adaptive_warps_per_sm = 0;   // 0 = kernel default
    int64_t opt_hcub_lpt = 1;                // pair-mode HCubature: pilot pass + longest-expected-first ordering
    double opt_single_round_fraction = 0.5;  // hcub_single.cu: share of the stored boxes one round may bisect
    int64_t opt_single_switch_boxes = 1 << 20;   // hcub_single.cu: level-1 store size at which a solve goes two-level (0 = never)
    int64_t opt_single_switch_factor = 0;        // ... provided the error is within this factor of the tolerance (or the store is 64x that size);
                                                 // 0 = automatic: min(4096, (arena / 16)^(6/dim)) -- a sub-problem needs about (E/tol)^(dim/6) rule
                                                 // applications, and the average one should fill a sixteenth of its arena at most (hcub_single.cu)
    int64_t opt_single_arena_mib = 0;            // device memory for the level-2 arenas, MiB (0 = 70 % of what is free)
    int64_t opt_single_sub_boxes = 1 << 17;      // boxes a level-2 sub-problem may hold (its warp's arena)
    int64_t opt_hcub_capacity = 0;           // boxes per in-flight problem, 0 = derive from maxevals
    int64_t opt_sampled_slabs = 0;           // 0 = auto
    int64_t opt_quadgk_mode = 0;             // 0 = auto, 1 = warp per problem, 2 = pair mode (thread per segment), 3 = thread per problem
    int64_t opt_hcub_mode = 0;               // 0 = auto, 1 = warp per problem, 2 = pair mode (thread per box)
    int64_t opt_hcub_compact = 1;            // round-based driver: 32-byte dyadic box records where they apply (hcubature_rb.cuh kCptBytes); 0 = full records
    int64_t opt_hcub_compact_level = 31;     // halvings per axis a compact record may hold (tests lower it to exercise the hand-back)

    // grow-only device scratch (the cacheval analogue: reused across calls)
    void *get(ScratchSlot s, size_t bytes) {
        Buf &b = scratch[s];
        if (bytes == 0) bytes = 8;
        if (b.cap < bytes) {
            if (b.p) { cudaStreamSynchronize(stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
            size_t want = bytes + bytes / 8;
            if (cudaMalloc(&b.p, want) != cudaSuccess) {
                cudaGetLastError();
                if (cudaMalloc(&b.p, bytes) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return nullptr; }
                want = bytes;
            }
            b.cap = want;
        }
        return b.p;
    }
};

extern thread_local std::string g_ib200_err;  // for failures without a context

inline int32_t ib200_fail(ib200_ctx *ctx, int32_t code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_ib200_err = msg;
    return code;
}

#define IB_CUDA(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                    \
            return ib200_fail((ctx), e__ == cudaErrorMemoryAllocation ? IB200_ERR_OOM : IB200_ERR_CUDA, \
                              std::string(#call) + ": " + cudaGetErrorString(e__));                \
        }                                                                                          \
    } while (0)

#define IB_REQUIRE(ctx, cond, msg)                                                                 \
    do { if (!(cond)) return ib200_fail((ctx), IB200_ERR_INVALID, (msg)); } while (0)

inline bool ib200_is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Input argument: device pointer used in place; host pointer staged into a scratch slot.
template <typename T>
struct DevIn {
    const T *d = nullptr;
    int32_t init(ib200_ctx *ctx, const T *p, size_t count, ScratchSlot slot) {
        if (!p || count == 0) { d = nullptr; return 0; }
        if (ib200_is_device_ptr(p)) { d = p; return 0; }
        void *s = ctx->get(slot, count * sizeof(T));
        if (!s) return ib200_fail(ctx, IB200_ERR_OOM, "device scratch allocation failed (input staging)");
        IB_CUDA(ctx, cudaMemcpyAsync(s, p, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        d = (const T *)s;
        return 0;
    }
};

// Output argument: device pointer written in place; host pointer gets a scratch slot + copy-back.
template <typename T>
struct DevOut {
    T *d = nullptr; T *host = nullptr; size_t count = 0;
    int32_t init(ib200_ctx *ctx, T *p, size_t n, ScratchSlot slot, bool required_scratch = false) {
        count = n;
        if (!p) {
            if (!required_scratch) { d = nullptr; return 0; }
        } else if (ib200_is_device_ptr(p)) { d = p; return 0; }
        void *s = ctx->get(slot, n * sizeof(T));
        if (!s) return ib200_fail(ctx, IB200_ERR_OOM, "device scratch allocation failed (output staging)");
        d = (T *)s; host = p;
        return 0;
    }
    int32_t finish(ib200_ctx *ctx) {
        if (host && count) IB_CUDA(ctx, cudaMemcpyAsync(host, d, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        return 0;
    }
    bool is_host() const { return host != nullptr; }
};

// kernel timing on the launching stream (CUDA events), reported by ib200_last_kernel_ms
inline void ib200_time_begin(ib200_ctx *ctx) { cudaEventRecord(ctx->ev0, ctx->stream); ctx->timing_open = true; }
inline void ib200_time_end(ib200_ctx *ctx) { cudaEventRecord(ctx->ev1, ctx->stream); }

#define IB_LAUNCH_CHECK(ctx)                                                                       \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess)                                                                    \
            return ib200_fail((ctx), IB200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
    } while (0)

// family dispatch: calls fn(std::integral_constant<int,FAM>{}) for a runtime family id
template <typename Fn>
inline bool ib200_dispatch_family(int fam, Fn &&fn) {
    switch (fam) {
#define IBF(k) case k: fn(std::integral_constant<int, k>{}); return true;
        IBF(0) IBF(1) IBF(2) IBF(3) IBF(4) IBF(5) IBF(6) IBF(7) IBF(8) IBF(9) IBF(10)
#undef IBF
    default: return false;
    }
}

// dimension dispatch: calls fn(std::integral_constant<int,D>{}) for 1 <= dim <= MAXD
template <int MAXD, typename Fn>
inline bool ib200_dispatch_dim(int dim, Fn &&fn) {
    switch (dim) {
#define IBD(k) case k: if constexpr (k <= MAXD) { fn(std::integral_constant<int, k>{}); return true; } else return false;
        IBD(1) IBD(2) IBD(3) IBD(4) IBD(5) IBD(6) IBD(7) IBD(8)
#undef IBD
    default: return false;
    }
}

// Are all of the `count` bounds finite?  Host pointers are scanned; small device arrays are copied
// back; large device arrays return false (the general transform path handles finite axes too).
inline bool ib200_all_finite(ib200_ctx *ctx, const double *lb, const double *ub, size_t count) {
    auto scan = [](const double *p, size_t n) { for (size_t i = 0; i < n; i++) if (!(p[i] - p[i] == 0.0)) return false; return true; };
    const double *ptrs[2] = {lb, ub};
    for (const double *p : ptrs) {
        if (!ib200_is_device_ptr(p)) { if (!scan(p, count)) return false; continue; }
        if (count > 4096) return false;
        std::vector<double> h(count);
        if (cudaMemcpyAsync(h.data(), p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return false; }
        if (!scan(h.data(), count)) return false;
    }
    return true;
}
