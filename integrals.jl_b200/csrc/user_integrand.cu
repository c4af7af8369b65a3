// user_integrand.cu -- device-resident USER integrands (second half): an integrand given as CUDA source is compiled at
// run time (NVRTC) into one elementwise kernel y[j] = f(x[:, j], p) and run by the engine between the two kernels of the
// BatchIntegralFunction bridge (csrc/batch_bridge.cu: points of a round -> values -> weighted sums, error estimates, refinement), so a
// solve with an arbitrary integrand makes no host round trip for its evaluations -- only the bridge's one state read-back per round.
//
// Replaces, for integrands outside the registered families, what the reference does with a Julia closure: QuadGK / HCubature call
// `f(x, p)` point by point on the host (src/Integrals.jl:327-335, 380-393), or on a batch of points for a BatchIntegralFunction
// (src/Integrals.jl:274-315).  CUDA cannot run a Julia closure; the caller states the integrand in CUDA C++ instead:
//
//     body of   __device__ double f(const double *x, const double *p)      e.g.  "return exp(-p[0] * x[0] * x[0]) * cos(p[1] * x[1]);"
//
// Refinement, rule, change of variables and stopping test are the bridge's (identical rounds to orc_hcubature_rounds with the same
// integrand).  NVRTC and the driver API are loaded with dlopen when first used: the library has no link-time dependency on either.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

extern "C" int32_t ib200_batch_open(ib200_ctx *ctx, int32_t dim, int32_t transform, const double *lb, const double *ub,
                                    double reltol, double abstol, int64_t max_boxes, int64_t store_boxes);
extern "C" int32_t ib200_batch_next(ib200_ctx *ctx, int64_t *npoints);
extern "C" int32_t ib200_batch_points(ib200_ctx *ctx, double *x, int64_t npoints);
extern "C" int32_t ib200_batch_values(ib200_ctx *ctx, const double *y, int64_t npoints);
extern "C" int32_t ib200_batch_result(ib200_ctx *ctx, double *out_I, double *out_E, int64_t *out_stats);

namespace ui {

typedef struct _nvrtcProgram *nvrtcProgram;
typedef struct CUmod_st *CUmodule;
typedef struct CUfunc_st *CUfunction;
typedef struct CUstream_st *CUstream;

struct Api {
    void *hn = nullptr, *hc = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    int (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    int (*ModuleUnload)(CUmodule) = nullptr;
    int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;

    bool load_nvrtc(std::string &why) {
        if (hn) return true;
        const char *names[] = {getenv("IB200_NVRTC"), "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (const char *n : names) { if (n && *n && (hn = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (!hn) { why = "libnvrtc not found (set IB200_NVRTC to its path)"; return false; }
#define UI_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(hn, name)); if (!field) { why = std::string("libnvrtc lacks ") + name; hn = nullptr; return false; }
        UI_SYM(CreateProgram, "nvrtcCreateProgram") UI_SYM(CompileProgram, "nvrtcCompileProgram") UI_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        UI_SYM(GetProgramLog, "nvrtcGetProgramLog") UI_SYM(GetCUBINSize, "nvrtcGetCUBINSize") UI_SYM(GetCUBIN, "nvrtcGetCUBIN")
        UI_SYM(DestroyProgram, "nvrtcDestroyProgram") UI_SYM(GetErrorString, "nvrtcGetErrorString")
#undef UI_SYM
        return true;
    }
    bool load_driver(std::string &why) {
        if (hc) return true;
        hc = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!hc) { why = "libcuda.so.1 not found"; return false; }
#define UI_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(hc, name)); if (!field) { why = std::string("libcuda lacks ") + name; hc = nullptr; return false; }
        UI_SYM(ModuleLoadData, "cuModuleLoadData") UI_SYM(ModuleGetFunction, "cuModuleGetFunction") UI_SYM(ModuleUnload, "cuModuleUnload")
        UI_SYM(LaunchKernel, "cuLaunchKernel")
#undef UI_SYM
        return true;
    }
};
static Api g_api;
static std::mutex g_mu;

// a compiled integrand: the cubin (kept, so that every context / device loads its own module) and the modules loaded so far
struct Integrand {
    int dim = 0;
    std::string cubin;
    std::vector<std::pair<int, std::pair<CUmodule, CUfunction>>> loaded;      // device -> module, function
    bool live = false;
};
static std::vector<Integrand> g_integrands;

static std::string make_source(const char *body, int dim) {
    std::string s;
    s += "__device__ __forceinline__ double ib200_user_f(const double *x, const double *p) {\n";
    s += body;
    s += "\n}\n";
    s += "extern \"C\" __global__ void ib200_user_eval(const double *__restrict__ x, const double *__restrict__ p, long long n, double *__restrict__ y) {\n"
         "    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n"
         "    if (j >= n) return;\n"
         "    double xj[" + std::to_string(dim) + "];\n"
         "    for (int k = 0; k < " + std::to_string(dim) + "; k++) xj[k] = x[j * " + std::to_string(dim) + " + k];\n"
         "    y[j] = ib200_user_f(xj, p);\n"
         "}\n";
    return s;
}

}  // namespace ui

// Compile `body` (the body of  __device__ double f(const double *x, const double *p)) for sm_100a.  Needs no GPU: the handle can be made on
// any machine with NVRTC; the module is loaded on a device when a solve first uses it there.
extern "C" int32_t ib200_integrand_compile(ib200_ctx *ctx, const char *body, int32_t dim, int32_t *handle_out) {
    using namespace ui;
    IB_REQUIRE(ctx, body != nullptr && handle_out != nullptr, "ib200_integrand_compile: NULL argument");
    IB_REQUIRE(ctx, dim >= 1 && dim <= IB200_MAX_DIM, "ib200_integrand_compile: 1 <= dim <= 8");
    std::lock_guard<std::mutex> lock(g_mu);
    std::string why;
    if (!g_api.load_nvrtc(why)) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_integrand_compile: " + why);
    const std::string src = make_source(body, dim);
    nvrtcProgram prog = nullptr;
    int rc = g_api.CreateProgram(&prog, src.c_str(), "ib200_user_integrand.cu", 0, nullptr, nullptr);
    if (rc != 0) return ib200_fail(ctx, IB200_ERR_CUDA, std::string("nvrtcCreateProgram: ") + g_api.GetErrorString(rc));
    // -fmad=false: a*b+c stays a multiply and an add, as in the CPU restatement the tests compare with (the caller writes fma() to fuse)
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-default-device"};
    rc = g_api.CompileProgram(prog, 4, opts);
    if (rc != 0) {
        size_t n = 0; g_api.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_api.GetProgramLog(prog, &log[0]);
        g_api.DestroyProgram(&prog);
        return ib200_fail(ctx, IB200_ERR_INVALID, "ib200_integrand_compile: the integrand does not compile:\n" + log);
    }
    size_t nb = 0;
    g_api.GetCUBINSize(prog, &nb);
    Integrand it; it.dim = dim; it.cubin.resize(nb); it.live = true;
    rc = g_api.GetCUBIN(prog, &it.cubin[0]);
    g_api.DestroyProgram(&prog);
    if (rc != 0 || nb == 0) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_integrand_compile: no cubin produced");
    g_integrands.push_back(std::move(it));
    *handle_out = (int32_t)g_integrands.size() - 1;
    return IB200_OK;
}

extern "C" int32_t ib200_integrand_release(ib200_ctx *ctx, int32_t handle) {
    using namespace ui;
    std::lock_guard<std::mutex> lock(g_mu);
    IB_REQUIRE(ctx, handle >= 0 && handle < (int32_t)g_integrands.size() && g_integrands[handle].live, "ib200_integrand_release: unknown handle");
    Integrand &it = g_integrands[handle];
    for (auto &m : it.loaded) if (g_api.ModuleUnload) g_api.ModuleUnload(m.second.first);
    it.loaded.clear(); it.cubin.clear(); it.live = false;
    return IB200_OK;
}

// One integral of the compiled integrand: the rounds of ib200_hcubature_single / the bridge (Genz-Malik 7/5 per box, GK(7,15) in 1-D),
// every evaluation on the device.  params: nparam doubles (host or device), handed to f as `p`.  out_stats as ib200_batch_result.
extern "C" int32_t ib200_integrate_compiled(ib200_ctx *ctx, int32_t handle, int32_t dim, int32_t transform,
                                            const double *lb, const double *ub, const double *params, int32_t nparam,
                                            double reltol, double abstol, int64_t max_boxes, int64_t store_boxes,
                                            double *out_I, double *out_E, int64_t *out_stats) {
    using namespace ui;
    IB_REQUIRE(ctx, ctx != nullptr, "ib200_integrate_compiled: ctx is NULL");
    IB_REQUIRE(ctx, nparam >= 0 && (nparam == 0 || params != nullptr), "ib200_integrate_compiled: params missing");
    CUfunction fn = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        IB_REQUIRE(ctx, handle >= 0 && handle < (int32_t)g_integrands.size() && g_integrands[handle].live, "ib200_integrate_compiled: unknown integrand handle");
        Integrand &it = g_integrands[handle];
        IB_REQUIRE(ctx, it.dim == dim, "ib200_integrate_compiled: the integrand was compiled for another dimension");
        IB_CUDA(ctx, cudaSetDevice(ctx->device));
        IB_CUDA(ctx, cudaFree(0));                                       // the primary context is current for the driver calls below
        std::string why;
        if (!g_api.load_driver(why)) return ib200_fail(ctx, IB200_ERR_UNSUPPORTED, "ib200_integrate_compiled: " + why);
        for (auto &m : it.loaded) if (m.first == ctx->device) fn = m.second.second;
        if (!fn) {
            CUmodule mod = nullptr;
            int rc = g_api.ModuleLoadData(&mod, it.cubin.data());
            if (rc != 0) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_integrate_compiled: cuModuleLoadData failed (" + std::to_string(rc) + ")");
            rc = g_api.ModuleGetFunction(&fn, mod, "ib200_user_eval");
            if (rc != 0) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_integrate_compiled: cuModuleGetFunction failed (" + std::to_string(rc) + ")");
            it.loaded.push_back({ctx->device, {mod, fn}});
        }
    }
    DevIn<double> dpar;
    int32_t rc;
    const double zero = 0.0;
    if ((rc = dpar.init(ctx, nparam > 0 ? params : &zero, (size_t)(nparam > 0 ? nparam : 1), SL_IN2))) return rc;
    if ((rc = ib200_batch_open(ctx, dim, transform, lb, ub, reltol, abstol, max_boxes, store_boxes))) return rc;
    for (;;) {
        int64_t n = 0;
        if ((rc = ib200_batch_next(ctx, &n))) return rc;
        if (n == 0) break;
        double *x = (double *)ctx->get(SL_WORK0, (size_t)n * dim * sizeof(double));
        double *y = (double *)ctx->get(SL_WORK2, (size_t)n * sizeof(double));
        if (!x || !y) return ib200_fail(ctx, IB200_ERR_OOM, "ib200_integrate_compiled: point buffer allocation failed (lower max_boxes to bound the round size)");
        if ((rc = ib200_batch_points(ctx, x, n))) return rc;
        long long nn = (long long)n;
        const double *pp = dpar.d;
        void *args[] = {(void *)&x, (void *)&pp, (void *)&nn, (void *)&y};
        const unsigned block = 256, grid = (unsigned)((n + block - 1) / block);
        const int lr = g_api.LaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, (CUstream)ctx->stream, args, nullptr);
        if (lr != 0) return ib200_fail(ctx, IB200_ERR_CUDA, "ib200_integrate_compiled: cuLaunchKernel failed (" + std::to_string(lr) + ")");
        ctx->launches += 1;
        if ((rc = ib200_batch_values(ctx, y, n))) return rc;             // same stream: ordered after the evaluation kernel
    }
    return ib200_batch_result(ctx, out_I, out_E, out_stats);
}
