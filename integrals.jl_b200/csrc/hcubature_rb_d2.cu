// hcubature_rb_d2.cu -- instantiates the round-based batched adaptive cubature kernels for dimension 2 (see hcubature_rb.cuh)
#include "hcubature_rb.cuh"
namespace hb {
template <> int32_t launch<2>(ib200_ctx *ctx, int family, bool allfin, Params &p, long long arena_mib) { return launch_impl<2>(ctx, family, allfin, p, arena_mib); }
}
