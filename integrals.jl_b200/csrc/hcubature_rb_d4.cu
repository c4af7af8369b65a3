// hcubature_rb_d4.cu -- instantiates the round-based batched adaptive cubature kernels for dimension 4 (see hcubature_rb.cuh)
#include "hcubature_rb.cuh"
namespace hb {
template <> int32_t launch<4>(ib200_ctx *ctx, int family, bool allfin, Params &p, long long arena_mib) { return launch_impl<4>(ctx, family, allfin, p, arena_mib); }
}
