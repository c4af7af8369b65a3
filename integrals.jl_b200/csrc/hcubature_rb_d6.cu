// hcubature_rb_d6.cu -- instantiates the round-based batched adaptive cubature kernels for dimension 6 (see hcubature_rb.cuh)
#include "hcubature_rb.cuh"
namespace hb {
template <> int32_t launch<6>(ib200_ctx *ctx, int family, bool allfin, Params &p, long long arena_mib) { return launch_impl<6>(ctx, family, allfin, p, arena_mib); }
}
