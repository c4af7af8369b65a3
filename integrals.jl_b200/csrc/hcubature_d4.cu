// hcubature_d4.cu -- instantiates the adaptive cubature kernels for dimension 4 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<4>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<4>(ctx, family, allfin, p); }
}
