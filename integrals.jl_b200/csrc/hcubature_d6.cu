// hcubature_d6.cu -- instantiates the adaptive cubature kernels for dimension 6 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<6>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<6>(ctx, family, allfin, p); }
}
