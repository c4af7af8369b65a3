// hcubature_d3.cu -- instantiates the adaptive cubature kernels for dimension 3 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<3>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<3>(ctx, family, allfin, p); }
}
