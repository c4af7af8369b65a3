// hcubature_d7.cu -- instantiates the adaptive cubature kernels for dimension 7 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<7>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<7>(ctx, family, allfin, p); }
}
