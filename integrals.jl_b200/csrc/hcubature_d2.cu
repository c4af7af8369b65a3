// hcubature_d2.cu -- instantiates the adaptive cubature kernels for dimension 2 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<2>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<2>(ctx, family, allfin, p); }
}
