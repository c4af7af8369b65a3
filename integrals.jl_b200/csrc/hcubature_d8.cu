// hcubature_d8.cu -- instantiates the adaptive cubature kernels for dimension 8 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<8>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<8>(ctx, family, allfin, p); }
}
