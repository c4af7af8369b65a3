// hcubature_d5.cu -- instantiates the adaptive cubature kernels for dimension 5 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<5>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<5>(ctx, family, allfin, p); }
}
