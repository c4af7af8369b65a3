// hcubature_d1.cu -- instantiates the adaptive cubature kernels for dimension 1 (see hcubature_impl.cuh)
#include "hcubature_impl.cuh"
namespace hc {
template <> int32_t launch<1>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<1>(ctx, family, allfin, p); }
}
