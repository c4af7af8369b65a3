// hcubature_vec_d8.cu -- the 8-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<8>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<8>(ctx, family, p); }
}
