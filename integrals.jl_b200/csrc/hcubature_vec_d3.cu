// hcubature_vec_d3.cu -- the 3-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<3>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<3>(ctx, family, p); }
}
