// hcubature_vec_d6.cu -- the 6-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<6>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<6>(ctx, family, p); }
}
