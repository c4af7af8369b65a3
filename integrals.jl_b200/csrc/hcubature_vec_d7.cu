// hcubature_vec_d7.cu -- the 7-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<7>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<7>(ctx, family, p); }
}
