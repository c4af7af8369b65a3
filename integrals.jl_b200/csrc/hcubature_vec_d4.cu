// hcubature_vec_d4.cu -- the 4-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<4>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<4>(ctx, family, p); }
}
