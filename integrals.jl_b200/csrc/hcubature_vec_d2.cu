// hcubature_vec_d2.cu -- the 2-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<2>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<2>(ctx, family, p); }
}
