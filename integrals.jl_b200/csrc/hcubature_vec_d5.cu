// hcubature_vec_d5.cu -- the 5-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<5>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<5>(ctx, family, p); }
}
