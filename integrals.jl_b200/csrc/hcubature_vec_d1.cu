// hcubature_vec_d1.cu -- the 1-dimensional instantiations of the vector-valued cubature kernel (hcubature_vec.cuh)
#include "hcubature_vec.cuh"
namespace hv {
template <> int32_t launch<1>(ib200_ctx *ctx, int family, Params &p) { return launch_impl<1>(ctx, family, p); }
}
