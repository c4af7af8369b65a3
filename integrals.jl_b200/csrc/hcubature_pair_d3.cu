// hcubature_pair_d3.cu -- instantiates the pair-mode adaptive cubature kernels for dimension 3 (see hcubature_pair.cuh)
#include "hcubature_pair.cuh"
namespace hp {
template <> int32_t launch<3>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<3>(ctx, family, allfin, p); }
}
