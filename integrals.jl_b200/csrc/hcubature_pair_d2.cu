// hcubature_pair_d2.cu -- instantiates the pair-mode adaptive cubature kernels for dimension 2 (see hcubature_pair.cuh)
#include "hcubature_pair.cuh"
namespace hp {
template <> int32_t launch<2>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<2>(ctx, family, allfin, p); }
}
