// hcubature_pair_d4.cu -- instantiates the pair-mode adaptive cubature kernels for dimension 4 (see hcubature_pair.cuh)
#include "hcubature_pair.cuh"
namespace hp {
template <> int32_t launch<4>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<4>(ctx, family, allfin, p); }
}
