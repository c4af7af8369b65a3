// hcubature_pair_d5.cu -- instantiates the pair-mode adaptive cubature kernels for dimension 5 (see hcubature_pair.cuh)
#include "hcubature_pair.cuh"
namespace hp {
template <> int32_t launch<5>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<5>(ctx, family, allfin, p); }
}
