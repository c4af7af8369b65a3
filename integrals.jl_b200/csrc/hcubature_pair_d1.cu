// hcubature_pair_d1.cu -- instantiates the pair-mode kernels for dimension 1: QuadGK's (7,15) Gauss-Kronrod bisection (see hcubature_pair.cuh)
#include "hcubature_pair.cuh"
namespace hp {
template <> int32_t launch<1>(ib200_ctx *ctx, int family, bool allfin, Params &p) { return launch_impl<1>(ctx, family, allfin, p); }
}
