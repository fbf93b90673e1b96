// kernels_d3q27.cu -- the D3Q27 instantiation of the step kernels (kernels_impl.cuh).
#include "kernels_impl.cuh"

namespace luma {
LUMA_INST(D3Q27)
}  // namespace luma
