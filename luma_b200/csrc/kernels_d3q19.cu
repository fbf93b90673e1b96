// kernels_d3q19.cu -- the D3Q19 instantiation of the step kernels (kernels_impl.cuh).
#include "kernels_impl.cuh"

namespace luma {
LUMA_INST(D3Q19)
}  // namespace luma
