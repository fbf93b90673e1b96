// kernels_d2q9.cu -- the D2Q9 instantiation of the step kernels (kernels_impl.cuh).
#include "kernels_impl.cuh"

namespace luma {
LUMA_INST(D2Q9)
}  // namespace luma
