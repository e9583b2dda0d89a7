// tcgen05 TF32 implicit-GEMM back end -- placeholder until the tensor-core kernels land.
#include "igemm.cuh"

namespace bre {
bool igemm_tc_supported(const GemmArgs&) { return false; }
int launch_igemm_tc(const GemmArgs&, cudaStream_t) {
  set_error("tcgen05 back end not available for this shape");
  return -4;
}
}  // namespace bre
