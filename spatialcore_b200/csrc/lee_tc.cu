// tcgen05 3xTF32 Lee's L contraction (impl 2).  Placeholder until the tensor-core kernel lands:
// reports "unsupported" so sc_lee_gemm(impl=0) uses the CUDA-core kernel.
#include "lee.cuh"

namespace sc {
bool lee_tc_supported(int64_t, int, int64_t, int64_t) { return false; }
size_t lee_tc_extra_workspace_bytes(int64_t, int) { return 0; }
int lee_tc_launch(const float*, int64_t, const float*, int64_t, int64_t, int, const LeePlan&, double*,
                  void*, cudaStream_t) {
  set_error("tcgen05 Lee kernel not built");
  return SC_ERR_UNSUPPORTED;
}
}  // namespace sc
