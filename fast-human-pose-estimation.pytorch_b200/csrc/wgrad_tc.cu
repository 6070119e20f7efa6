// wgrad_tc.cu -- placeholder until the tcgen05 weight-gradient GEMM lands (see DESIGN.md roadmap).
#include "common.cuh"
#include "kernels.h"
namespace fpd {
bool wgrad_tc_supported(int, int, int) { return false; }
size_t wgrad_tc_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
int wgrad_tc_launch(const float*, const float*, const float*, const float*, float*, float, int, int, int, int, int,
                    int, void*, size_t, int, cudaStream_t) {
  set_last_error("wgrad_tc: not built");
  return FPD_ERR_UNSUPPORTED;
}
}  // namespace fpd
