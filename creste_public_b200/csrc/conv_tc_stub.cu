// placeholder until conv_tc.cu (tcgen05 implicit GEMM) lands
#include "common.cuh"
namespace creste {
int conv_tc_launch(const creste_conv_desc*, const float*, const float*, const float*, const float*,
                   const float*, const float*, float*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return CRESTE_ERR_ARG;
}
size_t conv_tc_workspace_bytes(const creste_conv_desc*) { return 0; }
bool conv_tc_supported(const creste_conv_desc*) { return false; }
}  // namespace creste
