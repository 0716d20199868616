#include "head_tc.cuh"

namespace dpd {

bool tc_supported(const dpd_head_config&) { return false; }
size_t tc_packed_bytes(const dpd_head_config&, int) { return 0; }
size_t tc_workspace_bytes(const dpd_head_config&, size_t) { return 0; }
int tc_pack_weights(const dpd_head_config&, int, const float*, const float*, const float*, void*, cudaStream_t) {
  return set_error(DPD_E_UNSUPPORTED, "tensor-core head not built");
}
int tc_head_layers(const dpd_head_config&, int, const GatherDesc&, int, const void*, const float*, const float*,
                   const float*, float*, float*, void*, const float**, cudaStream_t) {
  return set_error(DPD_E_UNSUPPORTED, "tensor-core head not built");
}

}  // namespace dpd
