// tcgen05 (kind::tf32) versions of the two streaming products — placeholder until the kernels land.
#include "common.cuh"
namespace xb {
bool tc_supported(int64_t, int64_t, int64_t, const float*, int64_t) { return false; }
int64_t tc_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }
int project_S_tc(const float*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const float*,
                 int64_t, int64_t, float*, int64_t, void*, int64_t, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return XEOFS_E_UNSUPPORTED;
}
int project_T_tc(const float*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const float*,
                 int64_t, int64_t, float*, int64_t, void*, int64_t, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return XEOFS_E_UNSUPPORTED;
}
}  // namespace xb
