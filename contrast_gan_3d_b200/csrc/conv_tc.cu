// tcgen05 implicit-GEMM convolution path (placeholder until the kernels land).
#include "common.cuh"
#include "conv_internal.cuh"

namespace cg {
bool tc_supported(const cgan3d_conv_geom &, int, int) { return false; }
size_t tc_workspace_bytes(const cgan3d_conv_geom &, int, int) { return 0; }
int tc_gather(const cgan3d_conv_geom &, const void *, const void *, const float *, void *, void *, size_t, cudaStream_t) {
  return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 gather not built");
}
int tc_scatter(const cgan3d_conv_geom &, const void *, const void *, const float *, void *, void *, size_t, cudaStream_t) {
  return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 scatter not built");
}
int tc_wgrad(const cgan3d_conv_geom &, const void *, const void *, float *, float, void *, size_t, cudaStream_t) {
  return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 wgrad not built");
}
}  // namespace cg
