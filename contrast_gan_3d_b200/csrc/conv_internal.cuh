// Internal (non-ABI) entry points shared by conv_api.cu, conv_generic.cu and conv_tc.cu.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace cg {

int pack_weights(const float *w, void *packed, int dtype, int Cs, int Cb, int k, cudaStream_t st);
int generic_gather(const cgan3d_conv_geom &g, int dtype, const void *big, const void *wp, const float *bias,
                   void *small, cudaStream_t st);
int generic_scatter(const cgan3d_conv_geom &g, int dtype, const void *small, const void *wp, const float *bias,
                    void *big, void *ws, size_t ws_bytes, cudaStream_t st);
size_t generic_workspace_bytes(const cgan3d_conv_geom &g, int dtype, int op);
int generic_wgrad(const cgan3d_conv_geom &g, int dtype, const void *big, const void *small, float *dw, float beta,
                  cudaStream_t st);
int reflect_pad(const void *in, void *out, int dtype, int B, int X, int Y, int Z, int C, int pad, cudaStream_t st);
int reflect_pad_backward(const void *gp, void *gi, int dtype, int B, int X, int Y, int Z, int C, int pad,
                         cudaStream_t st);

// tcgen05 implicit-GEMM path (conv_tc.cu). op: 0 gather, 1 scatter, 2 wgrad.
bool tc_supported(const cgan3d_conv_geom &g, int dtype, int op);
size_t tc_workspace_bytes(const cgan3d_conv_geom &g, int dtype, int op);
// bn_sums (optional, fp64 [2*Cout], zeroed by the caller): the epilogue adds the per-channel sum / sum of squares of the
// fp32 accumulators (BatchNorm batch statistics fused into the convolution); see tc_fuses_bnstats.
int tc_gather(const cgan3d_conv_geom &g, const void *big, const void *wp, const float *bias, void *small,
              void *ws, size_t ws_bytes, cudaStream_t st, double *bn_sums = nullptr);
int tc_scatter(const cgan3d_conv_geom &g, const void *small, const void *wp, const float *bias, void *big,
               void *ws, size_t ws_bytes, cudaStream_t st, double *bn_sums = nullptr);
bool tc_fuses_bnstats(const cgan3d_conv_geom &g, int dtype, int op);
int tc_wgrad(const cgan3d_conv_geom &g, const void *big, const void *small, float *dw, float beta, void *ws,
             size_t ws_bytes, cudaStream_t st);

}  // namespace cg
