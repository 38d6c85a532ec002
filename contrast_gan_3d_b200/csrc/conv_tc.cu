// tcgen05 implicit-GEMM convolutions for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Kernel 1: conv_s1_tc_kernel — 3x3x3, stride 1, pad 1, Cin % 16 == 0, Cout % 16 == 0 (the 8 ResNet convs of the
// generator: 46 % of its FLOPs), used for fprop (gather) and for dgrad (scatter == gather with flipped, transposed
// filter).  Replaces aten::convolution / convolution_backward(input) at reference model/blocks.py:68-85.
//
// Design ("flattened-shift" implicit GEMM, no im2col materialisation, every input byte fetched from L2 once):
//   * One CTA owns an output slab (b, x0..x0+xlen, y0..y0+Yt, z0..z0+Zt) and marches along x.  For each input plane x
//     it TMA-loads the halo slab (Yt+2) x (Zt+2) x Cin ONCE, as Cin/8 boxes of 8 channels, into the UMMA
//     SWIZZLE_NONE K-major layout [Cin/8][row][8 ch] where row = yy*(Zt+2) + zz is the flattened halo position.
//     TMA's out-of-bounds zero fill implements the zero padding on all three axes.
//   * In that layout a filter tap (dx,dy,dz) is a pure ROW SHIFT of the A operand: the A descriptor for tap (dy,dz)
//     starts (dy*(Zt+2)+dz)*16 B further, plane dx selects one of three resident x-planes.  Output rows are the
//     flattened halo positions too; rows that fall on the halo columns are computed and discarded
//     (efficiency Yt*Zt / (mtiles*128)).
//   * Weights for one tap ([Cin/8][Cout][8] bf16, a K-major B operand) stream through a 4-stage ring with 1-D bulk
//     copies; each stage feeds mtiles*Cin/16 MMAs (M=128, N=Cout, K=16).
//   * Accumulators (mtiles x [128 x Cout] fp32) live in TMEM, double-buffered across output planes, so the
//     epilogue (tcgen05.ld -> bf16 -> 16 B global stores) of plane x overlaps the MMAs of plane x+1.
//   * Warp roles: 0-3 epilogue, 4 activation-plane TMA producer, 5 MMA issuer (+TMEM alloc), 6 weight producer.
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc_common.cuh"

namespace cg {

using bf16 = __nv_bfloat16;

struct TcPlan {
  int B, X, Y, Z, Cin, N;
  int Zt, nzt, Zh;
  int Yt, nslabs, Yh;
  int xseg, nxseg;
  int mtiles, rows_alloc, nitems, b_stages;
  uint32_t plane_bytes, btile_bytes, box_bytes, tmem_cols, smem_bytes;
};

constexpr int kPlaneSlots = 3;
constexpr int kThreads = 224;
constexpr uint32_t kSmemLimit = 232448 - 1024;

// Work = ncols columns (b, z-tile, y-slab) x X output planes, flattened as idx = col * X + x.  CTA c owns the contiguous
// range [c*total/grid, (c+1)*total/grid): at most one plane of imbalance; a range that crosses a column boundary is
// processed as two items (each item re-loads its two halo planes).
struct ItemIter {
  int idx, end;
  __device__ __forceinline__ ItemIter(const TcPlan &p) {
    const long long total = (long long)p.nitems * p.X;  // nitems == number of columns
    idx = (int)(total * blockIdx.x / gridDim.x);
    end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  }
  __device__ __forceinline__ bool next(const TcPlan &p, int &b, int &z0, int &zlen, int &y0, int &ylen, int &x0, int &xlen) {
    if (idx >= end) return false;
    int col = idx / p.X;
    x0 = idx - col * p.X;
    xlen = min(p.X - x0, end - idx);
    idx += xlen;
    const int sl = col % p.nslabs; col /= p.nslabs;
    const int zt = col % p.nzt;
    b = col / p.nzt;
    y0 = sl * p.Yt; ylen = min(p.Yt, p.Y - y0);
    z0 = zt * p.Zt; zlen = min(p.Zt, p.Z - z0);
    return true;
  }
};

template <int KSTEPS, int MT>
__global__ void __launch_bounds__(kThreads, 1)
conv_s1_tc_kernel(const __grid_constant__ CUtensorMap tmA, const bf16 *__restrict__ wB, bf16 *__restrict__ out, const TcPlan p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *planes = smem;
  uint8_t *bt = planes + (size_t)kPlaneSlots * p.plane_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(bt + (size_t)p.b_stages * p.btile_bytes);
  uint64_t *plane_full = bars, *plane_empty = bars + kPlaneSlots;
  uint64_t *b_full = bars + 2 * kPlaneSlots, *b_empty = b_full + p.b_stages;
  uint64_t *tm_full = b_empty + p.b_stages, *tm_empty = tm_full + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tm_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPlaneSlots; ++i) { tc::mbar_init(&plane_full[i], 1); tc::mbar_init(&plane_empty[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tm_full[i], 1); tc::mbar_init(&tm_empty[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 5) {
    tc::tmem_alloc(tmem_ptr, p.tmem_cols);
    tc::tmem_relinquish();
  }
  if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&tmA);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int kchunks8 = p.Cin >> 3;
  const int taps = 27;

  if (warp == 4) {
    // ------------------------------------------------ activation-plane producer
    if (lane == 0) {
      uint32_t e = 0;
      int b, z0, zlen, y0, ylen, x0, xlen;
      for (ItemIter it(p); it.next(p, b, z0, zlen, y0, ylen, x0, xlen);) {
        for (int px = x0 - 1; px <= x0 + xlen; ++px, ++e) {
          const uint32_t slot = e % kPlaneSlots, use = e / kPlaneSlots;
          if (use > 0) tc::mbar_wait(&plane_empty[slot], (use - 1) & 1);
          tc::mbar_expect_tx(&plane_full[slot], p.box_bytes * kchunks8);
          uint8_t *dst = planes + (size_t)slot * p.plane_bytes;
          for (int cc = 0; cc < kchunks8; ++cc)
            tc::tma_load_5d(dst + (size_t)cc * p.rows_alloc * 16, &tmA, &plane_full[slot], cc * 8, z0 - 1, y0 - 1, px, b);
        }
      }
    }
  } else if (warp == 6) {
    // ------------------------------------------------ weight-tap producer
    if (lane == 0) {
      uint32_t t = 0;
      int b, z0, zlen, y0, ylen, x0, xlen;
      for (ItemIter it(p); it.next(p, b, z0, zlen, y0, ylen, x0, xlen);) {
        for (int i = 0; i < xlen; ++i)
          for (int tap = 0; tap < taps; ++tap, ++t) {
            const uint32_t s = t % p.b_stages, use = t / p.b_stages;
            if (use > 0) tc::mbar_wait(&b_empty[s], (use - 1) & 1);
            tc::mbar_expect_tx(&b_full[s], p.btile_bytes);
            tc::bulk_g2s(bt + (size_t)s * p.btile_bytes, reinterpret_cast<const uint8_t *>(wB) + (size_t)tap * p.btile_bytes,
                         p.btile_bytes, &b_full[s]);
          }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer
    // The whole warp runs the (warp-uniform) control flow so that ptxas keeps descriptors, TMEM addresses and loop
    // state in UNIFORM registers; only the tcgen05.mma / tcgen05.commit themselves are issued by one elected lane.
    // (A lane-0-only loop forces per-MMA R2UR transfers and made the issue thread, not the tensor pipe, the limit.)
    {
      const bool leader = tc::elect_one();
      const uint32_t idesc = tc::make_idesc_bf16(128, p.N, 0, 0);
      const uint32_t planes_u32 = tc::smem_u32(planes), bt_u32 = tc::smem_u32(bt);
      const uint32_t a_lbo = (uint32_t)p.rows_alloc * 16, b_lbo = (uint32_t)p.N * 16;
      const uint64_t a_desc_hi = tc::make_desc(0, a_lbo, 128), b_desc_hi = tc::make_desc(0, b_lbo, 128);
      const uint32_t a_kstep = (2 * a_lbo) >> 4, b_kstep = (2 * b_lbo) >> 4;  // descriptor address units (16 B)
      uint32_t e_base = 0, t = 0, acc = 0;
      int b, z0, zlen, y0, ylen, x0, xlen;
      for (ItemIter it(p); it.next(p, b, z0, zlen, y0, ylen, x0, xlen);) {
        for (int i = 0; i < xlen; ++i, ++acc) {
          const uint32_t q = acc & 1, uq = acc >> 1;
          if (uq > 0) tc::mbar_wait(&tm_empty[q], (uq - 1) & 1);
          tc::tc_fence_after();
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t e = e_base + i + dx, slot = e % kPlaneSlots;
            tc::mbar_wait(&plane_full[slot], (e / kPlaneSlots) & 1);
            tc::tc_fence_after();
            const uint32_t a_plane = (planes_u32 + slot * p.plane_bytes) >> 4;
            for (int dy = 0; dy < 3; ++dy)
              for (int dz = 0; dz < 3; ++dz, ++t) {
                const int tap = (dx * 3 + dy) * 3 + dz;
                const uint32_t s = t % p.b_stages;
                tc::mbar_wait(&b_full[s], (t / p.b_stages) & 1);
                tc::tc_fence_after();
                const uint64_t b_desc0 = b_desc_hi | (uint64_t)(((bt_u32 + s * p.btile_bytes) >> 4) & 0x3FFF);
                const uint32_t a_tap = a_plane + (uint32_t)(dy * p.Zh + dz);
                const uint64_t a_desc0 = a_desc_hi | (uint64_t)(a_tap & 0x3FFF);
                const uint32_t d_tmem0 = tmem_base + q * (MT * p.N);
                if (leader) {
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk)
                      tc::umma_bf16(d_tmem0 + mt * p.N, a_desc0 + (uint64_t)(mt * 128 + kk * a_kstep),
                                    b_desc0 + (uint64_t)(kk * b_kstep), idesc, (uint32_t)((tap | kk) != 0));
                  }
                  tc::umma_commit(&b_empty[s]);
                }
                __syncwarp();
              }
            if (dx == 0 && leader) tc::umma_commit(&plane_empty[slot]);  // last use of input plane x-1
          }
          if (leader) tc::umma_commit(&tm_full[q]);
        }
        if (leader) {
          tc::umma_commit(&plane_empty[(e_base + xlen) % kPlaneSlots]);
          tc::umma_commit(&plane_empty[(e_base + xlen + 1) % kPlaneSlots]);
        }
        e_base += xlen + 2;
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 0..3 <-> TMEM lanes 32*warp..)
    uint32_t acc = 0;
    int b, z0, zlen, y0, ylen, x0, xlen;
    for (ItemIter it(p); it.next(p, b, z0, zlen, y0, ylen, x0, xlen);) {
      for (int i = 0; i < xlen; ++i, ++acc) {
        const uint32_t q = acc & 1;
        tc::mbar_wait(&tm_full[q], (acc >> 1) & 1);
        tc::tc_fence_after();
        for (int mt = 0; mt < p.mtiles; ++mt) {
          const int r = mt * 128 + warp * 32 + lane;
          const int oy = r / p.Zh, oz = r - oy * p.Zh;
          const bool valid = oy < ylen && oz < zlen;
          bf16 *dst = out + ((((size_t)b * p.X + (x0 + i)) * p.Y + (y0 + oy)) * p.Z + (z0 + oz)) * p.N;
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (q * p.mtiles + mt) * p.N;
          for (int c0 = 0; c0 < p.N; c0 += 16) {
            uint32_t v[16];
            tc::tmem_ld16(taddr + c0, v);
            tc::tmem_ld_wait();
            if (valid) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                pk[j] = *reinterpret_cast<uint32_t *>(&h);
              }
              uint4 *d4 = reinterpret_cast<uint4 *>(dst + c0);
              d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tm_empty[q]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

// [tap][Cb][Cs] (generic packed) -> [tap'][Cin/8][N][8]; flip = dgrad (reverse taps, swap channel roles)
__global__ void repack_b_kernel(const bf16 *__restrict__ wp, bf16 *__restrict__ wb, int Cb, int Cs, int taps, int flip) {
  const int Cin = flip ? Cs : Cb, N = flip ? Cb : Cs;
  const int64_t total = (int64_t)taps * Cin * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i & 7);
    int64_t t = i >> 3;
    const int n = (int)(t % N); t /= N;
    const int cc = (int)(t % (Cin >> 3));
    const int tap = (int)(t / (Cin >> 3));
    const int ci = cc * 8 + c8;
    const int src_tap = flip ? (taps - 1 - tap) : tap;
    const int cb = flip ? n : ci, cs = flip ? ci : n;
    wb[i] = wp[((int64_t)src_tap * Cb + cb) * Cs + cs];
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

void *tc_encode_fn_ptr() { return reinterpret_cast<void *>(encode_fn()); }

// conv_tc_prog.cu
bool tc_prog_supported(const cgan3d_conv_geom &g, int dtype, int op);
size_t tc_prog_workspace_bytes(const cgan3d_conv_geom &g, int dtype, int op);
int tc_prog_run(const cgan3d_conv_geom &g, int scatter, const void *in, const void *wp, void *outp, void *ws, size_t ws_bytes,
                cudaStream_t st);

static bool plan_s1(int B, int X, int Y, int Z, int Cin, int N, TcPlan &best) {
  if (N % 16 || N > 256 || N < 16) return false;
  if (Cin != 16 && Cin != 32 && Cin != 64 && Cin != 128) return false;
  TcPlan p{};
  p.B = B; p.X = X; p.Y = Y; p.Z = Z; p.Cin = Cin; p.N = N;
  p.nzt = (Z + 61) / 62;
  p.Zt = (Z + p.nzt - 1) / p.nzt;
  p.Zh = p.Zt + 2;
  p.btile_bytes = (uint32_t)Cin * N * 2;
  p.b_stages = p.btile_bytes <= 8192 ? 4 : (p.btile_bytes <= 16384 ? 3 : 2);
  double best_eff = 0;
  bool found = false;
  for (int Yt = 1; Yt <= Y && Yt + 2 <= 256; ++Yt) {
    const int mt = (Yt * p.Zh + 127) / 128;
    if (2 * mt * N > 512 || mt > 4) break;
    const int rows_alloc = ((mt * 128 + 2 * p.Zh + 2) + 7) / 8 * 8;
    const uint32_t plane_bytes = (uint32_t)(Cin / 8) * rows_alloc * 16;
    const uint32_t smem = kPlaneSlots * plane_bytes + p.b_stages * p.btile_bytes + 512;
    if (rows_alloc * 16 > 16383 * 16) break;
    if (smem > kSmemLimit) break;
    const int nslabs = (Y + Yt - 1) / Yt;
    const double eff = (double)Y * p.Zt / ((double)nslabs * mt * 128);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      found = true;
      best = p;
      best.Yt = Yt; best.Yh = Yt + 2; best.nslabs = nslabs; best.mtiles = mt; best.rows_alloc = rows_alloc;
      best.plane_bytes = plane_bytes; best.smem_bytes = smem;
    }
  }
  if (!found) return false;
  TcPlan &q = best;
  q.box_bytes = 16u * q.Zh * q.Yh;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * q.mtiles * N)) cols <<= 1;
  q.tmem_cols = cols;
  q.xseg = X;
  q.nxseg = 1;
  q.nitems = B * q.nzt * q.nslabs;  // columns; the kernel splits columns x planes evenly over the grid
  return true;
}


// ================================================================================================================
// Kernel 2: wgrad_s1_tc_kernel — weight gradient of the 3x3x3 stride-1 convolution (Cs == 64 output channels,
// Cb % 16 == 0 input channels, Z <= 62).  dW[co][ci][tap] = sum_v dY[v][co] * X[v + tap][ci].
//   * GEMM view per tap: D[co, ci] (M = 64, N = Cb) with K = voxels.  Both operands keep the channels-last
//     activation layout [C/8][row][8 ch] in shared memory, i.e. they are MN-major UMMA operands (a core matrix is
//     8 voxels x 8 channels); the tap is again a pure row shift of the X operand.
//   * A CTA owns one dx (filter x-offset) and a contiguous range of (b, y-slab, x) steps; per step it TMA-loads the
//     dY slab (rows = oy*Zh + oz, halo columns zero-filled by TMA OOB) and the X halo slab of plane x+dx-1, and issues
//     9 taps x K/16 MMAs into 9 TMEM accumulators (M = 64 accumulators are interleaved two per 64 columns).
//   * Split-K over CTAs: the partial 9 x 64 x Cb sums are added to the fp32 gradient with red.global.add.f32.
struct WgPlan {
  int B, X, Y, Z, Cb, Cs;
  int Zh, Yt, Yh, nslabs;
  int kpad, rowsA, rowsB;
  int steps_per_dx;
  uint32_t a_bytes, b_bytes, boxA_bytes, boxB_bytes, stage_bytes, smem_bytes, tmem_cols;
  int stages;
};

__global__ void __launch_bounds__(192, 1)
wgrad_s1_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, float *__restrict__ dw,
                   const WgPlan p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *stage_mem = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(stage_mem + (size_t)p.stages * p.stage_bytes);
  uint64_t *full = bars, *empty = bars + p.stages, *done = bars + 2 * p.stages;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the staging buffers once: rows that TMA never writes (K padding, tail of the shifted reads) must read as 0
  {
    uint4 *z = reinterpret_cast<uint4 *>(stage_mem);
    const uint32_t n16 = (uint32_t)p.stages * p.stage_bytes / 16;
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) {
    tc::tmem_alloc(tmem_ptr, p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work split: blockIdx.x % 3 selects dx; the CTAs of one dx share its steps evenly
  const int dx = blockIdx.x % 3;
  const int grp = blockIdx.x / 3, ngrp = (gridDim.x - dx + 2) / 3;
  const int s_begin = (int)((long long)p.steps_per_dx * grp / ngrp), s_end = (int)((long long)p.steps_per_dx * (grp + 1) / ngrp);
  const int kch_a = p.Cs >> 3, kch_b = p.Cb >> 3;

  if (warp == 4) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmY);
      tc::tma_prefetch_desc(&tmX);
      for (int st = s_begin, n = 0; st < s_end; ++st, ++n) {
        int t = st;
        const int x = t % p.X; t /= p.X;
        const int sl = t % p.nslabs;
        const int b = t / p.nslabs;
        const int y0 = sl * p.Yt;
        const uint32_t s = n % p.stages, use = n / p.stages;
        if (use > 0) tc::mbar_wait(&empty[s], (use - 1) & 1);
        tc::mbar_expect_tx(&full[s], p.boxA_bytes * kch_a + p.boxB_bytes * kch_b);
        uint8_t *a = stage_mem + (size_t)s * p.stage_bytes, *bb = a + p.a_bytes;
        for (int cc = 0; cc < kch_a; ++cc) tc::tma_load_5d(a + (size_t)cc * p.rowsA * 16, &tmY, &full[s], cc * 8, 0, y0, x, b);
        for (int cc = 0; cc < kch_b; ++cc)
          tc::tma_load_5d(bb + (size_t)cc * p.rowsB * 16, &tmX, &full[s], cc * 8, -1, y0 - 1, x + dx - 1, b);
      }
    }
  } else if (warp == 5) {
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_bf16(64, p.Cb, 1, 1);
    const uint32_t a_sbo = (uint32_t)p.rowsA * 16, b_sbo = (uint32_t)p.rowsB * 16;
    const uint64_t a_hi = tc::make_desc(0, 128, a_sbo), b_hi = tc::make_desc(0, 128, b_sbo);
    const uint32_t stage0 = tc::smem_u32(stage_mem);
    const int kblocks = p.kpad >> 4;
    for (int st = s_begin, n = 0; st < s_end; ++st, ++n) {
      const uint32_t s = n % p.stages;
      tc::mbar_wait(&full[s], (n / p.stages) & 1);
      tc::tc_fence_after();
      const uint32_t a0 = (stage0 + s * p.stage_bytes) >> 4, b0 = (stage0 + s * p.stage_bytes + p.a_bytes) >> 4;
      if (leader) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const uint64_t a_desc = a_hi | (uint64_t)((a0 + kb * 16) & 0x3FFF);
          const uint32_t acc_on = (n | kb) != 0;
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            const int dy = j / 3, dz = j % 3;
            const uint64_t b_desc = b_hi | (uint64_t)((b0 + kb * 16 + dy * p.Zh + dz) & 0x3FFF);
            const uint32_t d = tmem_base + (uint32_t)((j >> 1) * p.Cb) + ((uint32_t)((j & 1) * 16) << 16);
            tc::umma_bf16(d, a_desc, b_desc, idesc, acc_on);
          }
        }
        tc::umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (leader) tc::umma_commit(done);
    __syncwarp();
  } else {
    // epilogue: warps 0..3; lanes 0-15 hold accumulator 2g (row 16*warp + lane), lanes 16-31 accumulator 2g+1
    if (s_end > s_begin) {
      tc::mbar_wait(done, 0);
      tc::tc_fence_after();
      const int co = warp * 16 + (lane & 15);
      for (int g2 = 0; g2 < 5; ++g2) {
        const int j = g2 * 2 + (lane >> 4);
        const int tap = dx * 9 + j;
        for (int c0 = 0; c0 < p.Cb; c0 += 16) {
          uint32_t v[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g2 * p.Cb + c0), v);
          tc::tmem_ld_wait();
          if (j < 9) {
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(&dw[((size_t)co * p.Cb + (c0 + c)) * 27 + tap], __uint_as_float(v[c]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

static bool plan_wgrad(const cgan3d_conv_geom &g, WgPlan &p) {
  if (g.k != 3 || g.stride != 1 || g.pad != 1) return false;
  if (g.Cs != 64 || g.Cb % 16 || g.Cb < 16 || g.Cb > 64) return false;  // 5 column groups x Cb <= 512 TMEM columns
  if (g.Zb > 62) return false;
  p = WgPlan{};
  p.B = g.B; p.X = g.Xb; p.Y = g.Yb; p.Z = g.Zb; p.Cb = g.Cb; p.Cs = g.Cs;
  p.Zh = p.Z + 2;
  p.stages = 2;
  bool ok = false;
  for (int Yt = mn(p.Y, 64); Yt >= 1; --Yt) {
    const int kpad = (Yt * p.Zh + 15) / 16 * 16;
    const int rowsA = kpad, rowsB = (kpad + 2 * p.Zh + 2 + 7) / 8 * 8;
    const uint32_t a_bytes = (uint32_t)(g.Cs / 8) * rowsA * 16, b_bytes = (uint32_t)(g.Cb / 8) * rowsB * 16;
    if (rowsB * 16 > 16383 * 16) continue;
    if ((size_t)p.stages * (a_bytes + b_bytes) + 512 > kSmemLimit) continue;
    // prefer slabs of equal height
    p.Yt = Yt; p.Yh = Yt + 2; p.kpad = kpad; p.rowsA = rowsA; p.rowsB = rowsB; p.a_bytes = a_bytes; p.b_bytes = b_bytes;
    ok = true;
    break;
  }
  if (!ok) return false;
  p.nslabs = (p.Y + p.Yt - 1) / p.Yt;
  p.Yt = (p.Y + p.nslabs - 1) / p.nslabs;  // balance
  p.Yh = p.Yt + 2;
  p.kpad = (p.Yt * p.Zh + 15) / 16 * 16;
  p.rowsA = p.kpad;
  p.rowsB = (p.kpad + 2 * p.Zh + 2 + 7) / 8 * 8;
  p.a_bytes = (uint32_t)(g.Cs / 8) * p.rowsA * 16;
  p.b_bytes = (uint32_t)(g.Cb / 8) * p.rowsB * 16;
  p.stage_bytes = p.a_bytes + p.b_bytes;
  p.boxA_bytes = 16u * p.Zh * p.Yt;
  p.boxB_bytes = 16u * p.Zh * p.Yh;
  p.smem_bytes = p.stages * p.stage_bytes + 512;
  p.steps_per_dx = p.B * p.nslabs * p.X;
  uint32_t cols = 32;
  while (cols < (uint32_t)(5 * g.Cb)) cols <<= 1;
  p.tmem_cols = cols;
  return true;
}

static int encode_act_map(CUtensorMap *tm, const void *ptr, int C, int Z, int Y, int X, int B, int boxZ, int boxY) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(CGAN3D_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
  const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)B};
  const cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)Z * C * 2, (cuuint64_t)Y * Z * C * 2, (cuuint64_t)X * Y * Z * C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)boxZ, (cuuint32_t)boxY, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CGAN3D_E_SHAPE, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

static bool s1_shape_ok(const cgan3d_conv_geom &g, int dtype, int op) {
  if (dtype != CGAN3D_BF16 || op < 0 || op > 2) return false;
  if (g.k != 3 || g.stride != 1 || g.pad != 1) return false;
  if (g.Xb != g.Xs || g.Yb != g.Ys || g.Zb != g.Zs) return false;
  return true;
}

bool tc_supported(const cgan3d_conv_geom &g, int dtype, int op) {
  if (!cgan3d_device_supports_tc() || encode_fn() == nullptr) return false;
  if (!s1_shape_ok(g, dtype, op)) return tc_prog_supported(g, dtype, op);
  if (op == 2) {
    WgPlan w;
    return plan_wgrad(g, w);
  }
  TcPlan p;
  const int Cin = op == 0 ? g.Cb : g.Cs, N = op == 0 ? g.Cs : g.Cb;
  return plan_s1(g.B, g.Xb, g.Yb, g.Zb, Cin, N, p);
}

size_t tc_workspace_bytes(const cgan3d_conv_geom &g, int dtype, int op) {
  if (!s1_shape_ok(g, dtype, op)) return tc_prog_workspace_bytes(g, dtype, op);
  if (op == 2) return 0;
  return (size_t)27 * g.Cb * g.Cs * 2 + 256;
}

static int run_s1(const cgan3d_conv_geom &g, int flip, const void *in, const void *wp, void *outp, void *ws, size_t ws_bytes,
                  cudaStream_t st) {
  const int Cin = flip ? g.Cs : g.Cb, N = flip ? g.Cb : g.Cs;
  TcPlan p;
  if (!plan_s1(g.B, g.Xb, g.Yb, g.Zb, Cin, N, p)) return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 conv: no tiling for this shape");
  const size_t need = (size_t)27 * Cin * N * 2;
  if (ws == nullptr || ws_bytes < need) return fail(CGAN3D_E_WORKSPACE, "tcgen05 conv: workspace %zu < %zu", ws_bytes, need);
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(outp) & 15) || (reinterpret_cast<uintptr_t>(ws) & 15))
    return fail(CGAN3D_E_ARG, "tcgen05 conv: pointers must be 16-byte aligned");
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(CGAN3D_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
  bf16 *wb = reinterpret_cast<bf16 *>(ws);
  repack_b_kernel<<<64, 256, 0, st>>>(reinterpret_cast<const bf16 *>(wp), wb, g.Cb, g.Cs, 27, flip);
  CG_LAUNCH_CHECK("repack_b");
  CUtensorMap tm;
  const cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)p.Z, (cuuint64_t)p.Y, (cuuint64_t)p.X, (cuuint64_t)p.B};
  const cuuint64_t gstr[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)p.Z * Cin * 2, (cuuint64_t)p.Y * p.Z * Cin * 2,
                              (cuuint64_t)p.X * p.Y * p.Z * Cin * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)p.Zh, (cuuint32_t)p.Yh, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CGAN3D_E_SHAPE, "cuTensorMapEncodeTiled failed with %d", (int)r);
  const int grid = (int)mn<long long>((long long)p.nitems * p.X, (long long)num_sms());
  auto launch = [&](auto ks_tag, auto mt_tag) -> int {
    constexpr int KS = decltype(ks_tag)::value;
    constexpr int MT = decltype(mt_tag)::value;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(conv_s1_tc_kernel<KS, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kSmemLimit + 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_s1_tc_kernel)");
      attr_set = true;
    }
    conv_s1_tc_kernel<KS, MT><<<grid, kThreads, p.smem_bytes + 1024, st>>>(tm, wb, reinterpret_cast<bf16 *>(outp), p);
    CG_LAUNCH_CHECK("conv_s1_tc_kernel");
    return 0;
  };
  auto by_mt = [&](auto ks_tag) -> int {
    switch (p.mtiles) {
      case 1: return launch(ks_tag, std::integral_constant<int, 1>{});
      case 2: return launch(ks_tag, std::integral_constant<int, 2>{});
      case 3: return launch(ks_tag, std::integral_constant<int, 3>{});
      case 4: return launch(ks_tag, std::integral_constant<int, 4>{});
      default: return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 conv: mtiles %d not built", p.mtiles);
    }
  };
  switch (Cin >> 4) {
    case 1: return by_mt(std::integral_constant<int, 1>{});
    case 2: return by_mt(std::integral_constant<int, 2>{});
    case 4: return by_mt(std::integral_constant<int, 4>{});
    case 8: return by_mt(std::integral_constant<int, 8>{});
    default: return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 conv: Cin must be 16, 32, 64 or 128");
  }
}

int tc_gather(const cgan3d_conv_geom &g, const void *big, const void *wp, const float *bias, void *small, void *ws,
              size_t ws_bytes, cudaStream_t st) {
  if (bias) return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 conv: bias is applied by the bias_act pass");
  if (!s1_shape_ok(g, CGAN3D_BF16, 0)) return tc_prog_run(g, 0, big, wp, small, ws, ws_bytes, st);
  return run_s1(g, 0, big, wp, small, ws, ws_bytes, st);
}

int tc_scatter(const cgan3d_conv_geom &g, const void *small, const void *wp, const float *bias, void *big, void *ws,
               size_t ws_bytes, cudaStream_t st) {
  if (bias) return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 conv: bias is applied by the bias_act pass");
  if (!s1_shape_ok(g, CGAN3D_BF16, 1)) return tc_prog_run(g, 1, small, wp, big, ws, ws_bytes, st);
  return run_s1(g, 1, small, wp, big, ws, ws_bytes, st);
}

int tc_wgrad(const cgan3d_conv_geom &g, const void *big, const void *small, float *dw, float beta, void *, size_t,
             cudaStream_t st) {
  WgPlan p;
  if (!plan_wgrad(g, p)) return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 wgrad: shape not supported");
  if ((reinterpret_cast<uintptr_t>(big) & 15) || (reinterpret_cast<uintptr_t>(small) & 15))
    return fail(CGAN3D_E_ARG, "tcgen05 wgrad: pointers must be 16-byte aligned");
  if (beta == 0.f) {
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)g.Cs * g.Cb * 27 * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "tcgen05 wgrad memset");
  }
  CUtensorMap tmY, tmX;
  int r = encode_act_map(&tmY, small, g.Cs, p.Z, p.Y, p.X, p.B, p.Zh, p.Yt);
  if (r) return r;
  r = encode_act_map(&tmX, big, g.Cb, p.Z, p.Y, p.X, p.B, p.Zh, p.Yh);
  if (r) return r;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_s1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit + 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad_s1_tc_kernel)");
    attr_set = true;
  }
  const int grid = (int)mn<long long>((long long)3 * p.steps_per_dx, (long long)(num_sms() / 3) * 3);
  wgrad_s1_tc_kernel<<<grid, 192, p.smem_bytes + 1024, st>>>(tmY, tmX, dw, p);
  CG_LAUNCH_CHECK("wgrad_s1_tc_kernel");
  return 0;
}

}  // namespace cg
