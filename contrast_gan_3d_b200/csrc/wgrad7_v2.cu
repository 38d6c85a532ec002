// Weight gradient of the generator's two thin 7x7x7 layers, second generation (replaces wgrad7_thin_tc_kernel +
// expand_z_kernel of conv_thin_tc.cu on the shapes it supports).
//   `first`     (reference model/generator.py:31-38, aten::convolution_backward weight path): S16 = dY [B,Xs,Ys,Zs,16],
//               Q1 = reflection-padded input [B,Xs+6,Ys+6,Zs+6], P = 0
//   `last_conv` (generator.py:77-83): S16 = padded input, Q1 = dY [B,Xs-6,...], P = 6, taps come out flipped
//   R[c][dx,dy,dz] = sum_v S16[v, c] * Q1[v + (dx,dy,dz) - P]
//
// What the first kernel lost (profiles/r01_ncu_full_wgrad7_last.txt: 0.198 of the bf16 peak, 3.4 GB of DRAM traffic for a
// 1.3 GB problem) and what changes here:
//   * the z-expanded operand E[v][j] = Q1[v + j] was materialised in HBM by a pre-pass (16 B per voxel written, then read
//     2.75x through the y halo).  Here it is built IN SHARED MEMORY by the four epilogue warps from the raw 1-channel
//     lines (2 B per voxel, L2 resident): DRAM traffic = one read of S16.
//   * M was 64 (8 dy x 8 dz) with N = 64 (4 x-planes x 16 channels) and 2.5 MMAs per plane on average.  Here one E row
//     carries TWO x-planes (32-byte rows = [dx_lo][dz], SWIZZLE_32B MN-major, blocks of 16 M elements one slab LINE apart
//     = 8 dy blocks): M = 128.  N = 128 = the 8 S16 planes xe-6 .. xe+1 x 16 channels sitting in adjacent ring slots
//     (a ring of 10 planes whose first slots are mirrored behind its end, so a window never wraps): ONE M = 128, N = 128
//     MMA per 16 voxels and per PAIR of E planes, 14 of its 16 (dx_lo, plane) blocks and 49 of 64 rows useful.
//   * accumulator column = 16 * (plane - first plane of the window): the filter x-offset is dx = 6 - j + dx_lo.
// Split-K over CTAs (whole (b, y-tile) columns dealt round-robin, see SegIterW2), fp32 atomics at the end.
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace cg {

using bf16 = __nv_bfloat16;
constexpr uint32_t kSmemLimitW2 = 232448 - 1024;
constexpr int kMaxRingW2 = 14;  // logical S16 plane slots: 8 live + 4 or 6 in flight (TMA runs 2 - 3 steps ahead of the MMAs)
constexpr int kMaxPrefW2 = 11;     // staged 32-bit words per fetch thread and step
constexpr int kBuildWarpsW2 = 10;  // warps 0-3 (also the epilogue) and 6-11 expand the operand
constexpr int kFetchWarp0W2 = 12;  // warps 12-15 fetch and stage the raw Q1 words
constexpr int kFetchersW2 = 128;
constexpr int kThreadsW2 = 512;
constexpr int kBuildersW2 = (kBuildWarpsW2 + kFetchersW2 / 32) * 32;  // threads on the named barrier shared by the build and fetch warps (see bar_sync_builders)

struct ThinW2Plan {
  int B, Xs, Ys, Zs;  // S16 extents
  int Xq, Yq, Zq;     // Q1 extents
  int Xe;             // E planes = Xs + 6
  int P, flip;
  int Zt, Yt, nyt, L; // z rows per line (multiple of 16, >= Zs), S16 lines per step, y tiles, E lines per step (Yt + 7)
  int LW;             // staged words per Q1 line segment
  int rows, kblocks;  // rows per S16 slot (Yt * Zt), K blocks per step
  int ring, nphys, npairs;  // logical ring size; physical slots (>= ring: the first nphys - ring slots are mirrored)
  int nb16, inv_nb16;       // 16-row blocks per line and ceil(65536 / nb16)
  uint32_t slot_bytes, e2_bytes, box_bytes, stage_words, smem_bytes;
  int debug;  // CGAN3D_W2_DEBUG (profiling aid, results are wrong): 1 = skip the operand build, 2 = skip the MMAs, 4 = skip the plane loads
};

// Work = ncols columns (b, y-tile) x npairs plane pairs.  Columns are dealt to the CTAs ROUND-ROBIN (CTA c takes columns
// c, c + grid, ... whole), so that at any moment the 148 CTAs stream the same planes of 148 ADJACENT y-tiles, i.e. one
// contiguous region of the 16-channel tensor: DRAM sees long sequential bursts instead of 148 scattered 8 KB pieces.  The
// columns of the last, incomplete round are split evenly by plane pairs (a range that crosses a column boundary becomes
// two segments), so every CTA gets the same number of steps to within one.
struct SegIterW2 {
  int n, rounds, r;
  long long rem_col0, idx, end;
  __device__ __forceinline__ SegIterW2(long long ncols, int n_) : n(n_), r(0) {
    rounds = (int)(ncols / gridDim.x);
    rem_col0 = (long long)rounds * gridDim.x;
    const long long total = (ncols - rem_col0) * n;
    idx = total * blockIdx.x / gridDim.x;
    end = total * (blockIdx.x + 1) / gridDim.x;
  }
  __device__ __forceinline__ long long steps() const { return (long long)rounds * n + (end - idx); }
  // step k (0-based, in processing order) of this CTA -> (column, plane pair)
  __device__ __forceinline__ void locate(long long k, long long idx0, int &col, int &pair) const {
    if (k < (long long)rounds * n) {
      const int rr = (int)(k / n);
      col = rr * (int)gridDim.x + (int)blockIdx.x;
      pair = (int)(k - (long long)rr * n);
    } else {
      const long long flat = idx0 + (k - (long long)rounds * n);
      const long long c = flat / n;
      col = (int)(rem_col0 + c);
      pair = (int)(flat - c * n);
    }
  }
  __device__ __forceinline__ bool next(int &col, int &p0, int &plen) {
    if (r < rounds) {
      col = r * (int)gridDim.x + (int)blockIdx.x;
      p0 = 0;
      plen = n;
      ++r;
      return true;
    }
    if (idx >= end) return false;
    const long long c = idx / n;
    col = (int)(rem_col0 + c);
    p0 = (int)(idx - c * n);
    plen = (int)mn<long long>(n - p0, end - idx);
    idx += plen;
    return true;
  }
};

__device__ __forceinline__ void bar_sync_builders() { static_assert(kBuildersW2 == 448, "barrier count"); asm volatile("bar.sync 1, 448;" ::: "memory"); }

__global__ void __launch_bounds__(kThreadsW2, 1)
wgrad7_v2_kernel(const __grid_constant__ CUtensorMap tmS, const uint32_t *__restrict__ q1, float *__restrict__ dw,
                 const __grid_constant__ ThinW2Plan p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *ring = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // nphys S16 plane slots, [rows][16 ch], SWIZZLE_32B (TMA)
  uint8_t *e2 = ring + (size_t)p.nphys * p.slot_bytes;          // 2 stages of the expanded operand, [L * Zt rows][2][8], SWIZZLE_32B
  uint32_t *stage = reinterpret_cast<uint32_t *>(e2 + 2 * (size_t)p.e2_bytes);  // raw Q1 line segments [2][L][LW]
  uint64_t *bars = reinterpret_cast<uint64_t *>(stage + p.stage_words);
  uint64_t *s_full = bars, *s_empty = bars + kMaxRingW2, *e_full = s_empty + kMaxRingW2, *e_empty = e_full + 2, *done = e_empty + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxRingW2; ++i) { tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&e_full[i], kBuildWarpsW2); tc::mbar_init(&e_empty[i], 1); }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) {
    tc::tmem_alloc(tmem_ptr, 128);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp < 4) {  // every MMA accumulates: start from zero
    for (int c0 = 0; c0 < 128; c0 += 16) tc::tmem_zero16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  const long long ncols = (long long)p.B * p.nyt;

  if (warp == 4) {
    // ------------------------------------------------ S16 plane producer (TMA).  Measured (profiles/README.md, round 2): a
    // SWIZZLE_32B map is limited to 32-byte inner rows and the TMA engine spends ~3 cycles per row, so the planes of one
    // step (2 planes + mirrors = 660 rows) cost ~2000 cycles: this, not the MMAs (1350 cycles), is the kernel's floor.  A
    // 16-byte cp.async producer warp with the swizzle applied by hand was tried and is slower still (5400 cycles / step).  Planes are numbered by the order in which
    // this CTA loads them (q): slot q % ring, mirrored at ring + slot when that exists.  A segment loads its 6 warm-up
    // planes and then two planes per step, (ring - 8) / 2 steps ahead of the MMAs.
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmS);
      uint32_t q = 0;
      auto load_plane = [&](int xs, int b, int y0) {
        const uint32_t slot = q % (uint32_t)p.ring, use = q / (uint32_t)p.ring;
        if (use > 0) tc::mbar_wait(&s_empty[slot], (use - 1) & 1);
        if (p.debug & 4) { tc::mbar_arrive(&s_full[slot]); ++q; return; }  // profiling aid: no plane loads at all
        const bool mirror = (int)slot + p.ring < p.nphys;
        tc::mbar_expect_tx(&s_full[slot], mirror ? 2 * p.box_bytes : p.box_bytes);
        tc::tma_load_5d(ring + (size_t)slot * p.slot_bytes, &tmS, &s_full[slot], 0, 0, y0, xs, b);
        if (mirror) tc::tma_load_5d(ring + (size_t)(slot + p.ring) * p.slot_bytes, &tmS, &s_full[slot], 0, 0, y0, xs, b);
        ++q;
      };
      int col, p0, plen;
      for (SegIterW2 it(ncols, p.npairs); it.next(col, p0, plen);) {
        const int b = col / p.nyt, y0 = (col - b * p.nyt) * p.Yt;
        for (int xs = 2 * p0 - 6; xs < 2 * p0; ++xs) load_plane(xs, b, y0);
        for (int i = 0; i < plen; ++i) {
          load_plane(2 * (p0 + i), b, y0);
          load_plane(2 * (p0 + i) + 1, b, y0);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer
    const bool leader = tc::elect_one();
    // A: MN-major SWIZZLE_32B, 16 M elements per 32-byte row, further M blocks one slab line later (LBO), 8-row K groups 256 B
    const uint64_t a_hi = tc::make_desc_sw_mn(0, (uint32_t)p.Zt * 32u, 256, 32);
    // B: MN-major SWIZZLE_32B, one 16-channel N block per ring slot (LBO = slot stride)
    const uint64_t b_hi = tc::make_desc_sw_mn(0, p.slot_bytes, 256, 32);
    const uint32_t ring_u32 = tc::smem_u32(ring), e2_u32 = tc::smem_u32(e2);
    uint32_t waited = 0, n = 0, w0 = 0;  // planes waited for / steps issued / sequence number of the window's first plane
    int col, p0, plen;
    for (SegIterW2 it(ncols, p.npairs); it.next(col, p0, plen);) {
      w0 = waited;  // the segment's first window starts at its first warm-up plane
      for (int i = 0; i < plen; ++i, ++n, w0 += 2) {
        while (waited < w0 + 8) {
          tc::mbar_wait(&s_full[waited % (uint32_t)p.ring], (waited / (uint32_t)p.ring) & 1);
          ++waited;
        }
        const uint32_t st = n & 1;
        tc::mbar_wait(&e_full[st], (n >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t a0 = (e2_u32 + st * p.e2_bytes) >> 4;
        uint32_t j = 0;
        while (j < 8) {  // runs of window planes that are contiguous in the (mirrored) ring
          const uint32_t slot = (w0 + j) % (uint32_t)p.ring;
          const uint32_t run = mn<uint32_t>(8 - j, (uint32_t)p.nphys - slot);
          const uint32_t idesc = tc::make_idesc_bf16(128, (int)(16 * run), 1, 1);
          const uint32_t b0 = (ring_u32 + slot * p.slot_bytes) >> 4;
          const uint32_t d = tmem_base + 16 * j;
          if (leader && !(p.debug & 2)) {
            uint64_t a_desc = a_hi | (uint64_t)(a0 & 0x3FFF), b_desc = b_hi | (uint64_t)(b0 & 0x3FFF);
#pragma unroll 4
            for (int kb = 0; kb < p.kblocks; ++kb) {
              tc::umma_bf16(d, a_desc, b_desc, idesc, 1u);
              a_desc += 32;  // 16 rows of 32 B
              b_desc += 32;
            }
          }
          __syncwarp();
          j += run;
        }
        if (leader) {
          tc::umma_commit(&e_empty[st]);
          tc::umma_commit(&s_empty[w0 % (uint32_t)p.ring]);        // the two oldest planes leave the window
          tc::umma_commit(&s_empty[(w0 + 1) % (uint32_t)p.ring]);
        }
        __syncwarp();
      }
      // the six planes still resident belong to this segment only
      if (leader)
        for (uint32_t k = 0; k < 6; ++k) tc::umma_commit(&s_empty[(w0 + k) % (uint32_t)p.ring]);
      __syncwarp();
    }
    if (leader) tc::umma_commit(done);
    __syncwarp();
  } else {
    // ------------------------------------------------ warps 0..3, 6..11: build the expanded operand (0..3 then run the
    // epilogue); warps 12..15: fetch the raw Q1 words two steps ahead and stage them in shared memory.  The roles are split
    // because the generic->async proxy fence that publishes the operand compiles to MEMBAR.ALL.CTA, which waits for every
    // outstanding global load of the executing thread: a warp that prefetches AND builds exposes the full load latency
    // (~1 us under load) in every step.
    const int ZqW = p.Zq >> 1;
    // this CTA's steps are the flat (column, plane pair) indices [idx0, idx0 + total)
    const SegIterW2 range(ncols, p.npairs);
    const long long idx0 = range.idx;
    const uint32_t total = (uint32_t)range.steps();
    uint32_t n = 0;
    if (warp >= kFetchWarp0W2) {
      const int tid = (warp - kFetchWarp0W2) * 32 + lane;  // 0..127
      const int total_words = 2 * p.L * p.LW;
      // word i of this thread: staging index tid + 128 i = (h, l, w); rel = its offset from the step's base pointer; the
      // z-range check does not depend on the step (meta < 0: never loaded)
      int meta[kMaxPrefW2], rel[kMaxPrefW2];
#pragma unroll
      for (int i = 0; i < kMaxPrefW2; ++i) {
        const int idx = tid + kFetchersW2 * i;
        meta[i] = -1; rel[i] = 0;
        if (idx < total_words) {
          const int w = idx % p.LW, hl = idx / p.LW, l = hl % p.L, h = hl / p.L;
          const int wq = w - p.P / 2;  // staged word 0 holds Q1 z elements (-P, -P + 1); P is even
          if ((unsigned)wq < (unsigned)ZqW) { meta[i] = l | (h << 8); rel[i] = (h * p.Yq + l) * ZqW + w; }
        }
      }
      auto fetch = [&](uint32_t k, uint32_t (&pref)[kMaxPrefW2]) {  // raw Q1 words of this CTA's step k -> registers
        int col, pair;
        range.locate(k, idx0, col, pair);
        const int b = col / p.nyt, y0 = (col - b * p.nyt) * p.Yt;
        const int xb = 2 * pair - p.P, yb = y0 - p.P;
        // base may point outside the tensor (halo): it is only dereferenced for in-range (plane, line)
        const uint32_t *base = q1 + ((long long)b * p.Xq * p.Yq + (long long)xb * p.Yq + yb) * ZqW - p.P / 2;
        const bool hv0 = (unsigned)xb < (unsigned)p.Xq, hv1 = (unsigned)(xb + 1) < (unsigned)p.Xq;
#pragma unroll
        for (int i = 0; i < kMaxPrefW2; ++i) {
          uint32_t v = 0;
          if (meta[i] >= 0) {
            const bool hv = (meta[i] >> 8) ? hv1 : hv0;
            if (hv && (unsigned)(yb + (meta[i] & 0xFF)) < (unsigned)p.Yq) v = __ldg(base + rel[i]);
          }
          pref[i] = v;
        }
      };
      auto put = [&](uint32_t k, uint32_t (&pref)[kMaxPrefW2]) {
        bar_sync_builders();  // the builders have finished reading the previous step's staging area
#pragma unroll
        for (int i = 0; i < kMaxPrefW2; ++i)
          if (tid + kFetchersW2 * i < total_words) stage[tid + kFetchersW2 * i] = pref[i];
        bar_sync_builders();
        if (k + 2 < total) fetch(k + 2, pref);
      };
      uint32_t prefA[kMaxPrefW2], prefB[kMaxPrefW2];
      if (total > 0) fetch(0, prefA);
      if (total > 1) fetch(1, prefB);
      while (n < total) {
        put(n, prefA);
        if (++n >= total) break;
        put(n, prefB);
        ++n;
      }
    } else {
      const int bw = warp < 4 ? warp : warp - 2;  // builder warp 0..9
      // E2[l * Zt + r][h][j] = line(h, l)[r + j], j = 0..7.  A warp iteration = 32 rows of one staged line (h, l): 5
      // conflict-free LDS (neighbouring lanes read overlapping words), one funnel shift per output word (odd rows start
      // in the upper half-word), one 16-byte STS (SWIZZLE_32B makes the 8 rows of a quarter-warp hit 8 different 16-byte
      // bank groups).  The 2 L ipl iterations of a step are dealt to the 6 builder warps as contiguous ranges.
      const int ipl = (p.Zt + 31) >> 5, iters = 2 * p.L * ipl, per = (iters + kBuildWarpsW2 - 1) / kBuildWarpsW2;
      const int it0 = bw * per, it1 = mn(iters, it0 + per);
      const int hl0 = it0 / ipl, rb0 = it0 - hl0 * ipl;
      const uint32_t stage_u32 = tc::smem_u32(stage) + (uint32_t)(lane >> 1) * 4u;
      const uint32_t lane_dst = (uint32_t)lane * 32u;
      const uint32_t swz = (uint32_t)((lane >> 2) & 1);  // rows advance by 32 per iteration: the swizzle phase is the lane's
      const uint32_t sh = (uint32_t)(lane & 1) << 4;
      for (; n < total; ++n) {
        bar_sync_builders();
        bar_sync_builders();  // the fetch warps have written this step's staging area
        const uint32_t st = n & 1;
        if (n >= 2) tc::mbar_wait(&e_empty[st], ((n >> 1) - 1) & 1);
        if (!(p.debug & 1)) {
          const uint32_t dst_u32 = tc::smem_u32(e2) + st * p.e2_bytes + lane_dst;
          int it = it0, hl = hl0, rb = rb0;
          while (it < it1) {
            const int h = hl >= p.L ? 1 : 0, l = hl - h * p.L;
            uint32_t src = stage_u32 + (uint32_t)(hl * p.LW + rb * 16) * 4u;
            uint32_t d = dst_u32 + (uint32_t)(l * p.Zt + rb * 32) * 32u + (((uint32_t)h ^ swz) << 4);
            int r = rb * 32 + lane;
#pragma unroll 2
            for (; rb < ipl && it < it1; ++rb, ++it, src += 64, d += 1024, r += 32) {
              if (r < p.Zt) {
                uint32_t a0, a1, a2, a3, a4;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a0) : "r"(src));
                asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(a1) : "r"(src));
                asm volatile("ld.shared.b32 %0, [%1+8];" : "=r"(a2) : "r"(src));
                asm volatile("ld.shared.b32 %0, [%1+12];" : "=r"(a3) : "r"(src));
                asm volatile("ld.shared.b32 %0, [%1+16];" : "=r"(a4) : "r"(src));
                const uint32_t o0 = __funnelshift_r(a0, a1, sh), o1 = __funnelshift_r(a1, a2, sh), o2 = __funnelshift_r(a2, a3, sh),
                               o3 = __funnelshift_r(a3, a4, sh);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
              }
            }
            rb = 0;
            ++hl;
          }
        }
        tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&e_full[st]);
      }
    }
    // epilogue: TMEM lane m = (dy = m >> 4, dx_lo = (m >> 3) & 1, dz = m & 7); column = 16 * j + c, dx = 6 - j + dx_lo
    if (n > 0 && warp < 4) {
      tc::mbar_wait(done, 0);
      tc::tc_fence_after();
      const int m = warp * 32 + lane, dy = m >> 4, dxl = (m >> 3) & 1, dz = m & 7;
      for (int j = 0; j < 8; ++j) {
        uint32_t v[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(16 * j), v);
        tc::tmem_ld_wait();
        const int dx = 6 - j + dxl;
        if (dy < 7 && dz < 7 && dx >= 0 && dx < 7) {
          int tap = dx * 49 + dy * 7 + dz;
          if (p.flip) tap = 342 - tap;
#pragma unroll
          for (int c = 0; c < 16; ++c) atomicAdd(&dw[(size_t)c * 343 + tap], __uint_as_float(v[c]));
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFnW2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void *tc_encode_fn_ptr();              // conv_tc.cu
CUtensorMapL2promotion tc_l2_promo();  // conv_tc.cu

static bool plan_w2(const cgan3d_conv_geom &g, ThinW2Plan &p) {
  static int off = -1;
  if (off < 0) off = getenv("CGAN3D_WGRAD7_V1") ? 1 : 0;
  if (off) return false;
  if (g.k != 7 || g.stride != 1) return false;
  const bool first = g.Cb == 1 && g.Cs == 16, last = g.Cb == 16 && g.Cs == 1;
  if (!first && !last) return false;
  p = ThinW2Plan{};
  p.B = g.B;
  if (first) { p.Xs = g.Xs; p.Ys = g.Ys; p.Zs = g.Zs; p.Xq = g.Xb; p.Yq = g.Yb; p.Zq = g.Zb; p.P = g.pad; p.flip = 0; }
  else       { p.Xs = g.Xb; p.Ys = g.Yb; p.Zs = g.Zb; p.Xq = g.Xs; p.Yq = g.Ys; p.Zq = g.Zs; p.P = 6 - g.pad; p.flip = 1; }
  if ((p.P & 1) || (p.Zq & 1)) return false;  // staged words hold aligned element pairs
  p.Xe = p.Xs + 6;
  p.npairs = (p.Xe + 1) / 2;
  p.Zt = (p.Zs + 15) / 16 * 16;
  if (p.Zt > 256) return false;
  p.LW = (p.Zt + 8) / 2 + 1;
  p.nb16 = p.Zt / 16;
  p.inv_nb16 = (65536 + p.nb16 - 1) / p.nb16;
  for (int Yt = mn(p.Ys, 8); Yt >= 1; --Yt) {
    const int L = Yt + 7;
    if (2 * L * p.LW > kFetchersW2 * kMaxPrefW2) continue;
    const uint32_t slot = (uint32_t)Yt * p.Zt * 32u;
    const uint32_t e2b = ((uint32_t)L * p.Zt * 32u + 1023u) / 1024u * 1024u;
    const uint32_t stage_words = ((uint32_t)(2 * L * p.LW) + 3u) & ~3u;
    const uint32_t fixed = 2 * e2b + stage_words * 4 + 512;
    if (slot % 1024) continue;
    int nphys = (int)((kSmemLimitW2 - mn(kSmemLimitW2, fixed)) / slot);
    if (nphys > kMaxRingW2 + 6) nphys = kMaxRingW2 + 6;
    // ring of 14 (TMA three steps ahead) when at least four of its slots can be mirrored, else 12; a mirrored slot keeps the
    // window of 8 planes contiguous, so most steps need ONE N = 128 MMA per K block
    const int ring = nphys >= kMaxRingW2 + 4 ? kMaxRingW2 : 12;
    if (nphys < ring + 2 && Yt > 1) continue;
    if (nphys < ring) continue;
    if (nphys > ring + 6) nphys = ring + 6;  // windows start at even slots <= ring - 2: six mirrors make every window contiguous
    if (const char *e = getenv("CGAN3D_W2_NPHYS")) nphys = mx(ring, mn(nphys, atoi(e)));  // experiment: fewer mirrored slots
    p.Yt = Yt; p.L = L; p.nphys = nphys; p.ring = ring;
    p.slot_bytes = slot; p.e2_bytes = e2b; p.stage_words = stage_words;
    p.smem_bytes = fixed + (uint32_t)nphys * slot;
    break;
  }
  if (p.Yt == 0) return false;
  p.nyt = (p.Ys + p.Yt - 1) / p.Yt;
  p.rows = p.Yt * p.Zt;
  p.kblocks = p.rows / 16;
  p.box_bytes = p.slot_bytes;
  return true;
}

bool thin_w2_supported(const cgan3d_conv_geom &g) {
  ThinW2Plan p;
  return plan_w2(g, p);
}

int thin_w2_run(const cgan3d_conv_geom &g, const void *big, const void *small, float *dw, float beta, cudaStream_t st) {
  ThinW2Plan p;
  if (!plan_w2(g, p)) return fail(CGAN3D_E_UNSUPPORTED, "tcgen05 thin wgrad v2: shape not supported");
  if (const char *e = getenv("CGAN3D_W2_DEBUG")) p.debug = atoi(e);
  const bool first = p.flip == 0;
  const void *s16 = first ? small : big, *q1 = first ? big : small;
  if ((reinterpret_cast<uintptr_t>(s16) & 15) || (reinterpret_cast<uintptr_t>(q1) & 3))
    return fail(CGAN3D_E_ARG, "tcgen05 thin wgrad v2: the 16-channel tensor must be 16-byte aligned, the 1-channel tensor 4-byte aligned");
  if (beta == 0.f) {
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)16 * 343 * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "tcgen05 thin wgrad v2 memset");
  }
  EncodeTiledFnW2 enc = reinterpret_cast<EncodeTiledFnW2>(tc_encode_fn_ptr());
  if (!enc) return fail(CGAN3D_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
  CUtensorMap tmS;
  {
    const cuuint64_t gdim[5] = {16, (cuuint64_t)p.Zs, (cuuint64_t)p.Ys, (cuuint64_t)p.Xs, (cuuint64_t)p.B};
    const cuuint64_t gstr[4] = {32, (cuuint64_t)p.Zs * 32, (cuuint64_t)p.Ys * p.Zs * 32, (cuuint64_t)p.Xs * p.Ys * p.Zs * 32};
    const cuuint32_t box[5] = {16, (cuuint32_t)p.Zt, (cuuint32_t)p.Yt, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(s16), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, tc_l2_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CGAN3D_E_SHAPE, "cuTensorMapEncodeTiled (thin wgrad v2) failed with %d", (int)r);
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad7_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimitW2 + 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(wgrad7_v2_kernel)");
    attr_set = true;
  }
  const long long total = (long long)p.B * p.nyt * p.npairs;
  const int grid = (int)mn<long long>(total, (long long)num_sms());
  wgrad7_v2_kernel<<<grid, kThreadsW2, p.smem_bytes + 1024, st>>>(tmS, reinterpret_cast<const uint32_t *>(q1), dw, p);
  CG_LAUNCH_CHECK("wgrad7_v2_kernel");
  return 0;
}

}  // namespace cg
