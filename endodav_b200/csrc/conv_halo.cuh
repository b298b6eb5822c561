// 3x3 convolution (pad 1, stride 1) for the 64-channel DPT refinement maps, tcgen05 implicit GEMM
// with HALO REUSE: ResidualConvUnit convs, layer1_rn, output_conv1 (util/blocks.py:78-91,
// dpt.py:100-117).
//
// gemm_tc_kernel<CONV=true> fetches the A tile of every tap with its own TMA load, i.e. each input
// pixel crosses L2 -> SM nine times; at 64 channels that traffic (not the tensor pipe, not HBM) is
// what bounds those convs.  Here the 18 x 10 pixel halo of a 16 x 8 output tile is brought into
// shared memory ONCE (cp.async, 16 B per pixel per 8-channel chunk, zero-filled outside the image)
// in the chunk-planar layout [chunk][pixel] x 16 B.  That layout is the canonical no-swizzle
// K-major UMMA operand for any tap shift (8 pixels of a tile row = one core matrix, SBO = halo row
// pitch, LBO = plane size), so the nine taps are nine descriptors into the same tile.  The
// weights [N][9*C] stay resident in shared memory for the whole (persistent) CTA.
//
//   warps 0..3 / 4..7  two epilogue warpgroups, one per TMEM accumulator (gt_epilogue: fused
//                      bias / ReLU / residual(s) / relu-copy, coalesced 128-bit stores)
//   warp  8            TMEM allocator + MMA issuer: 9 x C/16 tcgen05.mma (M=128, N=BN, K=16)
//   warps 9..12        producers: cp.async halo loads, 4-deep ring, two tiles in flight
#pragma once
#include "gemm_tc.cuh"

namespace tc {

constexpr int CH_TH = 16, CH_TW = 8;
constexpr int CH_HW = CH_TW + 2, CH_HH = CH_TH + 2;
constexpr int CH_PIX = CH_HH * CH_HW;          // 180
constexpr int CH_PLANE = CH_PIX * 16;
constexpr int CH_THREADS = 13 * 32;
constexpr int CH_PRODUCERS = 4 * 32;
constexpr int CH_STAGES = 4;
constexpr int CH_LOOK = 2;                     // tiles whose cp.async groups are in flight

template <int CIN, int BN> constexpr size_t ch_smem_bytes() {
  return 1024 + (size_t)9 * (CIN / 8) * BN * 16 + CH_STAGES * (size_t)(CIN / 8) * CH_PLANE + 8 * (size_t)GT_STG_WORDS * 4 + 256 + 64 * 4 + 16;
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int CIN, int BN>
__global__ void __launch_bounds__(CH_THREADS, 1)
    conv3x3_halo_kernel(const T* __restrict__ x, const T* __restrict__ w, Epi e, int F, int H, int W, int tiles_x,
                        int tiles_y, int total_tiles) {
  pdl_launch();   // PDL: the next kernel of the stream may start its prologue (common.cuh)
  constexpr int NCH = CIN / 8;
  constexpr uint32_t W_BYTES = 9 * NCH * BN * 16;
  constexpr uint32_t HALO_BYTES = NCH * CH_PLANE;
  constexpr uint32_t TMEM_COLS = 2 * BN < 64 ? 64 : 2 * BN;
  constexpr int NT = (CH_PIX * NCH + CH_PRODUCERS - 1) / CH_PRODUCERS;   // tasks per producer thread per tile
  extern __shared__ __align__(1024) unsigned char ch_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ch_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* wsm = smem;
  unsigned char* halo = smem + W_BYTES;
  uint32_t* stg_base = reinterpret_cast<uint32_t*>(halo + CH_STAGES * HALO_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + 8 * GT_STG_WORDS);
  uint64_t* halo_full = bars;
  uint64_t* halo_empty = bars + CH_STAGES;
  uint64_t* acc_full = bars + 2 * CH_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);   // BN floats, resident (see gt_applyN)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < BN) bias_s[threadIdx.x] = e.bias ? e.bias[threadIdx.x] : 0.f;

  // weights [BN][9*CIN] (ky,kx,c) -> [tap][chunk][n] x 16 B (canonical no-swizzle K-major B operand)
  for (int i = threadIdx.x; i < 9 * NCH * BN; i += CH_THREADS) {
    const int n = i % BN;
    const int tc_ = i / BN;  // tap * NCH + chunk
    *reinterpret_cast<uint4*>(wsm + ((size_t)tc_ * BN + n) * 16) =
        *reinterpret_cast<const uint4*>(w + (size_t)n * (9 * CIN) + (size_t)tc_ * 8);
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < CH_STAGES; ++b) {
      mbar_init(&halo_full[b], CH_PRODUCERS);
      mbar_init(&halo_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 8) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // PDL: everything above ran under the previous kernel's tail; its results are visible from here
  const int per_frame = tiles_x * tiles_y;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 9) {
    // ===== producers =====
    const int pt = threadIdx.x - 9 * 32;
    // this thread's (halo pixel, chunk) tasks are tile-independent
    int dsto[NT], hyx[NT];
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const int t = pt + u * CH_PRODUCERS;
      const int p = t / NCH, kc = t - p * NCH;
      const int hy = p / CH_HW, hx = p - hy * CH_HW;
      dsto[u] = (t < CH_PIX * NCH) ? kc * CH_PLANE + p * 16 : -1;
      hyx[u] = (hy << 16) | (hx << 8) | kc;
    }
    const uint32_t halo_u32 = smem_u32(halo);
    auto issue = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const uint32_t b = it % CH_STAGES;
      const int f = tile / per_frame;
      const int r = tile - f * per_frame;
      const int ty = r / tiles_x;
      const int y0t = ty * CH_TH - 1, x0t = (r - ty * tiles_x) * CH_TW - 1;
      mbar_wait(&halo_empty[b], ((it / CH_STAGES) & 1) ^ 1);
      const T* xf = x + (size_t)f * H * W * CIN;
      const uint32_t hb = halo_u32 + b * HALO_BYTES;
#pragma unroll
      for (int u = 0; u < NT; ++u) {
        if (dsto[u] < 0) continue;
        const int iy = y0t + (hyx[u] >> 16), ix = x0t + ((hyx[u] >> 8) & 255), kc = hyx[u] & 255;
        const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const T* src = in ? xf + ((size_t)iy * W + ix) * CIN + kc * 8 : xf;
        cp_async16_zfill(hb + dsto[u], src, in ? 16u : 0u);
      }
      cp_async_commit();
    };
    // software pipeline: the loads of tile j are issued CH_LOOK iterations before tile j is published
#pragma unroll 1
    for (int j = 0; j < my_tiles + CH_LOOK; ++j) {
      if (j < my_tiles) issue(j);
      else cp_async_commit();                   // empty group keeps the wait_group arithmetic uniform
      const int it = j - CH_LOOK;
      if (it < 0) continue;
      cp_async_wait<CH_LOOK>();
      fence_proxy_async();                      // cp.async (generic proxy) -> tcgen05 (async proxy)
      mbar_arrive(&halo_full[it % CH_STAGES]);
    }
  } else if (warp == 8) {
    // ===== MMA issuer (warp-uniform loop, elected lane issues: tc_common.cuh elect_one_sync) =====
    constexpr uint32_t idesc = make_idesc<T>(128, BN, 0);
    const uint32_t wa = smem_u32(wsm);
    const uint32_t leader = elect_one_sync();
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t b = it & 1, ph = (it >> 1) & 1;
      const uint32_t hs = it % CH_STAGES;
      mbar_wait(&acc_empty[b], ph ^ 1);
      mbar_wait(&halo_full[hs], (it / CH_STAGES) & 1);
      fence_after_sync();
      const uint32_t ha = smem_u32(halo + hs * HALO_BYTES);
      // descriptors differ only in the 14-bit start-address field: build once, add offsets (>>4)
      const uint64_t adesc0 = make_smem_desc(ha, CH_HW * 16, CH_PLANE, 0);
      const uint64_t bdesc0 = make_smem_desc(wa, 128, BN * 16, 0);
      if (leader) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint64_t at = adesc0 + (uint64_t)(ky * CH_HW + kx);                 // 16 B per pixel
          const uint64_t bt = bdesc0 + (uint64_t)(tap * NCH * BN);                  // BN x 16 B per chunk
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k)
            mma_ss(tmem_base + b * BN, at + (uint64_t)(2 * k * (CH_PLANE / 16)), bt + (uint64_t)(2 * k * BN), idesc, (tap | k) ? 1u : 0u);
        }
        mma_commit(&halo_empty[hs]);
        mma_commit(&acc_full[b]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warpgroups =====
    const uint32_t g = warp >> 2;           // accumulator buffer
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int dy = r >> 3, dx = r & 7;
    uint32_t* stg = stg_base + warp * GT_STG_WORDS;
    for (int it = (int)g; it < my_tiles; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int f = tile / per_frame;
      const int rr = tile - f * per_frame;
      const int ty = rr / tiles_x;
      const int y = ty * CH_TH + dy, xx = (rr - ty * tiles_x) * CH_TW + dx;
      const bool valid = (y < H) && (xx < W);
      const long long m = ((long long)f * H + y) * W + xx;
      mbar_wait(&acc_full[g], (it >> 1) & 1);
      fence_after_sync();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + g * BN;
#pragma unroll 1
      for (int half = 0; half < (BN >= 64 ? 2 : 1); ++half)
        gt_epilogue<T, BN>(e, trow, valid, m, m, 0, 0, stg, lane, half, e.bias ? bias_s + half * (BN >= 64 ? BN / 2 : BN) : nullptr);
      fence_before_sync();
      mbar_arrive(&acc_empty[g]);
    }
  }
  __syncthreads();
  if (warp == 8) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
