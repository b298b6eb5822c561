// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   C[M,N] = A[M,K] * W[N,K]^T   (16-bit operands, fp32 accumulation in TMEM)
//
// One CTA computes a 128 x BN tile.  Warp roles (192 threads):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor loads of the A and W k-blocks into a
//                 ring of 128B(64B)-swizzled shared-memory stages, completion on mbarriers
//   warp 1      : TMEM allocator + MMA issuer -- one lane issues tcgen05.mma (M=128, N=BN,
//                 K=16) per 16-element k-step, tcgen05.commit releases the stage / signals
//                 the epilogue
//   warps 2..5  : epilogue -- tcgen05.ld the accumulator (lane = row), fused
//                 bias / row-bias / GELU / GEGLU / residual / ReLU / pixel-shuffle /
//                 disparity-head epilogue (common.cuh), 16-byte vector stores
//
// CONV=false: A is a row-major [M,K] matrix (2-D tensor map), M tail rows are zero-filled by
//             TMA and masked in the epilogue.
// CONV=true : A is an NHWC activation [F,H,W,C]; the M tile is a TH x TW pixel patch of one
//             frame and k-block (tap, c0) is the patch shifted by the tap offset, fetched
//             with a 4-D tensor map whose out-of-bounds fill supplies the zero padding.
//             K = 9*C in (ky,kx,c) order.
#pragma once
#include "tc_common.cuh"

namespace tc {

struct ConvTile {
  int H, W, C;          // NHWC input == output spatial size (stride 1, pad 1)
  int th, tw;           // tile shape, th*tw == 128
  int tiles_y, tiles_x; // per frame
};

constexpr int GT_BM = 128;
constexpr int GT_THREADS = 192;

template <int BN, int BK> constexpr int gt_stage_bytes() { return (GT_BM + BN) * BK * 2; }

template <typename T, int BN, int BK, bool CONV>
__global__ void __launch_bounds__(GT_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, Epi e, int M,
                                                             int N, int K, int stages, ConvTile ct) {
  constexpr uint32_t ROW_BYTES = BK * 2;                    // 128 (BK=64) or 64 (BK=32)
  constexpr uint32_t SWZ = (BK == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * ROW_BYTES;                   // 8-row swizzle atom
  constexpr uint32_t A_BYTES = GT_BM * ROW_BYTES;
  constexpr uint32_t B_BYTES = BN * ROW_BYTES;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN must be a power of two in [32,256]");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [stages x (A|B)] then barriers
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* accum_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = N / BN;
  const int tile_n = blockIdx.x % n_tiles;
  const int tile_m = blockIdx.x / n_tiles;
  const int n0 = tile_n * BN;
  const int kblocks = K / BK;

  // conv tile coordinates
  int cf = 0, y0 = 0, x0 = 0;
  if (CONV) {
    int per = ct.tiles_y * ct.tiles_x;
    cf = tile_m / per;
    int r = tile_m - cf * per;
    int tyi = r / ct.tiles_x;
    y0 = tyi * ct.th;
    x0 = (r - tyi * ct.tiles_x) * ct.tw;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int cblocks = CONV ? ct.C / BK : 1;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % stages;
        const uint32_t it = kb / stages;
        mbar_wait(&empty_bar[s], (it & 1) ^ 1);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        if (CONV) {
          const int tap = kb / cblocks;
          const int c0 = (kb - tap * cblocks) * BK;
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d(sa, &tmA, &full_bar[s], c0, x0 + kx - 1, y0 + ky - 1, cf);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, tile_m * GT_BM);
        }
        tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<T>(GT_BM, BN, 0);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % stages;
        const uint32_t it = kb / stages;
        mbar_wait(&full_bar[s], it & 1);
        fence_after_sync();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        const uint64_t adesc = make_smem_desc(sa, SBO, 16, SWZ);
        const uint64_t bdesc = make_smem_desc(sb, SBO, 16, SWZ);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the >>4 address field
          mma_ss(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        mma_commit(&empty_bar[s]);  // stage reusable once these MMAs have read it
      }
      mma_commit(accum_bar);        // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the tile
    mbar_wait(accum_bar, 0);
    fence_after_sync();
    long long m, orow;
    bool valid;
    if (CONV) {
      const int dy = r / ct.tw, dx = r - dy * ct.tw;
      const int y = y0 + dy, x = x0 + dx;
      valid = (y < ct.H) && (x < ct.W);
      m = ((long long)cf * ct.H + y) * ct.W + x;
      orow = m;
    } else {
      m = (long long)tile_m * GT_BM + r;
      valid = m < M;
      orow = epi_row(e, m);
    }
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    if (e.act == ACT_GEGLU) {
      // BN == 128: columns [0,64) value, [64,128) gate (pack.py pairs them per tile)
      if constexpr (BN == 128) {
#pragma unroll 1
        for (int c = 0; c < 64; c += 16) {
          float hv[16], gv[16];
          tmem_ld16(trow + c, hv);
          tmem_ld16(trow + 64 + c, gv);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float hh = hv[i] + __ldg(e.bias + n0 + c + i);
              float gg = gv[i] + __ldg(e.bias + n0 + 64 + c + i);
              hv[i] = hh * gelu_erf(gg);
            }
            store_vec<T, 16>((T*)e.out + orow * e.ldo + tile_n * 64 + c, hv);
          }
        }
      }
    } else if (e.act == ACT_HEAD) {
      // whole row in one tile (BN == N == 32): relu(conv+b) . w + b -> relu  (dpt.py:118-123)
      if constexpr (BN == 32) {
        float v[32];
        tmem_ld32(trow, v);
        if (valid) {
          float s = e.head_b + __ldg(e.head_w + 32);  // head_w[32] carries the 1x1 conv bias
#pragma unroll
          for (int i = 0; i < 32; ++i) s = fmaf(fmaxf(v[i] + __ldg(e.bias + i), 0.f), __ldg(e.head_w + i), s);
          if (e.sig_sign == 0.f) s = fmaxf(s, 0.f);             // output_conv2: trailing ReLU
          else s = 1.f / (1.f + expf(-e.sig_sign * s));         // HeadDepth + sigmoid
          ((float*)e.out)[orow] = s;
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        if (valid) epi_apply<T, 16>(e, m, orow, n0 + c, v);
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
