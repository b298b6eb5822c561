// Persistent tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   C[M,N] = A[M,K] * W[N,K]^T   (16-bit operands, fp32 accumulation in TMEM)
//
// One CTA per SM loops over 128 x BN output tiles (static round-robin, n fastest so that CTAs
// running together share A tiles through L2).  640 threads:
//   warp 0       TMA producer  -- cp.async.bulk.tensor loads of the A and W k-blocks into a ring
//                of 128B(64B)-swizzled shared-memory stages, completion on mbarriers; the ring
//                runs ahead across tile boundaries
//   warp 1       TMEM allocator + MMA issuer -- one lane issues tcgen05.mma (M=128, N=BN, K=16)
//                into one of TWO accumulator buffers in TMEM; tcgen05.commit releases the smem
//                stage / hands the finished accumulator to its epilogue warpgroup
//   warps 2,3    idle (keep the epilogue warpgroups 4-warp aligned for the TMEM lane quarters)
//   warps 4..19  four epilogue warpgroups: warpgroup wg drains columns [half*BN/2, +BN/2) of
//                accumulator buffer (wg & 1) (even / odd tiles of this CTA), half = wg >> 1
//                tcgen05.ld (lane = row), fused bias / row-bias / GELU / GEGLU / residual / ReLU /
//                pixel-shuffle / disparity-head epilogue (common.cuh), 16-byte vector stores.
// So the epilogue of tile i overlaps the mainloop of tile i+1 -- with K = 384 (ViT-S) the
// epilogue is as long as the mainloop and hiding it is what matters.
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
constexpr int GT_THREADS = 640;

template <int BN, int BK> constexpr int gt_stage_bytes() { return (GT_BM + BN) * BK * 2; }
template <int BN> constexpr uint32_t gt_tmem_cols() { return 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512; }

constexpr int GT_STG_WORDS = 32 * 32;   // per-warp transposition buffer: 32 rows x 32 fp32, 16-byte chunks XOR-swizzled

// Row-owner -> row-segment transposition of one warp's 32 x 32 fp32 block through shared memory.
// write: thread = row, 8 x STS.128 with the 16-byte chunk index XORed by (row & 7)  (conflict-free)
// read : lane = (row it*4 + lane/8, columns 4*(lane%8)..+3), LDS.128                (conflict-free)
__device__ __forceinline__ void gt_stage_write(uint32_t* stg, int lane, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ float4 gt_stage_read(const uint32_t* stg, int rr, int j) {
  return *reinterpret_cast<const float4*>(stg + rr * 32 + ((j ^ (rr & 7)) << 2));
}

__device__ __forceinline__ void f4_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
template <typename T> __device__ __forceinline__ float4 ld4_as_f32(const void* p, int is_f32, long long off) {
  if (is_f32) return *reinterpret_cast<const float4*>((const float*)p + off);   // plain loads: may alias the output (x += ...)
  float v[4];
  load_vec<T, 4>((const T*)p + off, v);
  return make_float4(v[0], v[1], v[2], v[3]);
}

// Compile-time epilogue kinds.  The runtime-flag version of this code is a chain of
// `if (flag) { unrolled block }`: every skipped block is a taken branch to a far target, and with
// the epilogue far larger than the L0 instruction cache each of them is an instruction-fetch stall
// (ncu: stall_no_inst on every other line, ~2 500 cycles per 32-column chunk).  The hot call sites
// of the forward therefore get straight-line instantiations (F >= 0: bit mask below); anything
// else (row-bias tables, pixel shuffle, sigmoid) runs the generic F = -1 version.
// (the EF_* bit mask is declared in common.cuh: the launcher side needs it without the kernels)

// The fused epilogue on NS row segments (4 consecutive columns starting at n of rows mm[it] ->
// output rows oo[it], oo < 0 = masked).  Same order of operations as epi_apply (common.cuh):
// bias, per-frame row bias, GELU, residual(s), ReLU / sigmoid, stores.  Every stage is a loop over
// the NS segments so the loads of a stage are all in flight together.
template <typename T, int NS, int F>
__device__ __forceinline__ void gt_applyN(const Epi& e, float4* a, const int* mm, const int* oo, int n,
                                          const float* bias4) {
  constexpr bool S = F >= 0;
  const bool has_bias = S ? ((F & EF_BIAS) != 0) : (bias4 != nullptr);
  const bool do_gelu = S ? ((F & EF_GELU) != 0) : (e.act == ACT_GELU);
  const bool do_relu = S ? ((F & EF_RELU) != 0) : (e.act == ACT_RELU);
  const bool has_res1 = S ? ((F & (EF_RES1_F32 | EF_RES1_T)) != 0) : (e.res1 != nullptr);
  const int res1_f32 = S ? ((F & EF_RES1_F32) != 0) : e.res1_f32;
  const bool has_res2 = S ? ((F & EF_RES2_T) != 0) : (e.res2 != nullptr);
  const int res2_f32 = S ? 0 : e.res2_f32;
  const bool has_out = S ? true : (e.out != nullptr);
  const int out_f32 = S ? ((F & EF_OUT_F32) != 0) : e.out_f32;
  const bool has_out_relu = S ? ((F & EF_OUT_RELU) != 0) : (e.out_relu != nullptr);
  if (has_bias) {
    // the bias slice was staged in shared memory before the accumulator wait (with ~200 KB of the
    // SM's memory carved out as shared, L1 holds almost nothing)
    const float4 b = *reinterpret_cast<const float4*>(bias4);
#pragma unroll
    for (int it = 0; it < NS; ++it) f4_add(a[it], b);
  }
  const bool has_rowbias = S ? ((F & EF_ROWBIAS) != 0) : (e.rowbias != nullptr);
  if (has_rowbias) {
#pragma unroll
    for (int it = 0; it < NS; ++it)
      if (oo[it] >= 0) f4_add(a[it], *reinterpret_cast<const float4*>(e.rowbias + (long long)((mm[it] / e.rb_div) % e.rb_mod) * e.rb_ld + n));
  }
  if (do_gelu) {
#pragma unroll
    for (int it = 0; it < NS; ++it) {
      a[it].x = gelu_act<T>(a[it].x); a[it].y = gelu_act<T>(a[it].y);
      a[it].z = gelu_act<T>(a[it].z); a[it].w = gelu_act<T>(a[it].w);
    }
  }
  long long orow[NS];
  int ocol = n;
  if (!S && e.map == MAP_PIXSHUF) {
    // m = (f, y, x) over the ps_h x ps_w grid; n = (ky*k + kx)*ps_c + c
    const int tap = n / e.ps_c;
    ocol = n - tap * e.ps_c;
    const int ky = tap / e.ps_k, kx = tap - ky * e.ps_k;
    const int hw = e.ps_h * e.ps_w;
#pragma unroll
    for (int it = 0; it < NS; ++it) {
      const int f = mm[it] / hw;
      const int r = mm[it] - f * hw;
      const int y = r / e.ps_w, x = r - y * e.ps_w;
      orow[it] = ((long long)f * (e.ps_h * e.ps_k) + (long long)y * e.ps_k + ky) * ((long long)e.ps_w * e.ps_k) + (long long)x * e.ps_k + kx;
    }
  } else {
#pragma unroll
    for (int it = 0; it < NS; ++it) orow[it] = oo[it];
  }
  if (has_res1) {
    float4 r[NS];
#pragma unroll
    for (int it = 0; it < NS; ++it) r[it] = (oo[it] >= 0) ? ld4_as_f32<T>(e.res1, res1_f32, orow[it] * e.ld_res1 + ocol) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < NS; ++it) f4_add(a[it], r[it]);
  }
  if (has_res2) {
    float4 r[NS];
#pragma unroll
    for (int it = 0; it < NS; ++it) r[it] = (oo[it] >= 0) ? ld4_as_f32<T>(e.res2, res2_f32, orow[it] * e.ld_res2 + ocol) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < NS; ++it) f4_add(a[it], r[it]);
  }
  if (do_relu) {
#pragma unroll
    for (int it = 0; it < NS; ++it) { a[it].x = fmaxf(a[it].x, 0.f); a[it].y = fmaxf(a[it].y, 0.f); a[it].z = fmaxf(a[it].z, 0.f); a[it].w = fmaxf(a[it].w, 0.f); }
  } else if (!S && e.act == ACT_SIGMOID) {
#pragma unroll
    for (int it = 0; it < NS; ++it) {
      a[it].x = 1.f / (1.f + __expf(-e.sig_sign * a[it].x)); a[it].y = 1.f / (1.f + __expf(-e.sig_sign * a[it].y));
      a[it].z = 1.f / (1.f + __expf(-e.sig_sign * a[it].z)); a[it].w = 1.f / (1.f + __expf(-e.sig_sign * a[it].w));
    }
  }
  if (has_out) {
    if (out_f32) {
#pragma unroll
      for (int it = 0; it < NS; ++it)
        if (oo[it] >= 0) *reinterpret_cast<float4*>((float*)e.out + orow[it] * e.ldo + ocol) = a[it];
    } else {
#pragma unroll
      for (int it = 0; it < NS; ++it)
        if (oo[it] >= 0) {
          uint2 w;
          w.x = pack2(from_f<T>(a[it].x), from_f<T>(a[it].y));
          w.y = pack2(from_f<T>(a[it].z), from_f<T>(a[it].w));
          *reinterpret_cast<uint2*>((T*)e.out + orow[it] * e.ldo + ocol) = w;
        }
    }
  }
  if (has_out_relu) {
#pragma unroll
    for (int it = 0; it < NS; ++it)
      if (oo[it] >= 0) {
        uint2 w;
        w.x = pack2(from_f<T>(fmaxf(a[it].x, 0.f)), from_f<T>(fmaxf(a[it].y, 0.f)));
        w.y = pack2(from_f<T>(fmaxf(a[it].z, 0.f)), from_f<T>(fmaxf(a[it].w, 0.f)));
        *reinterpret_cast<uint2*>((T*)e.out_relu + orow[it] * e.ldo + ocol) = w;
      }
  }
}

// The row-segment path of gt_epilogue for columns [half*BN/2, +BN/2) of one accumulator.
template <typename T, int BN, int F>
__device__ __forceinline__ void gt_epi_rows(const Epi& e, uint32_t trow, int m32, int o32, int n0, uint32_t* stg, int lane,
                                            int half, const float* bias_s) {
  constexpr int CH0 = (BN >= 64) ? BN / 2 : BN;
  const int j = lane & 7, rsub = lane >> 3;
#pragma unroll 1
  for (int c = half * CH0; c < (BN >= 64 ? (half + 1) * CH0 : BN); c += 32) {
    float v[32];
    tmem_ld32(trow + c, v);
    gt_stage_write(stg, lane, v);
    __syncwarp();
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {   // 2 x 4 row segments: keeps the epilogue under 96 registers
      float4 a[4];
      int mm[4], oo[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int rr = (hb * 4 + it) * 4 + rsub;
        mm[it] = __shfl_sync(0xffffffffu, m32, rr);
        oo[it] = __shfl_sync(0xffffffffu, o32, rr);
        a[it] = gt_stage_read(stg, rr, j);
      }
      gt_applyN<T, 4, F>(e, a, mm, oo, n0 + c + 4 * j, bias_s ? bias_s + (c - half * CH0) + 4 * j : nullptr);
    }
    __syncwarp();
  }
}

// epilogue of one 128 x BN accumulator: thread = row `r` of the tile, TMEM address `trow`.
// The accumulator leaves TMEM in the row-owner domain (tcgen05.ld: lane = row) and is transposed
// per warp through `stg`, so that the fused epilogue math (common.cuh: epi_apply) and every
// global access -- bias / residual loads, stores -- run with 8 lanes covering 32 consecutive
// columns of one row (128-bit accesses, 4 full rows per instruction).
template <typename T, int BN>
__device__ __forceinline__ void gt_epilogue(const Epi& e, uint32_t trow, bool valid, long long m, long long orow_lin,
                                            int n0, int tile_n, uint32_t* stg, int lane, int half,
                                            const float* bias_s) {
  const int m32 = (int)m;                          // rows < 2^31 (checked by the launcher)
  const int o32 = valid ? (int)orow_lin : -1;
  const int j = lane & 7, rsub = lane >> 3;
  // the two warpgroups of an accumulator split its columns: [0, BN/2) and [BN/2, BN)
  constexpr int CH0 = (BN >= 64) ? BN / 2 : BN;
  if (e.act == ACT_HEAD) {
    // whole row in one tile (BN == N == 32): relu(conv+b) . w + b -> relu  (dpt.py:118-123)
    if (half != 0) return;
    if constexpr (BN == 32) {
      float v[32];
      tmem_ld32(trow, v);
      if (valid) {
        float s = e.head_b + __ldg(e.head_w + 32);  // head_w[32] carries the 1x1 conv bias
#pragma unroll
        for (int i = 0; i < 32; ++i) s = fmaf(fmaxf(v[i] + __ldg(e.bias + i), 0.f), __ldg(e.head_w + i), s);
        if (e.sig_sign == 0.f) s = fmaxf(s, 0.f);             // output_conv2: trailing ReLU
        else s = 1.f / (1.f + expf(-e.sig_sign * s));         // HeadDepth + sigmoid
        ((float*)e.out)[orow_lin] = s;
      }
    }
  } else if (e.act == ACT_GEGLU) {
    // BN == 128: columns [0,64) value, [64,128) gate (pack.py pairs them per tile)
    if constexpr (BN == 128) {
      {
        const int c = half * 32;
        float v[32];
        float4 hv[8];
        tmem_ld32(trow + c, v);
        gt_stage_write(stg, lane, v);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) hv[it] = gt_stage_read(stg, it * 4 + rsub, j);
        __syncwarp();
        tmem_ld32(trow + 64 + c, v);
        gt_stage_write(stg, lane, v);
        __syncwarp();
        const float4 bh = *reinterpret_cast<const float4*>(e.bias + n0 + c + 4 * j);
        const float4 bg = *reinterpret_cast<const float4*>(e.bias + n0 + 64 + c + 4 * j);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rsub;
          const int oo = __shfl_sync(0xffffffffu, o32, rr);
          const float4 gv = gt_stage_read(stg, rr, j);
          if (oo >= 0) {
            float o[4];
            o[0] = (hv[it].x + bh.x) * gelu_act<T>(gv.x + bg.x);
            o[1] = (hv[it].y + bh.y) * gelu_act<T>(gv.y + bg.y);
            o[2] = (hv[it].z + bh.z) * gelu_act<T>(gv.z + bg.z);
            o[3] = (hv[it].w + bh.w) * gelu_act<T>(gv.w + bg.w);
            store_vec<T, 4>((T*)e.out + (long long)oo * e.ldo + tile_n * 64 + c + 4 * j, o);
          }
        }
        __syncwarp();
      }
    }
  } else {
    if (BN < 64 && half != 0) return;
    switch (e.kind) {
#define EDV_EPI_CASE(F_) case F_: gt_epi_rows<T, BN, F_>(e, trow, m32, o32, n0, stg, lane, half, bias_s); break;
      EDV_EPI_CASE(EF_BIAS)                                              // qkv, projects, out_conv, output_conv1
      EDV_EPI_CASE(EF_BIAS | EF_GELU)                                    // fc1
      EDV_EPI_CASE(EF_BIAS | EF_RES1_F32 | EF_OUT_F32)                   // attn.proj, fc2, temporal to_out: x += ...
      EDV_EPI_CASE(EF_BIAS | EF_OUT_F32)                                 // temporal proj_in
      EDV_EPI_CASE(EF_BIAS | EF_RES1_F32)                                // temporal ff.net.2
      EDV_EPI_CASE(EF_BIAS | EF_RES1_T)                                  // temporal proj_out, RCU conv2
      EDV_EPI_CASE(EF_BIAS | EF_RES1_T | EF_OUT_RELU)
      EDV_EPI_CASE(EF_BIAS | EF_RES1_T | EF_RES2_T)
      EDV_EPI_CASE(EF_BIAS | EF_RES1_T | EF_RES2_T | EF_OUT_RELU)
      EDV_EPI_CASE(EF_BIAS | EF_RELU)                                    // RCU conv1
      EDV_EPI_CASE(EF_OUT_RELU)                                          // layer*_rn
      EDV_EPI_CASE(0)                                                    // temporal q|k|v (no bias)
#undef EDV_EPI_CASE
      default: gt_epi_rows<T, BN, -1>(e, trow, m32, o32, n0, stg, lane, half, bias_s); break;
    }
  }
}

template <typename T, int BN, int BK, bool CONV>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, Epi e, int M,
                                                                int N, int K, int stages, ConvTile ct, int n_tiles,
                                                                int total_tiles) {
  constexpr uint32_t ROW_BYTES = BK * 2;                    // 128 (BK=64) or 64 (BK=32)
  constexpr uint32_t SWZ = (BK == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * ROW_BYTES;                   // 8-row swizzle atom
  constexpr uint32_t A_BYTES = GT_BM * ROW_BYTES;
  constexpr uint32_t B_BYTES = BN * ROW_BYTES;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = gt_tmem_cols<BN>();
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32,256]");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;   // 2: accumulator buffer b complete
  uint64_t* tempty_bar = tfull_bar + 2;       // 2: accumulator buffer b drained by its epilogue warpgroup
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint32_t* stg_base = tmem_slot + 4;          // 16 warps x GT_STG_WORDS
  float* bias_base = reinterpret_cast<float*>(stg_base + 16 * GT_STG_WORDS);   // 16 warps x 128 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 256);
    }
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
      uint32_t kc = 0;  // k-blocks issued so far (ring position)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tile_n = tile % n_tiles, tile_m = tile / n_tiles;
        int cf = 0, y0 = 0, x0 = 0;
        if (CONV) {
          const int per = ct.tiles_y * ct.tiles_x;
          cf = tile_m / per;
          const int r = tile_m - cf * per;
          const int tyi = r / ct.tiles_x;
          y0 = tyi * ct.th;
          x0 = (r - tyi * ct.tiles_x) * ct.tw;
        }
        for (int kb = 0; kb < kblocks; ++kb, ++kc) {
          const int s = kc % stages;
          mbar_wait(&empty_bar[s], ((kc / stages) & 1) ^ 1);
          unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          if (CONV) {
            const int tap = kb / cblocks;
            const int c0 = (kb - tap * cblocks) * BK;
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d(sa, &tmA, &full_bar[s], c0, x0 + kx - 1, y0 + ky - 1, cf);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, tile_m * GT_BM);
          }
          tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop: all 32 lanes wait on the barriers and build the descriptors, the elected lane
    // issues tcgen05.mma / tcgen05.commit (see elect_one_sync in tc_common.cuh)
    constexpr uint32_t idesc = make_idesc<T>(GT_BM, BN, 0);
    const uint32_t leader = elect_one_sync();
    uint32_t kc = 0, it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1;
      mbar_wait(&tempty_bar[b], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
      fence_after_sync();
      const uint32_t acc = tmem_base + b * BN;
      for (int kb = 0; kb < kblocks; ++kb, ++kc) {
        const int s = kc % stages;
        mbar_wait(&full_bar[s], (kc / stages) & 1);
        fence_after_sync();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa, SBO, 16, SWZ);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, SBO, 16, SWZ);
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the >>4 address field
            mma_ss(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          mma_commit(&empty_bar[s]);  // stage reusable once these MMAs have read it
        }
        __syncwarp();
      }
      if (leader) mma_commit(&tfull_bar[b]);    // accumulator complete
      __syncwarp();
    }
  } else if (warp >= 4) {
    // epilogue warpgroups 0..3: accumulator buffer g = wg & 1, column half = wg >> 1;
    // warp w may touch TMEM lanes [32*(w%4), +32)
    const uint32_t wg = (warp - 4) >> 2;
    const uint32_t g = wg & 1;
    const int half = wg >> 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the tile
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != g) continue;
      const int tile_n = tile % n_tiles, tile_m = tile / n_tiles;
      long long m, orow;
      bool valid;
      if (CONV) {
        const int per = ct.tiles_y * ct.tiles_x;
        const int cf = tile_m / per;
        const int rr = tile_m - cf * per;
        const int tyi = rr / ct.tiles_x;
        const int dy = r / ct.tw, dx = r - dy * ct.tw;
        const int y = tyi * ct.th + dy, x = (rr - tyi * ct.tiles_x) * ct.tw + dx;
        valid = (y < ct.H) && (x < ct.W);
        m = ((long long)cf * ct.H + y) * ct.W + x;
        orow = m;
      } else {
        m = (long long)tile_m * GT_BM + r;
        valid = m < M;
        orow = epi_row(e, m);
      }
      // stage this warp's bias slice (columns [half*BN/2, +BN/2) of the tile) while the mainloop runs
      float* bias_s = nullptr;
      if (e.bias && e.act != ACT_GEGLU && e.act != ACT_HEAD) {
        constexpr int CH0 = (BN >= 64) ? BN / 2 : BN;
        bias_s = bias_base + (warp - 4) * 128;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < (CH0 + 31) / 32; ++k)
          if (lane + 32 * k < CH0) bias_s[lane + 32 * k] = __ldg(e.bias + tile_n * BN + half * CH0 + lane + 32 * k);
        __syncwarp();
      }
      mbar_wait(&tfull_bar[g], (it >> 1) & 1);
      fence_after_sync();
      gt_epilogue<T, BN>(e, tmem_base + ((uint32_t)(q * 32) << 16) + g * BN, valid, m, orow, tile_n * BN, tile_n,
                         stg_base + (warp - 4) * GT_STG_WORDS, lane, half, bias_s);
      fence_before_sync();
      mbar_arrive(&tempty_bar[g]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
