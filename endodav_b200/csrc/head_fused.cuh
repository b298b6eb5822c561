// Fused disparity-head tail (the "memory-bound tail" of the forward):
//   bilinear upsample (align_corners=True) of the F/2-channel map to the output resolution
//   -> 3x3 conv (F/2 -> 32) + bias -> ReLU -> 1x1 conv (32 -> 1) + bias -> ReLU | sigmoid
// Replaces dpt_pyramid.py:90-93 + dpt.py:118-124 (output_conv2) and the tail of HeadDepth
// (models/endodav/layers.py:211-217 + dpt_pyramid.py:104-109).  Unfused, the upsampled map and
// the 32-channel conv output would make three HBM round trips at the full output resolution
// (~75 MB / frame at 518^2); fused, the kernel reads the low-resolution map once and writes one
// float per pixel (6.7 MB / frame).
//
// Persistent CTAs, one per SM, loop over 16 x 8 output-pixel tiles (= the 128 rows of one MMA):
//   warps 0..3   epilogue: tcgen05.ld the 128 x 32 accumulator (lane = pixel), bias/ReLU/dot/ReLU, store
//   warp  4      TMEM allocator + MMA issuer: 9 taps x (CIN/16) tcgen05.mma (M=128, N=32, K=16);
//                the A operand of tap (ky,kx) is simply the halo tile read at a shifted start
//                address -- the halo is stored chunk-planar ([8-channel chunk][pixel] x 16 B), which
//                IS the canonical no-swizzle K-major UMMA layout for any shift (SBO = halo row pitch)
//   warps 5..12  producers: compute the 18 x 10 halo of the upsampled map straight into that
//                shared-memory layout (4-deep ring), zero outside the image (conv padding)
// Measured (round 2): staging the <= 20 x 12 source pixels of a tile in shared memory first (cp.async, double
// buffered) and interpolating from there is SLOWER (615 vs 531 us at 32 x 296^2 -> 518^2): the ~9x re-fetch of source
// pixels through L1/L2 is not what bounds the kernel; the producers' instruction stream is (ncu: 250 M instructions,
// issue slots 45 % busy with only 8 producer warps, top stall long_scoreboard).  A software pipeline over the producers'
// load batches (the next tile's 12 loads issued before the current batch is interpolated, two register sets) is slower too
// (601 us): the 4-deep halo ring already lets the producers run ahead of the MMAs, the extra registers spill.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int HF_TH = 16, HF_TW = 8;              // output tile
constexpr int HF_HH = HF_TH + 2, HF_HW = HF_TW + 2;  // halo
constexpr int HF_PIX = HF_HH * HF_HW;             // 180
constexpr int HF_PLANE = HF_PIX * 16;             // bytes per 8-channel chunk plane
constexpr int HF_THREADS = 13 * 32;
constexpr int HF_PRODUCERS = 8 * 32;
template <int CIN> constexpr int hf_stages() { return CIN <= 32 ? 4 : 2; }   // halo ring depth (shared-memory budget)

template <int CIN> constexpr size_t hf_smem_bytes() {
  return 1024 + (size_t)9 * (CIN / 8) * 512 + hf_stages<CIN>() * (size_t)(CIN / 8) * HF_PLANE + 256;
}

// w00*a + w01*b + w10*c + w11*d on 8 packed 16-bit channels, in packed 16-bit arithmetic
// (1 HMUL2 + 3 HFMA2 per pair; no fp32 round trip: the producers are instruction-bound)
__device__ __forceinline__ uint4 hf_lerp4(const uint4& a, const uint4& b, const uint4& c, const uint4& d, float w00,
                                          float w01, float w10, float w11, f16) {
  const __half2 h00 = __float2half2_rn(w00), h01 = __float2half2_rn(w01), h10 = __float2half2_rn(w10), h11 = __float2half2_rn(w11);
  uint4 o;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  const __half2* pc = reinterpret_cast<const __half2*>(&c);
  const __half2* pd = reinterpret_cast<const __half2*>(&d);
  __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) po[i] = __hfma2(h11, pd[i], __hfma2(h10, pc[i], __hfma2(h01, pb[i], __hmul2(h00, pa[i]))));
  return o;
}
__device__ __forceinline__ uint4 hf_lerp4(const uint4& a, const uint4& b, const uint4& c, const uint4& d, float w00,
                                          float w01, float w10, float w11, bf16) {
  // bf16 has too few mantissa bits for packed accumulation: fp32 lerp, one rounding
  const bf16* ea = reinterpret_cast<const bf16*>(&a);
  const bf16* eb = reinterpret_cast<const bf16*>(&b);
  const bf16* ec = reinterpret_cast<const bf16*>(&c);
  const bf16* ed = reinterpret_cast<const bf16*>(&d);
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    v[i] = w00 * to_f<bf16>(ea[i]) + w01 * to_f<bf16>(eb[i]) + w10 * to_f<bf16>(ec[i]) + w11 * to_f<bf16>(ed[i]);
  uint4 o;
  o.x = pack2(from_f<bf16>(v[0]), from_f<bf16>(v[1]));
  o.y = pack2(from_f<bf16>(v[2]), from_f<bf16>(v[3]));
  o.z = pack2(from_f<bf16>(v[4]), from_f<bf16>(v[5]));
  o.w = pack2(from_f<bf16>(v[6]), from_f<bf16>(v[7]));
  return o;
}

template <typename T, int CIN>
__global__ void __launch_bounds__(HF_THREADS, 1)
    head_fused_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ head_w, float* __restrict__ out, int F, int H1, int W1, int OH, int OW,
                      float sig_sign, int tiles_x, int tiles_y, int total_tiles) {
  pdl_launch();   // PDL: the next kernel of the stream may start its prologue (common.cuh)
  constexpr int NCH = CIN / 8;                     // 16-byte channel chunks per pixel
  constexpr uint32_t W_BYTES = 9 * NCH * 512;
  constexpr uint32_t HALO_BYTES = NCH * HF_PLANE;
  constexpr int HF_STAGES = hf_stages<CIN>();
  extern __shared__ __align__(1024) unsigned char hf_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(hf_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* wsm = smem;
  unsigned char* halo = smem + W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(halo + HF_STAGES * HALO_BYTES);
  uint64_t* halo_full = bars;                    // HF_STAGES
  uint64_t* halo_empty = bars + HF_STAGES;       // HF_STAGES
  uint64_t* acc_full = bars + 2 * HF_STAGES;     // 2
  uint64_t* acc_empty = acc_full + 2;            // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights [32][9*CIN] (ky,kx,c) -> canonical no-swizzle K-major: [tap][chunk][n] x 16 B
  for (int i = threadIdx.x; i < 9 * NCH * 32; i += HF_THREADS) {
    const int n = i & 31;
    const int tc_ = i >> 5;  // tap * NCH + chunk
    const uint4 v = *reinterpret_cast<const uint4*>(w + (size_t)n * (9 * CIN) + (size_t)tc_ * 8);
    *reinterpret_cast<uint4*>(wsm + (size_t)tc_ * 512 + n * 16) = v;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < HF_STAGES; ++b) {
      mbar_init(&halo_full[b], HF_PRODUCERS);
      mbar_init(&halo_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_barrier_init();
  }
  fence_proxy_async();   // weight stores (generic proxy) -> visible to tcgen05 (async proxy)
  if (warp == 4) tmem_alloc(tmem_slot, 64);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // PDL: everything above ran under the previous kernel's tail; its results are visible from here
  const int per_frame = tiles_x * tiles_y;

  if (warp >= 5) {
    // ===== producers =====
    const int pt = threadIdx.x - 5 * 32;
    const float sy = (OH > 1) ? (float)(H1 - 1) / (float)(OH - 1) : 0.f;
    const float sx = (OW > 1) ? (float)(W1 - 1) / (float)(OW - 1) : 0.f;
    // the (halo pixel, channel chunk) tasks of this thread are the same for every tile
    constexpr int NT = (HF_PIX * NCH + HF_PRODUCERS - 1) / HF_PRODUCERS;
    static_assert(NT <= 12, "too many tasks per producer thread");
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t b = it % HF_STAGES;
      const int f = tile / per_frame;
      const int r = tile - f * per_frame;
      const int ty = r / tiles_x;
      const int y0t = ty * HF_TH - 1, x0t = (r - ty * tiles_x) * HF_TW - 1;
      mbar_wait(&halo_empty[b], ((it / HF_STAGES) & 1) ^ 1);
      unsigned char* hb = halo + b * HALO_BYTES;
      const T* xf = x + (size_t)f * H1 * W1 * CIN;
      constexpr int NB = NT < 3 ? NT : 3;   // tasks per batch: all 4*NB loads are issued before the first use
#pragma unroll 1
      for (int u0 = 0; u0 < NT; u0 += NB) {
        uint4 va[NB], vb[NB], vc[NB], vd[NB];
        float w00[NB], w01[NB], w10[NB], w11[NB];
        int dst[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const int t = pt + (u0 + u) * HF_PRODUCERS;
          const int p = t / NCH, kc = t - p * NCH;
          const int hy = p / HF_HW, hx = p - hy * HF_HW;
          const int oy = y0t + hy, ox = x0t + hx;
          dst[u] = (t < HF_PIX * NCH) ? kc * HF_PLANE + p * 16 : -1;
          const bool in = dst[u] >= 0 && oy >= 0 && oy < OH && ox >= 0 && ox < OW;
          // source coordinates exactly as upsample_nhwc_kernel / ATen (align_corners=True, float32)
          const float fy = sy * oy, fx = sx * ox;
          const int ya = (int)fy, xa = (int)fx;
          const float ly = fy - ya, lx = fx - xa;
          w00[u] = in ? (1.f - ly) * (1.f - lx) : 0.f;
          w01[u] = in ? (1.f - ly) * lx : 0.f;
          w10[u] = in ? ly * (1.f - lx) : 0.f;
          w11[u] = in ? ly * lx : 0.f;
          const int yc = min(max(ya, 0), H1 - 1), xc = min(max(xa, 0), W1 - 1);   // clamped: loads stay in bounds, weights are 0 outside
          const int dxo = (xc + 1 < W1) ? CIN : 0, dyo = (yc + 1 < H1) ? W1 * CIN : 0;
          const T* s0 = xf + ((size_t)yc * W1 + xc) * CIN + kc * 8;
          if (dst[u] >= 0) {
            va[u] = *reinterpret_cast<const uint4*>(s0);
            vb[u] = *reinterpret_cast<const uint4*>(s0 + dxo);
            vc[u] = *reinterpret_cast<const uint4*>(s0 + dyo);
            vd[u] = *reinterpret_cast<const uint4*>(s0 + dyo + dxo);
          }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u)
          if (dst[u] >= 0) *reinterpret_cast<uint4*>(hb + dst[u]) = hf_lerp4(va[u], vb[u], vc[u], vd[u], w00[u], w01[u], w10[u], w11[u], T());
      }
      fence_proxy_async();
      mbar_arrive(&halo_full[b]);
    }
  } else if (warp == 4) {
    // ===== MMA issuer (warp-uniform loop, elected lane issues: tc_common.cuh elect_one_sync) =====
    constexpr uint32_t idesc = make_idesc<T>(128, 32, 0);
    const uint32_t leader = elect_one_sync();
    const uint32_t wa = smem_u32(wsm);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1, ph = (it >> 1) & 1;
      const uint32_t hs = it % HF_STAGES;
      mbar_wait(&acc_empty[b], ph ^ 1);
      mbar_wait(&halo_full[hs], (it / HF_STAGES) & 1);
      fence_after_sync();
      const uint32_t ha = smem_u32(halo + hs * HALO_BYTES);
      const uint64_t adesc0 = make_smem_desc(ha, HF_HW * 16, HF_PLANE, 0);
      const uint64_t bdesc0 = make_smem_desc(wa, 128, 512, 0);
      if (leader) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k) {
            // start-address field is in 16-byte units: halo pixel = 16 B, chunk plane = HF_PLANE B, weight chunk = 512 B
            const uint64_t adesc = adesc0 + (uint64_t)((2 * k) * (HF_PLANE / 16) + (ky * HF_HW + kx));
            const uint64_t bdesc = bdesc0 + (uint64_t)((tap * NCH + 2 * k) * 32);
            mma_ss(tmem_base + b * 32, adesc, bdesc, idesc, (tap | k) ? 1u : 0u);
          }
        }
        mma_commit(&halo_empty[hs]);
        mma_commit(&acc_full[b]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue =====
    const int q = warp;                 // warps 0..3 own TMEM lane quarters 0..3
    const int r = q * 32 + lane;
    const int dy = r >> 3, dx = r & 7;
    float bv[32], hw_[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { bv[i] = __ldg(bias + i); hw_[i] = __ldg(head_w + i); }
    const float hb_ = __ldg(head_w + 32);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1;
      const int f = tile / per_frame;
      const int rr = tile - f * per_frame;
      const int ty = rr / tiles_x;
      const int oy = ty * HF_TH + dy, ox = (rr - ty * tiles_x) * HF_TW + dx;
      mbar_wait(&acc_full[b], (it >> 1) & 1);
      fence_after_sync();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + b * 32, v);
      fence_before_sync();
      mbar_arrive(&acc_empty[b]);
      if (oy < OH && ox < OW) {
        float s = hb_;
#pragma unroll
        for (int i = 0; i < 32; ++i) s = fmaf(fmaxf(v[i] + bv[i], 0.f), hw_[i], s);
        if (sig_sign == 0.f) s = fmaxf(s, 0.f);
        else s = 1.f / (1.f + expf(-sig_sign * s));
        out[((size_t)f * OH + oy) * OW + ox] = s;
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace tc
