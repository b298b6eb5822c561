// tcgen05 flash attention for the spatial (per-frame) self-attention of the ViT encoder:
//   out[f, i, h, :] = softmax_j(q[f,i,h,:] . k[f,j,h,:]) v[f,j,h,:]     head dim 64, S ~ 1.4k tokens
// Replaces Attention.forward's q@k^T / softmax / @v (layers/attention.py:60-66); q is
// pre-scaled by 64^-0.5 at pack time, so the kernel applies no extra scale.
//
// One CTA handles 256 queries (two 128-row tiles, A and B) of one (frame, head) and streams the
// keys/values of that (frame, head) in 128-row tiles.  384 threads = 3 warpgroups; setmaxnreg
// moves registers from warpgroup 0 (56/thread) to the softmax warpgroups (224/thread):
//   warp 0       TMA producer: Q tiles once, then K_j | V_j into a 4-stage 128B-swizzled ring
//   warp 1       TMEM allocator + MMA issuer (one lane):  S_t = Q_t K_j^T  (SS, 128x128x64) and
//                O_t += P_t V_j (A = P from TMEM, B = V tile read MN-major, 128x64x128)
//   warps 2,3    idle (fill warpgroup 0)
//   warps 4..7   softmax warpgroup of tile A: thread = query row.  tcgen05.ld the 128 scores of
//   warps 8..11  softmax warpgroup of tile B  the row into registers, row max / exp2 / row sum
//                entirely thread-local (no shuffles), P rounded to 16 bit and written back to
//                TMEM with tcgen05.st, O rescaled in TMEM only when the running max grew by
//                more than 2^8 (lazy rescale), final O / l written straight to global.
// Software pipeline: S_t(j+1) = Q_t K_{j+1}^T is issued as soon as warpgroup t has pulled S_t(j)
// into registers (barrier s_free), i.e. it runs UNDER the softmax of S_t(j); O_t += P_t(j) V_j
// follows when P_t(j) is in TMEM (p_full) and its completion (pv_done) gates only the next P
// store / O rescale.  So a warpgroup's iteration is just its softmax; the kernel is MUFU(ex2)
// bound at head dim 64 in theory; in practice (ncu + in-kernel timers) it is bound by
// the tensor pipe running these small MMAs at ~45 % of their floor plus softmax issue latency.
//
// TMEM columns (512): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512)
#pragma once
#include "attention_simt.cuh"
#include "launch.h"
#include <cstdlib>

#include "tc_common.cuh"

namespace tc {

constexpr int FA_BM = 128;      // query rows per tile
constexpr int FA_BN = 128;      // keys per iteration
constexpr int FA_HD = 64;
constexpr int FA_STAGES = 4;
constexpr int FA_THREADS = 384;
constexpr int FA_POLY_DEFAULT = 0;   // measured: the polynomial path only adds issue pressure (0: all MUFU)
constexpr uint32_t FA_TILE_BYTES = FA_BM * FA_HD * 2;  // 16 KB: one 128 x 64 16-bit tile
constexpr size_t FA_SMEM = 1024 + (2 + 2 * FA_STAGES) * (size_t)FA_TILE_BYTES + 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-4
// polynomial for 2^f (max relative error 7e-6, far below the 16-bit rounding of P), exponent added
// with an integer shift.  Used for every PM-th score so that the ex2 work is shared between the
// XU pipe (16 lanes/clk/SM) and the otherwise idle FMA pipe -- softmax at head dim 64 is MUFU-bound.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;              // 1.5 * 2^23: n lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.009666374f, f, 0.055838343f);
  p = fmaf(p, f, 0.24022348f);
  p = fmaf(p, f, 0.69313675f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack_pair(float a, float b, bf16) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_pair(float a, float b, f16) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <typename T, int PM>
__global__ void __launch_bounds__(FA_THREADS, 1)
    flash_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, T* __restrict__ out, int S, int heads) {
  extern __shared__ __align__(1024) unsigned char fa_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = smem;                                   // 2 tiles
  unsigned char* sKV = smem + 2 * FA_TILE_BYTES;              // stages x (K | V)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + (size_t)FA_STAGES * 2 * FA_TILE_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // FA_STAGES
  uint64_t* kv_empty = kv_full + FA_STAGES;
  uint64_t* s_full = kv_empty + FA_STAGES; // 2
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* o_done = p_full + 2;           // 2
  uint64_t* s_free = o_done + 2;           // 2: warpgroup t holds S_t(j) in registers
  uint64_t* pv_done = s_free + 2;          // 2: O_t += P_t(j) V_j has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * (2 * FA_BM);
  const int D = heads * FA_HD;
  const bool b_active = (q0 + FA_BM) < S;        // tile B holds at least one real query
  const int n_iter = (S + FA_BN - 1) / FA_BN;
  const int row_base = f * S;                    // first token row of this frame in qkv / out

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_done[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(q_full, (b_active ? 2u : 1u) * FA_TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, q_full, h * FA_HD, row_base + q0);
      if (b_active) tma_load_2d(sQ + FA_TILE_BYTES, &tmQKV, q_full, h * FA_HD, row_base + q0 + FA_BM);
      for (int j = 0; j < n_iter; ++j) {
        const int s = j % FA_STAGES;
        mbar_wait(&kv_empty[s], ((j / FA_STAGES) & 1) ^ 1);
        unsigned char* sk = sKV + (size_t)s * 2 * FA_TILE_BYTES;
        mbar_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
        tma_load_2d(sk, &tmQKV, &kv_full[s], D + h * FA_HD, row_base + j * FA_BN);
        tma_load_2d(sk + FA_TILE_BYTES, &tmQKV, &kv_full[s], 2 * D + h * FA_HD, row_base + j * FA_BN);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc<T>(FA_BM, FA_BN, 0);   // K-major A and B
      constexpr uint32_t idesc_pv = make_idesc<T>(FA_BM, FA_HD, 1);   // B = V tile, MN-major
      const int nt = b_active ? 2 : 1;
      auto issue_qk = [&](int t, int s) {
        const uint32_t qa = smem_u32(sQ + (size_t)t * FA_TILE_BYTES);
        const uint32_t ka = smem_u32(sKV + (size_t)s * 2 * FA_TILE_BYTES);
        const uint64_t adesc = make_smem_desc(qa, 1024, 16, SWZ_128B);
        const uint64_t bdesc = make_smem_desc(ka, 1024, 16, SWZ_128B);
#pragma unroll
        for (int k = 0; k < FA_HD / 16; ++k)
          mma_ss(tmem_base + t * FA_BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_qk, k ? 1u : 0u);
        mma_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      fence_after_sync();
      for (int t = 0; t < nt; ++t) issue_qk(t, 0);
      for (int j = 0; j < n_iter; ++j) {
        const int s = j % FA_STAGES;
        const uint32_t va = smem_u32(sKV + (size_t)s * 2 * FA_TILE_BYTES + FA_TILE_BYTES);
        // 1) S_t(j+1): as soon as warpgroup t holds S_t(j) in registers and K_{j+1} has landed
        if (j + 1 < n_iter) {
          const int s1 = (j + 1) % FA_STAGES;
          mbar_wait(&kv_full[s1], ((j + 1) / FA_STAGES) & 1);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&s_free[t], j & 1);
            fence_after_sync();
            issue_qk(t, s1);
          }
        }
        // 2) O_t (+)= P_t(j) V_j : A from TMEM (8 columns per 16-key step), B rows = keys (MN-major,
        //    128 B per key, 8-key groups 1024 B apart) -> +2048 B per 16-key step
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&p_full[t], j & 1);   // P_t(j) in TMEM, O_t rescaled
          fence_after_sync();
          const uint64_t vdesc = make_smem_desc(va, 1024, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < FA_BN / 16; ++k)
            mma_ts(tmem_base + 256 + t * FA_HD, tmem_base + 384 + t * 64 + k * 8, vdesc + (uint64_t)(128 * k), idesc_pv,
                   (j | k) ? 1u : 0u);
          mma_commit(&pv_done[t]);
          if (t == nt - 1) mma_commit(&kv_empty[s]);   // K_j and V_j no longer needed once these complete
          if (j + 1 == n_iter) mma_commit(&o_done[t]);
        }
      }
    }
   }
  } else {
    // ===== softmax warpgroups =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;        // 0: tile A, 1: tile B
    if (t == 0 || b_active) {
      const int q = warp & 3;             // TMEM lane quarter this warp may access
      const int r = q * 32 + lane;        // row inside the tile
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + t * FA_BN;
      const uint32_t tO = tmem_base + lane_off + 256 + t * FA_HD;
      const uint32_t tP = tmem_base + lane_off + 384 + t * 64;
      constexpr float LOG2E = 1.4426950408889634f;
      float m_used = -INFINITY;           // max the exponentials are currently taken against (raw units)
      float l = 0.f;
      for (int j = 0; j < n_iter; ++j) {
        mbar_wait(&s_full[t], j & 1);
        fence_after_sync();
        uint32_t sv[128];
        tmem_ld32_nowait(tS, sv);
        tmem_ld32_nowait(tS + 32, sv + 32);
        tmem_ld32_nowait(tS + 64, sv + 64);
        tmem_ld32_nowait(tS + 96, sv + 96);
        tmem_ld_wait();
        fence_before_sync();
        mbar_arrive(&s_free[t]);          // S_t(j) is in registers: the tensor core may overwrite it with S_t(j+1)
        const int valid = S - j * FA_BN;  // keys of this tile that belong to the frame
        if (valid < FA_BN) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) sv[i] = 0xff800000u;  // -inf
        }
        // 8 independent partial maxima (a single running max would be a 127-deep dependent chain)
        float pm[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pm[i] = __uint_as_float(sv[i]);
#pragma unroll
        for (int i = 8; i < 128; ++i) pm[i & 7] = fmaxf(pm[i & 7], __uint_as_float(sv[i]));
        const float mx = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])), fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));
        // lazy rescale: keep the old reference max unless the row max grew by more than 2^8
        float factor = 1.f;
        const bool grow = mx > m_used + 8.f / LOG2E;
        if (grow) {
          factor = ex2_approx((m_used - mx) * LOG2E);   // 0 on the first tile (m_used = -inf)
          m_used = mx;
          l *= factor;
        }
        if (j > 0) {
          mbar_wait(&pv_done[t], (j - 1) & 1);   // O_t += P_t(j-1) V_{j-1} done: O_t stable, P_t free
          fence_after_sync();
        }
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
#pragma unroll
          for (int c = 0; c < FA_HD; c += 32) {
            uint32_t ov[32];
            tmem_ld32_nowait(tO + c, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * factor);
            tmem_st32(tO + c, ov);
          }
        }
        const float neg = -m_used * LOG2E;
        float ls[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial row sums
#pragma unroll
        for (int c = 0; c < 128; c += 64) {
          uint32_t pv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x0 = fmaf(__uint_as_float(sv[c + 2 * i]), LOG2E, neg);
            const float x1 = fmaf(__uint_as_float(sv[c + 2 * i + 1]), LOG2E, neg);
            const float p0 = (PM > 0 && ((2 * i) % (PM > 0 ? PM : 1)) == 0) ? ex2_poly(x0) : ex2_approx(x0);
            const float p1 = (PM > 0 && ((2 * i + 1) % (PM > 0 ? PM : 1)) == 0) ? ex2_poly(x1) : ex2_approx(x1);
            ls[i & 3] += p0 + p1;
            pv[i] = pack_pair(p0, p1, T());
          }
          tmem_st32(tP + c / 2, pv);
        }
        l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        tmem_st_wait();
        fence_before_sync();
        mbar_arrive(&p_full[t]);
      }
      // ---- epilogue: O / l -> global ----
      mbar_wait(&o_done[t], 0);
      fence_after_sync();
      const float inv = 1.f / l;
      const int qi = q0 + t * FA_BM + r;
      T* orow = out + ((long long)row_base + qi) * D + h * FA_HD;
#pragma unroll
      for (int c = 0; c < FA_HD; c += 32) {
        uint32_t ov[32];
        tmem_ld32_nowait(tO + c, ov);
        tmem_ld_wait();
        if (qi < S) {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(ov[i]) * inv;
          store_vec<T, 32>(orow + c, o);
        }
      }
      fence_before_sync();
    }
  }
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename T>
void launch_attention_tc(edv::Launch& L, int dtype, const void* qkv, void* out, int F, int S, int heads,
                         bool (*make_map)(edv::Launch&, CUtensorMap*, int, const void*, int, const uint64_t*,
                                          const uint64_t*, const uint32_t*, int)) {
  const int D = heads * FA_HD;
  CUtensorMap tm;
  uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)F * S};
  uint64_t str[1] = {(uint64_t)3 * D * 2};
  uint32_t box[2] = {(uint32_t)FA_HD, (uint32_t)FA_BM};
  if (!make_map(L, &tm, dtype, qkv, 2, dims, str, box, 128)) return;
  auto kern = flash_attention_tc_kernel<T, FA_POLY_DEFAULT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM);
    attr_done = true;
  }
  dim3 grid((S + 2 * FA_BM - 1) / (2 * FA_BM), heads, F);
  kern<<<grid, FA_THREADS, FA_SMEM, L.stream>>>(tm, (T*)out, S, heads);
  L.check("flash_attention_tc");
}

}  // namespace tc
