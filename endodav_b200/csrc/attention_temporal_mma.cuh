// Temporal (frame-axis) attention of the motion modules for the 16-bit paths -- warp-level tensor-core
// version of the short-sequence kernel in attention_simt.cuh (which stays the fp32 path and the cross-check).
// Replaces TemporalAttention / CrossAttention._attention (motion_module.py:232-295; motion_module/attention.py:182-211).
//
//   qkv : [B*T*hw, 3C] row = (clip b, frame f, position d), q|k|v contiguous (q pre-scaled by hd^-0.5), 8 heads
//   out : [B*T*hw, C]
//   grid (ceil(hw/PB), B); a CTA owns PB consecutive positions x all 8 heads; 8 warps, warp = (position, head) task.
//   1. the T frames' rows of the PB positions (3C contiguous elements each) are staged in shared memory with
//      batched 16-byte loads and 16-byte stores: Q, K and V row-major [frame][C+8] -- the 16-byte row padding
//      makes every fragment load (32-bit for Q/K, ldmatrix.trans for V) bank-conflict free;
//   2. per task and 16-query tile: S = Q K^T with mma.sync m16n8k8 (fp32 accumulate) -- the whole T x T
//      (T <= 32) score tile lives in 16 registers per lane; softmax over the key axis with two quad shuffles
//      per statistic; the C fragments of S are, unchanged, the A fragments of P for O = P V (4 k-steps);
//   3. O overwrites the task's own q slots and leaves through 16-byte stores.
//   The T x T problem is far too small for tcgen05 (M = 128 tiles, TMEM round trip per task); mma.sync keeps
//   everything in registers, which is what this shape wants.
#pragma once
#include "common.cuh"

namespace tmma {

template <typename T> struct Mma;
template <> struct Mma<f16> {
  static __device__ __forceinline__ void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static constexpr bool kSplitP = false;   // 11-bit probabilities are inside the 1e-2 gate
  static __device__ __forceinline__ uint32_t residual(float, float, uint32_t) { return 0u; }
};
template <> struct Mma<bf16> {
  static __device__ __forceinline__ void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  // bf16 keeps 8 bits of a probability; P = P_hi + P_lo (two PV mma per k-step) restores ~16 bits so that the
  // bf16 path stays on its operand-rounding floor (the CUDA-core kernel this replaces kept P in fp32)
  static constexpr bool kSplitP = true;
  static __device__ __forceinline__ uint32_t residual(float lo, float hi, uint32_t packed) {
    const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&packed));
    return pack(lo - r.x, hi - r.y);
  }
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int TA_WARPS = 8;

// four 8x8 b16 tiles, transposed: with V row-major [key][channel] this yields, for the 4 key blocks of one
// 8-channel tile, exactly the mma B fragments (k = key 2*t4,+1 ; n = channel g)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}

template <int HD> constexpr int ta_row() { return 8 * HD + 8; }                       // Q/K row pitch in elements
template <int HD> constexpr size_t ta_smem_per_pos() { return (size_t)(3 * 32 * ta_row<HD>()) * 2; }

template <typename T, int HD, int PB>
__global__ void __launch_bounds__(TA_WARPS * 32) temporal_attention_mma_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn,
                                                                               int hw, const float2* __restrict__ rope) {
  pdl_launch();
  pdl_wait();
  constexpr int C = 8 * HD, ROW = ta_row<HD>();
  constexpr int CPS = C / 8;            // 16-byte chunks per q / k / v segment
  constexpr int CPR = 3 * CPS;          // chunks per (frame, position) row
  extern __shared__ __align__(16) unsigned char ta_raw[];
  T* Qs = reinterpret_cast<T*>(ta_raw);                  // [PB][32][ROW]
  T* Ks = Qs + (size_t)PB * 32 * ROW;                    // [PB][32][ROW]
  T* Vs = Ks + (size_t)PB * 32 * ROW;                    // [PB][32][ROW]
  const int d0 = blockIdx.x * PB, b = blockIdx.y;
  const int npos = min(PB, hw - d0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- 0. frames >= Tn must read as zeros (finite) inside the fragments ----
  if (Tn < 32) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* all = reinterpret_cast<uint4*>(ta_raw);
    const int n16 = (int)((size_t)PB * ta_smem_per_pos<HD>() / 16);
    for (int i = tid; i < n16; i += blockDim.x) all[i] = z;
    __syncthreads();
  }

  // ---- 1. load: 4 independent 16-byte loads in flight per thread ----
  const int total = Tn * npos * CPR;
  const T* gbase = qkv + ((long long)b * Tn * hw + d0) * (3LL * C);
  for (int base = tid; base < total; base += blockDim.x * 4) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * (int)blockDim.x;
      if (i < total) {
        const int row = i / CPR, cc = i - row * CPR;
        const int f = row / npos, p = row - f * npos;
        raw[u] = *reinterpret_cast<const uint4*>(gbase + ((long long)f * hw + p) * (3LL * C) + cc * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * (int)blockDim.x;
      if (i < total) {
        const int row = i / CPR, cc = i - row * CPR;
        const int f = row / npos, p = row - f * npos;
        const int which = cc / CPS, c = (cc - which * CPS) * 8;
        uint4 r = raw[u];
        if (rope && which < 2) {
          // RoPE on q and k (pe='rope', motion_module/attention.py:403-429): channel pair i at frame f rotated by
          // the angle whose (cos, sin) sits in rope[f * C/2 + i]
          const T* e = reinterpret_cast<const T*>(&r);
          const float2* rt = rope + (size_t)f * (C / 2) + c / 2;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 cs = rt[j / 2];
            const float a = to_f<T>(e[j]), bq = to_f<T>(e[j + 1]);
            v[j] = a * cs.x - bq * cs.y;
            v[j + 1] = a * cs.y + bq * cs.x;
          }
          r.x = Mma<T>::pack(v[0], v[1]); r.y = Mma<T>::pack(v[2], v[3]);
          r.z = Mma<T>::pack(v[4], v[5]); r.w = Mma<T>::pack(v[6], v[7]);
        }
        T* dst = (which == 0 ? Qs : which == 1 ? Ks : Vs) + ((size_t)p * 32 + f) * ROW + c;
        *reinterpret_cast<uint4*>(dst) = r;
      }
    }
  }
  __syncthreads();

  // ---- 2. attention: one (position, head) task per warp iteration ----
  const int g = lane >> 2, t4 = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;
  for (int task = warp; task < npos * 8; task += TA_WARPS) {
    const int p = task >> 3, h = task & 7;
    T* q0 = Qs + (size_t)p * 32 * ROW + h * HD;
    const T* k0 = Ks + (size_t)p * 32 * ROW + h * HD;
    const T* v0 = Vs + (size_t)p * 32 * ROW + h * HD;
    const int mtiles = Tn > 16 ? 2 : 1;
    for (int mt = 0; mt < mtiles; ++mt) {
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 8; ++ks) {
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(q0 + (size_t)(mt * 16 + g) * ROW + ks * 8 + 2 * t4);
        const uint32_t a1 = *reinterpret_cast<const uint32_t*>(q0 + (size_t)(mt * 16 + g + 8) * ROW + ks * 8 + 2 * t4);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(k0 + (size_t)(nt * 8 + g) * ROW + ks * 8 + 2 * t4);
          Mma<T>::k8(s[nt], a0, a1, b0);
        }
      }
      // softmax over the keys: lane holds keys nt*8 + 2*t4 (+1) of rows g (s[.][0..1]) and g+8 (s[.][2..3])
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = nt * 8 + 2 * t4;
        if (key >= Tn) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= Tn) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
      uint32_t pa[4][2], pr[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float e0 = ex2((s[nt][0] - m0) * LOG2E), e1 = ex2((s[nt][1] - m0) * LOG2E);
        const float e2 = ex2((s[nt][2] - m1) * LOG2E), e3 = ex2((s[nt][3] - m1) * LOG2E);
        l0 += e0 + e1;
        l1 += e2 + e3;
        pa[nt][0] = Mma<T>::pack(e0, e1);   // A fragment of P for k-step nt: row g / row g+8
        pa[nt][1] = Mma<T>::pack(e2, e3);
        pr[nt][0] = Mma<T>::residual(e0, e1, pa[nt][0]);
        pr[nt][1] = Mma<T>::residual(e2, e3, pa[nt][1]);
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      // O = P V, then normalise; the output overwrites this task's q slots (only this warp reads them)
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t vb[4];   // lane -> key row `lane` (tile lane/8 = key block, row lane%8) of channels nt*8 .. nt*8+7
        ldmatrix_x4_trans(vb, v0 + (size_t)lane * ROW + nt * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          Mma<T>::k8(o, pa[j][0], pa[j][1], vb[j]);
          if (Mma<T>::kSplitP) Mma<T>::k8(o, pr[j][0], pr[j][1], vb[j]);
        }
        __syncwarp();   // every lane has finished reading this m-tile's q fragments (loop above) before they are overwritten
        *reinterpret_cast<uint32_t*>(q0 + (size_t)(mt * 16 + g) * ROW + nt * 8 + 2 * t4) = Mma<T>::pack(o[0] * i0, o[1] * i0);
        *reinterpret_cast<uint32_t*>(q0 + (size_t)(mt * 16 + g + 8) * ROW + nt * 8 + 2 * t4) = Mma<T>::pack(o[2] * i1, o[3] * i1);
      }
    }
  }
  __syncthreads();

  // ---- 3. store ----
  const int ototal = Tn * npos * CPS;
  T* obase = out + ((long long)b * Tn * hw + d0) * (long long)C;
  for (int i = tid; i < ototal; i += blockDim.x) {
    const int row = i / CPS, c = (i - row * CPS) * 8;
    const int f = row / npos, p = row - f * npos;
    *reinterpret_cast<uint4*>(obase + ((long long)f * hw + p) * C + c) =
        *reinterpret_cast<const uint4*>(Qs + ((size_t)p * 32 + f) * ROW + c);
  }
}

}  // namespace tmma
