// CUDA-core attention kernels.
//  * spatial_attention_simt: flash-style (online softmax) attention over the ~1.4k patch
//    tokens of one frame, head dim 64, fp32 math.  It is the fp32 path and the on-device
//    cross-check of the tcgen05 kernel in attention_tc.cuh.
//    Replaces Attention.forward's softmax(q k^T) v (layers/attention.py:60-66).
//  * temporal_attention: the dedicated short-sequence kernel for the motion modules:
//    one warp per (clip, position, head), lane = query frame, the T x T (T<=32) score row
//    lives in registers, K/V rows in shared memory, softmax without any cross-lane traffic.
//    Reads q|k|v once and writes o once in the "(b f) d c" layout, i.e. the rearranges of
//    motion_module.py:232,295 and attention.py:93-98,182-211 never materialise.
#pragma once
#include "common.cuh"

// qkv: [F*S, 3*heads*64] token-major (q pre-scaled by 1/8 at pack time); out: [F*S, heads*64]
template <typename T>
__global__ void __launch_bounds__(128) spatial_attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                                    int S, int heads) {
  constexpr int HD = 64, KT = 32;
  __shared__ __align__(16) float Ks[KT][HD];
  __shared__ __align__(16) float Vs[KT][HD];
  const int f = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * 128 + threadIdx.x;
  const int D = heads * HD;
  const long long ld = 3LL * D;
  const T* base = qkv + (long long)f * S * ld;
  float q[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  if (qi < S) {
#pragma unroll
    for (int d = 0; d < HD; d += 8) load_vec<T, 8>(base + (long long)qi * ld + h * HD + d, q + d);
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int k0 = 0; k0 < S; k0 += KT) {
    __syncthreads();
    // 128 threads load 32 keys x 64 dims of K and V: thread -> (key = tid/4, 16 dims)
    {
      int kr = threadIdx.x >> 2, dc = (threadIdx.x & 3) * 16;
      int kg = k0 + kr;
      float kv[16], vv[16];
      if (kg < S) {
        load_vec<T, 16>(base + (long long)kg * ld + D + h * HD + dc, kv);
        load_vec<T, 16>(base + (long long)kg * ld + 2 * D + h * HD + dc, vv);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { kv[i] = 0.f; vv[i] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) { Ks[kr][dc + i] = kv[i]; Vs[kr][dc + i] = vv[i]; }
    }
    __syncthreads();
    const int nk = min(KT, S - k0);
    float s[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
        a = fmaf(q[d], kk.x, a); a = fmaf(q[d + 1], kk.y, a); a = fmaf(q[d + 2], kk.z, a); a = fmaf(q[d + 3], kk.w, a);
      }
      s[j] = (j < nk) ? a : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mnew = fmaxf(mrun, tmax);
    const float corr = __expf(mrun - mnew);  // mrun = -inf on the first tile -> 0
    lrun *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float p = __expf(s[j] - mnew);  // masked keys: exp(-inf) = 0
      lrun += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
        acc[d] = fmaf(p, vv.x, acc[d]); acc[d + 1] = fmaf(p, vv.y, acc[d + 1]);
        acc[d + 2] = fmaf(p, vv.z, acc[d + 2]); acc[d + 3] = fmaf(p, vv.w, acc[d + 3]);
      }
    }
    mrun = mnew;
  }
  if (qi < S) {
    const float inv = 1.f / lrun;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= inv;
    T* o = out + ((long long)f * S + qi) * D + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 8) store_vec<T, 8>(o + d, acc + d);
  }
}

// ---------------------------------------------------------------------------------------
// temporal attention
//   qkv : [B*T*hw, 3C]  row (b*T + f)*hw + d,  columns [q | k | v], head h at h*HD
//   out : [B*T*hw, C]
//   grid (ceil(hw/PB), B, 8/HG); block = 32 * HG * PB threads: one warp per (position, head),
//   lane = query frame.  A CTA owns PB consecutive positions x HG heads:
//     1. cooperative 16-byte loads of the T frames' q|k|v segments (consecutive positions are
//        consecutive rows, so each frame contributes one contiguous run per segment); K and V
//        are converted to fp32 once, q stays in the activation dtype
//     2. per warp: T x T scores in registers (K rows are shared-memory broadcasts), softmax
//        without any cross-lane traffic, P.V with V broadcasts
//     3. the output overwrites the warp's own q slots and leaves through cooperative 16-byte stores
//   The activation is read once and written once; no rearranged copy exists.
// ---------------------------------------------------------------------------------------
template <int HD> constexpr int ta_max_threads() { return HD >= 64 ? 256 : 512; }

template <typename T, int HD>
__global__ void __launch_bounds__(ta_max_threads<HD>()) temporal_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int hw, int C, int PB,
                                          int HG, const float2* __restrict__ rope) {
  constexpr int VEC = 16 / (int)sizeof(T);       // elements per 16-byte chunk
  constexpr int QPAD = 16 / (int)sizeof(T);      // q rows are padded by 16 B: lanes (frames) read their own row
  extern __shared__ __align__(16) unsigned char tsm_raw[];
  const int W = HG * HD;                         // channels of this head group
  const int rowKV = PB * W;                      // floats per frame in Ks / Vs
  const int rowQ = PB * W + QPAD;                // elements per frame in Qs
  float* Ks = reinterpret_cast<float*>(tsm_raw);
  float* Vs = Ks + (size_t)Tn * rowKV;
  T* Qs = reinterpret_cast<T*>(Vs + (size_t)Tn * rowKV);
  const int d0 = blockIdx.x * PB, b = blockIdx.y, hg = blockIdx.z;
  const int npos = min(PB, hw - d0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ld = 3LL * C;
  const int c0 = hg * W;
  const int cpr = W / VEC;                       // 16-byte chunks per (frame, position, segment)

  // ---- 1. load ----
  const int total = Tn * npos * 3 * cpr;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    int c = i % cpr;
    int r = i / cpr;
    const int which = r % 3;
    r /= 3;
    const int pos = r % npos, f = r / npos;
    const T* src = qkv + ((long long)(b * Tn + f) * hw + d0 + pos) * ld + (long long)which * C + c0 + c * VEC;
    const uint4 raw = *reinterpret_cast<const uint4*>(src);
    const T* e = reinterpret_cast<const T*>(&raw);
    float v[VEC];
#pragma unroll
    for (int u = 0; u < VEC; ++u) v[u] = to_f<T>(e[u]);
    if (rope && which < 2) {
      // RoPE on q and k (pe='rope', motion_module/attention.py:403-429): channel pair i at frame f is
      // rotated by the angle whose (cos, sin) sits in rope[f * C/2 + i]
      const float2* rt = rope + (size_t)f * (C / 2) + (c0 + c * VEC) / 2;
#pragma unroll
      for (int u = 0; u < VEC; u += 2) {
        const float2 cs = rt[u / 2];
        const float a = v[u], bq = v[u + 1];
        v[u] = a * cs.x - bq * cs.y;
        v[u + 1] = a * cs.y + bq * cs.x;
      }
    }
    if (which == 0) {
      T* dst = Qs + (size_t)f * rowQ + pos * W + c * VEC;
      if (rope) store_vec<T, VEC>(dst, v);
      else *reinterpret_cast<uint4*>(dst) = raw;
    } else {
      float* dst = (which == 1 ? Ks : Vs) + (size_t)f * rowKV + pos * W + c * VEC;
#pragma unroll
      for (int u = 0; u < VEC; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
    }
  }
  __syncthreads();

  // ---- 2. attention: warp = (position, head), lane = query frame ----
  const int pos = warp / HG, hl = warp - pos * HG;
  if (pos < npos && lane < Tn) {
    const int off = pos * W + hl * HD;
    const T* qrow = Qs + (size_t)lane * rowQ + off;
    float s[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] = 0.f;
#pragma unroll
    for (int dd = 0; dd < HD; dd += 8) {
      float qv[8];
      load_vec<T, 8>(qrow + dd, qv);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < Tn) {
          const float4 k0 = *reinterpret_cast<const float4*>(Ks + (size_t)j * rowKV + off + dd);
          const float4 k1 = *reinterpret_cast<const float4*>(Ks + (size_t)j * rowKV + off + dd + 4);
          float a = s[j];
          a = fmaf(qv[0], k0.x, a); a = fmaf(qv[1], k0.y, a); a = fmaf(qv[2], k0.z, a); a = fmaf(qv[3], k0.w, a);
          a = fmaf(qv[4], k1.x, a); a = fmaf(qv[5], k1.y, a); a = fmaf(qv[6], k1.z, a); a = fmaf(qv[7], k1.w, a);
          s[j] = a;
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < Tn) mx = fmaxf(mx, s[j]);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s[j] = (j < Tn) ? expf(s[j] - mx) : 0.f;
      l += s[j];
    }
    const float inv = 1.f / l;
    float o[HD];
#pragma unroll
    for (int dd = 0; dd < HD; ++dd) o[dd] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < Tn) {
        const float pj = s[j] * inv;
        const float* vrow = Vs + (size_t)j * rowKV + off;
#pragma unroll
        for (int dd = 0; dd < HD; dd += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(vrow + dd);
          o[dd] = fmaf(pj, v4.x, o[dd]); o[dd + 1] = fmaf(pj, v4.y, o[dd + 1]);
          o[dd + 2] = fmaf(pj, v4.z, o[dd + 2]); o[dd + 3] = fmaf(pj, v4.w, o[dd + 3]);
        }
      }
    }
    // only this lane ever read these q slots: reuse them for the output
    T* orow = Qs + (size_t)lane * rowQ + off;
#pragma unroll
    for (int dd = 0; dd < HD; dd += 8) store_vec<T, 8>(orow + dd, o + dd);
  }
  __syncthreads();

  // ---- 3. store ----
  const int ototal = Tn * npos * cpr;
  for (int i = threadIdx.x; i < ototal; i += blockDim.x) {
    const int c = i % cpr;
    const int r = i / cpr;
    const int p2 = r % npos, f = r / npos;
    *reinterpret_cast<uint4*>(out + ((long long)(b * Tn + f) * hw + d0 + p2) * C + c0 + c * VEC) =
        *reinterpret_cast<const uint4*>(Qs + (size_t)f * rowQ + p2 * W + c * VEC);
  }
}
