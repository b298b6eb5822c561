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
//   grid (hw, B, head_groups); block = 32 * HPB threads (one warp per head of the group)
// shared memory: q|k|v rows of the T frames for the HPB heads, padded so that lanes
// (frames) hit different banks when each reads its own row.
// ---------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void temporal_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int hw, int C,
                                          int HPB) {
  extern __shared__ __align__(16) unsigned char tsm_raw[];
  float* sm = reinterpret_cast<float*>(tsm_raw);
  const int d = blockIdx.x, b = blockIdx.y, hg = blockIdx.z;
  const int W3 = 3 * HPB * HD;   // floats per frame row in smem (q|k|v of this head group)
  const int RS = W3 + 1;         // padded row stride
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthreads = blockDim.x;
  const long long ld = 3LL * C;
  const int c0 = hg * HPB * HD;  // first channel of the head group
  // cooperative, coalesced load: for each frame, 3 segments (q,k,v) of HPB*HD contiguous elements
  const int seg = HPB * HD;
  for (int i = threadIdx.x; i < Tn * 3 * seg; i += nthreads) {
    int f = i / (3 * seg);
    int r = i - f * 3 * seg;
    int which = r / seg, c = r - which * seg;
    const T* row = qkv + ((long long)(b * Tn + f) * hw + d) * ld + (long long)which * C + c0 + c;
    sm[f * RS + which * seg + c] = to_f<T>(*row);
  }
  __syncthreads();
  if (warp < HPB) {
    const int hoff = warp * HD;
    float s[32];
    if (lane < Tn) {
      const float* qrow = sm + lane * RS + hoff;
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = 0.f;
      // scores: q_t . k_j  (k_j broadcast across lanes, q_t conflict-free thanks to the padding)
#pragma unroll
      for (int dd = 0; dd < HD; dd += 8) {
        float qv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) qv[u] = qrow[dd + u];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < Tn) {
            const float* krow = sm + j * RS + seg + hoff + dd;
#pragma unroll
            for (int u = 0; u < 8; ++u) s[j] = fmaf(qv[u], krow[u], s[j]);
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < Tn) mx = fmaxf(mx, s[j]);
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
          const float* vrow = sm + j * RS + 2 * seg + hoff;
          const float p = s[j] * inv;
#pragma unroll
          for (int dd = 0; dd < HD; ++dd) o[dd] = fmaf(p, vrow[dd], o[dd]);
        }
      }
      // every lane of this warp has finished reading q (its own row) -> reuse the q slot
      // of the row for the output so the global store below is coalesced.
      __syncwarp(__activemask());
      float* orow = sm + lane * RS + hoff;
#pragma unroll
      for (int dd = 0; dd < HD; ++dd) orow[dd] = o[dd];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tn * seg; i += nthreads) {
    int f = i / seg, c = i - f * seg;
    out[((long long)(b * Tn + f) * hw + d) * C + c0 + c] = from_f<T>(sm[f * RS + c]);
  }
}
