// On-GPU window stitching: the sequential scale/shift alignment + 8-frame linear
// cross-fade of endodav.infer_video_depth (models/endodav/endodav.py:213-254; utils/util.py:40-74),
// one window per call, stream-ordered, no host synchronisation.
//
// Reference arithmetic that is kept op for op (float32, round-to-nearest, NO fused multiply-add):
//   det = a00*a11 - a01*a01;  scale = (a11*b0 - a01*b1)/det;  shift = (-a01*b0 + a00*b1)/det
//   g = frame*scale + shift;  g[g<0] = 0;  tail = pre*(1-w) + g*w   (w = float32(i*(1/7)))
// The five sums (np.sum of float32 products, numpy's pairwise order) are accumulated here in float64 in
// a fixed block/warp order (bit-reproducible run to run) over the same float32-rounded products and
// then rounded to float32: they agree with numpy's to a few float32 ulps.
#pragma once
#include <cuda_runtime.h>

namespace edv {

constexpr int STITCH_BLOCKS = 296;   // 2 CTAs per SM on 148 SMs
constexpr int STITCH_THREADS = 256;

// partial[b][0..3] = sum p*p, sum p, sum p*t, sum t over block b's grid-stride share (p = post, t = pre)
__global__ void __launch_bounds__(STITCH_THREADS) stitch_stats_kernel(const float* __restrict__ pre,
                                                                      const float* __restrict__ post, long long n,
                                                                      double* __restrict__ partial) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n & 3) == 0 && (((uintptr_t)pre | (uintptr_t)post) & 15) == 0) {
    const float4* p4 = (const float4*)post;
    const float4* t4 = (const float4*)pre;
    for (long long j = i; j < n / 4; j += stride) {
      const float4 p = p4[j], t = t4[j];
      s0 += (double)__fmul_rn(p.x, p.x) + (double)__fmul_rn(p.y, p.y) + (double)__fmul_rn(p.z, p.z) + (double)__fmul_rn(p.w, p.w);
      s1 += (double)p.x + (double)p.y + (double)p.z + (double)p.w;
      s2 += (double)__fmul_rn(p.x, t.x) + (double)__fmul_rn(p.y, t.y) + (double)__fmul_rn(p.z, t.z) + (double)__fmul_rn(p.w, t.w);
      s3 += (double)t.x + (double)t.y + (double)t.z + (double)t.w;
    }
  } else {
    for (long long j = i; j < n; j += stride) {
      const float p = post[j], t = pre[j];
      s0 += (double)__fmul_rn(p, p);
      s1 += (double)p;
      s2 += (double)__fmul_rn(p, t);
      s3 += (double)t;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    s3 += __shfl_down_sync(0xffffffffu, s3, o);
  }
  __shared__ double sm[STITCH_THREADS / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm[warp][0] = s0; sm[warp][1] = s1; sm[warp][2] = s2; sm[warp][3] = s3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0;
    for (int w = 0; w < STITCH_THREADS / 32; ++w) a += sm[w][threadIdx.x];
    partial[blockIdx.x * 4 + threadIdx.x] = a;
  }
}

// one warp: fixed-order reduction of the block partials, then the 2x2 solve in float32 (utils/util.py:54-60)
__global__ void stitch_solve_kernel(const double* __restrict__ partial, int nblocks, long long n, float* __restrict__ ss) {
  __shared__ double tot[4];
  if (threadIdx.x < 4) {
    double a = 0;
    for (int b = 0; b < nblocks; ++b) a += partial[b * 4 + threadIdx.x];
    tot[threadIdx.x] = a;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const float a00 = (float)tot[0], a01 = (float)tot[1], a11 = (float)n, b0 = (float)tot[2], b1 = (float)tot[3];
    const float det = __fsub_rn(__fmul_rn(a00, a11), __fmul_rn(a01, a01));
    float scale = 1.f, shift = 0.f;
    if (det != 0.f) {
      scale = __fdiv_rn(__fsub_rn(__fmul_rn(a11, b0), __fmul_rn(a01, b1)), det);
      shift = __fdiv_rn(__fadd_rn(__fmul_rn(-a01, b0), __fmul_rn(a00, b1)), det);
    }
    ss[0] = scale;
    ss[1] = shift;
  }
}

__global__ void stitch_identity_kernel(float* __restrict__ ss) {
  ss[0] = 1.f;
  ss[1] = 0.f;
}

struct StitchFade {
  float w0[8], w1[8];   // float32(1 - w_i), float32(w_i), w_i computed in double like the Python list
};

// tail: out[tail0 + j] for j < 8 is cross-faded in place, j in [8,30) receives the aligned fresh frames;
// window slot of output j is j + 2 (slots 2..9 = post frames, 10..31 = fresh frames)
__global__ void stitch_apply_kernel(const float* __restrict__ win, float* __restrict__ tail, long long hw,
                                    const float* __restrict__ ss, StitchFade fade) {
  const long long total = 30 * hw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float scale = ss[0], shift = ss[1];
  float g = __fadd_rn(__fmul_rn(win[2 * hw + i], scale), shift);
  if (g < 0.f) g = 0.f;
  const int j = (int)(i / hw);
  if (j < 8) g = __fadd_rn(__fmul_rn(tail[i], fade.w0[j]), __fmul_rn(g, fade.w1[j]));
  tail[i] = g;
}

}  // namespace edv
