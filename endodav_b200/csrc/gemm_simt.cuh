// CUDA-core GEMM / implicit-GEMM convolution: C[M,N] = A[M,K] * W[N,K]^T, fp32 accumulate.
// This is the fp32 ("tight") path of the engine and the on-device cross-check for the
// tcgen05 kernels in gemm_tc.cuh.  A is either a plain row-major matrix or an NHWC
// activation gathered as 3x3 taps (pad 1, stride 1 or 2) with K = 9*Cin in (ky,kx,c) order.
#pragma once
#include "common.cuh"

struct ConvGeom {
  int H, W, C;      // input NHWC spatial size and channels
  int OH, OW;       // output spatial size
  int stride;       // 1 or 2 (resize_layers[3], dpt.py:85-90)
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename T, bool CONV>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ Wt, Epi e,
                                                       int M, int N, int K, long long lda, ConvGeom g) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * SG_BM;   // M tiles on grid.x: 32 x 518 x 518 output pixels are 134 k tiles (> 65535)
  const int n0 = blockIdx.y * SG_BN;
  // loader mapping: 64 rows x 16 k, 4 consecutive k per thread
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const long long am = m0 + lr;
  const int wn = n0 + lr;

  // conv: decompose the output pixel once
  int cf = 0, coy = 0, cox = 0;
  if (CONV) {
    long long ohw = (long long)g.OH * g.OW;
    cf = (int)(am / ohw);
    int r = (int)(am - (long long)cf * ohw);
    coy = r / g.OW;
    cox = r - coy * g.OW;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + lk;
    if (am < M && k < K) {
      if (CONV) {
        int tap = k / g.C, c = k - tap * g.C;
        int ky = tap / 3, kx = tap - ky * 3;
        int iy = coy * g.stride + ky - 1, ix = cox * g.stride + kx - 1;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
          load_vec<T, 4>(A + (((long long)cf * g.H + iy) * g.W + ix) * g.C + c, av);
      } else {
        load_vec<T, 4>(A + am * lda + k, av);
      }
    }
    if (wn < N && k < K) load_vec<T, 4>(Wt + (long long)wn * K + k, wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = av[i];
      Ws[lk + i][lr] = wv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    epi_apply<T, 4>(e, m, epi_row(e, m), n, acc[i]);
  }
}

// value * gelu(gate) for the tile-paired GEGLU weight layout produced by pack.py:
// columns are grouped in blocks of 128 = [64 value | 64 gate]  (attention.py:382-384).
template <typename T>
__global__ void geglu_pair_kernel(const T* __restrict__ hg, T* __restrict__ out, long long M, int C4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = M * C4;
  if (i >= total) return;
  long long m = i / C4;
  int c = (int)(i - m * C4);
  int blk = c >> 6, j = c & 63;
  const T* row = hg + m * (2LL * C4) + blk * 128;
  out[i] = from_f<T>(to_f<T>(row[j]) * gelu_erf(to_f<T>(row[64 + j])));
}

// disparity head tail on a [M,32] ReLU'd feature map: relu(sum_c x*w + b)  (dpt.py:121-123)
template <typename T>
__global__ void head_dot_kernel(const T* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                long long M, int C, int final_relu, float sig_sign, int sigmoid) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = __ldg(w + C);  // w[C] carries the 1x1 conv bias
  for (int c = 0; c < C; c += 8) {
    float v[8];
    load_vec<T, 8>(x + m * C + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(v[i], __ldg(w + c + i), s);
  }
  if (final_relu) s = fmaxf(s, 0.f);
  if (sigmoid) s = 1.f / (1.f + expf(-sig_sign * s));
  out[m] = s;
}
