// Depth conversion and evaluation metrics of the reference's evaluate scripts on the GPU:
//   disp_to_depth   (utils/layers.py:11-20)      -- elementwise, float32 op for op (bit-identical to numpy)
//   compute_errors  (utils/utils.py:112-133) with the per-frame masking / scaling / clamping of
//                   evaluate_depth_video.py:197-204 -- one block per frame, float64 sums in a fixed order
// so that the caller of infer_video_depth can keep the stitched disparity on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// scaled_disp = min_disp + (max_disp - min_disp) * disp ; depth = 1 / scaled_disp  (numpy float32 arithmetic:
// the Python-float constants are rounded to float32 first, no FMA contraction)
__global__ void disp_to_depth_kernel(const float* __restrict__ disp, float* __restrict__ scaled, float* __restrict__ depth,
                                     long long n, float min_disp, float range) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = __fadd_rn(min_disp, __fmul_rn(range, disp[i]));
  if (scaled) scaled[i] = s;
  depth[i] = __fdiv_rn(1.0f, s);
}

constexpr int CE_THREADS = 1024;

// out[frame][8] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, number of valid pixels.
// mask: optional uint8 [frames][hw] (compute_errors' `mask` argument); otherwise valid = gt in (gt_lo, gt_hi)
// (evaluate_depth_video.py:198).  pred is multiplied by pred_scale and clamped to [clamp_lo, clamp_hi] first
// (:199-201; pass pred_scale = 1 and clamp_lo > clamp_hi to skip).
__global__ void __launch_bounds__(CE_THREADS) compute_errors_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                                   const uint8_t* __restrict__ mask, long long hw, float gt_lo,
                                                                   float gt_hi, float pred_scale, float clamp_lo, float clamp_hi,
                                                                   double* __restrict__ out) {
  const int f = blockIdx.x;
  const float* g = gt + (long long)f * hw;
  const float* p = pred + (long long)f * hw;
  const uint8_t* mk = mask ? mask + (long long)f * hw : nullptr;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = threadIdx.x; i < hw; i += CE_THREADS) {
    const float gv = g[i];
    const bool ok = mk ? (mk[i] != 0) : (gv > gt_lo && gv < gt_hi);
    if (!ok) continue;
    float pv = __fmul_rn(p[i], pred_scale);
    if (clamp_lo <= clamp_hi) pv = fminf(fmaxf(pv, clamp_lo), clamp_hi);
    const float th = fmaxf(__fdiv_rn(gv, pv), __fdiv_rn(pv, gv));
    const float d = __fsub_rn(gv, pv);
    const float d2 = __fmul_rn(d, d);
    const float lg = __fsub_rn(logf(gv), logf(pv));
    acc[0] += (double)__fdiv_rn(fabsf(d), gv);
    acc[1] += (double)__fdiv_rn(d2, gv);
    acc[2] += (double)d2;
    acc[3] += (double)__fmul_rn(lg, lg);
    acc[4] += th < 1.25f ? 1.0 : 0.0;
    acc[5] += th < 1.5625f ? 1.0 : 0.0;           // 1.25 ** 2
    acc[6] += th < 1.953125f ? 1.0 : 0.0;         // 1.25 ** 3
    acc[7] += 1.0;
  }
  __shared__ double red[8][CE_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0;
    for (int w = 0; w < CE_THREADS / 32; ++w) v += red[threadIdx.x][w];
    red[threadIdx.x][0] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = red[7][0];
    double* o = out + (long long)f * 8;
    o[0] = red[0][0] / n;                 // 0 / 0 -> NaN, like numpy's mean of an empty selection
    o[1] = red[1][0] / n;
    o[2] = sqrt(red[2][0] / n);
    o[3] = sqrt(red[3][0] / n);
    o[4] = red[4][0] / n;
    o[5] = red[5][0] / n;
    o[6] = red[6][0] / n;
    o[7] = n;
  }
}
