// On-GPU window stitching: the sequential scale/shift alignment + 8-frame linear
// cross-fade of endodav.infer_video_depth (models/endodav/endodav.py:213-254; utils/util.py:40-74),
// one window per call, stream-ordered, no host synchronisation.
//
// Reference arithmetic that is kept op for op (float32, round-to-nearest, NO fused multiply-add):
//   det = a00*a11 - a01*a01;  scale = (a11*b0 - a01*b1)/det;  shift = (-a01*b0 + a00*b1)/det
//   g = frame*scale + shift;  g[g<0] = 0;  tail = pre*(1-w) + g*w   (w = float32(i*(1/7)))
// The five sums (np.sum of float32 products) are evaluated in float32 in
// numpy's own pairwise association order (see below), so the whole chain is bit-identical to the numpy one.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace edv {

// ---- numpy's float32 pairwise summation, restated ------------------------------------------------------
// np.sum over a contiguous float32 array (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum):
//   n <= 128 : eight strided accumulators r[j] += a[8i+j], res = ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then
//              the n%8 trailing elements are added one by one;
//   n  > 128 : n2 = n/2 rounded down to a multiple of 8; res = sum(a[:n2]) + sum(a[n2:]).
// The split tree depends only on n.  stitch_plan_build() flattens it on the host: the leaves (contiguous
// blocks of 8..128 elements, in order) and the internal nodes sorted by height, so that the device evaluates
// exactly the same float32 additions in the same association order -> the sums are bit-identical to numpy's.
// Plan layout (int32): [0]=L leaves, [1]=I internal nodes, [2]=levels, [3]=root value index,
//   [4 .. 4+levels] level starts (into the internal list), then L+1 leaf offsets, then I (left,right) pairs
//   (value indices: leaves 0..L-1, internal node i -> L+i).
struct StitchNode {
  int l, r, h;
};

static int stitch_plan_rec(long long off, long long n, std::vector<int>& leaf_off, std::vector<StitchNode>& nodes, int* height) {
  if (n <= 128) {
    leaf_off.push_back((int)off);
    *height = 0;
    return (int)leaf_off.size() - 1;            // leaf id >= 0
  }
  long long n2 = n / 2;
  n2 -= n2 % 8;
  int hl, hr;
  const int l = stitch_plan_rec(off, n2, leaf_off, nodes, &hl);
  const int r = stitch_plan_rec(off + n2, n - n2, leaf_off, nodes, &hr);
  *height = 1 + (hl > hr ? hl : hr);
  nodes.push_back(StitchNode{l, r, *height});
  return -(int)nodes.size();                    // internal id: -(index+1)
}

static std::vector<int> stitch_plan_build(long long n) {
  std::vector<int> leaf_off;
  std::vector<StitchNode> nodes;
  int h = 0;
  const int root = stitch_plan_rec(0, n, leaf_off, nodes, &h);
  const int L = (int)leaf_off.size(), I = (int)nodes.size();
  leaf_off.push_back((int)n);
  // stable counting sort of the internal nodes by height (children always have a smaller height)
  std::vector<int> start(h + 2, 0), order(I), pos(I);
  for (const StitchNode& nd : nodes) ++start[nd.h + 1];
  for (int i = 1; i <= h + 1; ++i) start[i] += start[i - 1];
  {
    std::vector<int> fill(start.begin(), start.end());
    for (int i = 0; i < I; ++i) {
      pos[i] = fill[nodes[i].h]++;
      order[pos[i]] = i;
    }
  }
  auto value_index = [&](int id) { return id >= 0 ? id : L + pos[-id - 1]; };
  std::vector<int> plan;
  plan.reserve(4 + h + 1 + L + 1 + 2 * (size_t)I);
  plan.push_back(L);
  plan.push_back(I);
  plan.push_back(h);
  plan.push_back(value_index(root));
  for (int lv = 1; lv <= h + 1; ++lv) plan.push_back(start[lv]);   // h+1 entries: start of level 1..h, then I
  plan.insert(plan.end(), leaf_off.begin(), leaf_off.end());
  for (int p = 0; p < I; ++p) {
    plan.push_back(value_index(nodes[order[p]].l));
    plan.push_back(value_index(nodes[order[p]].r));
  }
  return plan;
}

constexpr int STITCH_THREADS = 256;

// eight lanes per leaf, lane j = numpy's accumulator r[j]; val[leaf] = (sum p*p, sum p, sum p*t, sum t), p = post, t = pre
__global__ void __launch_bounds__(STITCH_THREADS) stitch_leaf_kernel(const float* __restrict__ pre, const float* __restrict__ post,
                                                                     const int* __restrict__ plan, float4* __restrict__ val) {
  const int L = plan[0], levels = plan[2];
  const int* leaf_off = plan + 4 + levels + 1;
  const int leaf = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  const int j = threadIdx.x & 7;
  const bool live = leaf < L;
  const int off = live ? leaf_off[leaf] : 0;
  const int n = live ? leaf_off[leaf + 1] - off : 0;
  const int m = n - (n & 7);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (m > 0) {
    float p = post[off + j], t = pre[off + j];
    s0 = __fmul_rn(p, p); s1 = p; s2 = __fmul_rn(p, t); s3 = t;        // r[j] = a[j]
    for (int i = 8; i < m; i += 8) {
      p = post[off + i + j];
      t = pre[off + i + j];
      s0 = __fadd_rn(s0, __fmul_rn(p, p));
      s1 = __fadd_rn(s1, p);
      s2 = __fadd_rn(s2, __fmul_rn(p, t));
      s3 = __fadd_rn(s3, t);
    }
  }
  // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)): xor-butterfly over the 8 lanes (float addition is commutative)
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    s0 = __fadd_rn(s0, __shfl_xor_sync(0xffffffffu, s0, o));
    s1 = __fadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, o));
    s2 = __fadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, o));
    s3 = __fadd_rn(s3, __shfl_xor_sync(0xffffffffu, s3, o));
  }
  if (live && j == 0) {
    for (int i = m; i < n; ++i) {                                       // the n % 8 trailing elements
      const float p = post[off + i], t = pre[off + i];
      s0 = __fadd_rn(s0, __fmul_rn(p, p));
      s1 = __fadd_rn(s1, p);
      s2 = __fadd_rn(s2, __fmul_rn(p, t));
      s3 = __fadd_rn(s3, t);
    }
    val[leaf] = make_float4(s0, s1, s2, s3);
  }
}

// one CTA: the pairwise tree level by level (children of a level-h node sit on levels < h), then the 2x2 solve
// in float32 exactly as utils/util.py:54-60 writes it
__global__ void __launch_bounds__(1024) stitch_tree_solve_kernel(const int* __restrict__ plan, float4* __restrict__ val, long long n,
                                                                 float* __restrict__ ss) {
  const int L = plan[0], levels = plan[2], root = plan[3];
  const int* lv = plan + 4;                      // nodes of height h: [lv[h-1], lv[h]); lv[levels] == I
  const int* pairs = plan + 4 + levels + 1 + L + 1;
  for (int h = 1; h <= levels; ++h) {
    const int begin = lv[h - 1], end = lv[h];
    for (int i = begin + (int)threadIdx.x; i < end; i += (int)blockDim.x) {
      const float4 a = val[pairs[2 * i]], b = val[pairs[2 * i + 1]];
      val[L + i] = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float4 t = val[root];
    const float a00 = t.x, a01 = t.y, a11 = (float)n, b0 = t.z, b1 = t.w;   // np.sum(ones) == n exactly (n < 2^24)
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
