// Host-side launchers shared by the forward graph (engine.cu) and the per-kernel C entry
// points.  Everything is stream-ordered; no allocation, no synchronisation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/endodav_b200.h"
#include "launch.h"
#include "attention_simt.cuh"
#include "attention_temporal_mma.cuh"
#include "attention_tc.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "conv_halo.cuh"
#include "head_fused.cuh"

namespace edv {

inline size_t dtype_size(int dtype) { return dtype == EDV_F32 ? 4 : 2; }

// ---- TMA descriptor encoding (driver entry point fetched through the runtime, so the
// library carries no link-time dependency on libcuda) --------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// rank-R tiled map over a 16-bit tensor; dims/box innermost first; strides in bytes for dims 1..R-1
inline bool make_tmap(Launch& L, CUtensorMap* m, int dtype, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  if (!fn) {
    L.fail(EDV_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    return false;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dt = dtype == EDV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu box %u,%u", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    L.fail(EDV_ERR_CUDA, buf);
    return false;
  }
  return true;
}

// ---- GEMM -----------------------------------------------------------------------------------
struct GemmArgs {
  const void* A = nullptr;   // [M,K] (lda) or NHWC activation for conv
  const void* W = nullptr;   // [N,K]
  int M = 0, N = 0, K = 0;
  long long lda = 0;
  Epi e{};
  // conv (3x3, pad 1)
  bool conv = false;
  int F = 0, H = 0, Wd = 0, C = 0, stride = 1;
};

inline Epi epi_zero() {
  Epi e;
  memset(&e, 0, sizeof e);
  e.rb_div = 1;
  e.rb_mod = 1;
  return e;
}

inline void note_gemm(Launch& L, const GemmArgs& a, size_t es) {
  // algorithmic work: 2*M*N*K; bytes: A (conv: the activation once) + W + C once
  const double a_elems = a.conv ? (double)a.F * a.H * a.Wd * a.C : (double)a.M * a.K;
  L.note(2.0 * a.M * a.N * a.K, (a_elems + (double)a.N * a.K + (double)a.M * a.N) * es);
}

template <typename T> void launch_gemm_simt(Launch& L, const GemmArgs& a) {
  note_gemm(L, a, sizeof(T));
  dim3 grid((a.N + SG_BN - 1) / SG_BN, (unsigned)((a.M + SG_BM - 1) / SG_BM));
  ConvGeom g{};
  if (a.conv) {
    g.H = a.H; g.W = a.Wd; g.C = a.C; g.stride = a.stride;
    g.OH = (a.H - 1) / a.stride + 1;
    g.OW = (a.Wd - 1) / a.stride + 1;
    gemm_simt_kernel<T, true><<<grid, 256, 0, L.stream>>>((const T*)a.A, (const T*)a.W, a.e, a.M, a.N, a.K, a.lda, g);
  } else {
    gemm_simt_kernel<T, false><<<grid, 256, 0, L.stream>>>((const T*)a.A, (const T*)a.W, a.e, a.M, a.N, a.K, a.lda, g);
  }
  L.check("gemm_simt");
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// choose the conv tile shape (th*tw == 128) wasting the fewest pixels
inline void pick_conv_tile(int H, int W, int* th, int* tw) {
  const int cand[6][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {1, 128}};
  long long best = -1;
  for (auto& c : cand) {
    long long cover = (long long)((H + c[0] - 1) / c[0]) * c[0] * ((W + c[1] - 1) / c[1]) * c[1];
    if (best < 0 || cover < best) {
      best = cover;
      *th = c[0];
      *tw = c[1];
    }
  }
}

template <typename T, int BN, int BK, bool CONV>
void launch_gemm_tc_inst(Launch& L, int dtype, const GemmArgs& a) {
  using namespace tc;
  CUtensorMap tmA, tmB;
  ConvTile ct{};
  long long m_tiles;
  const int swz = BK * 2;
  if (CONV) {
    ct.H = a.H; ct.W = a.Wd; ct.C = a.C;
    pick_conv_tile(a.H, a.Wd, &ct.th, &ct.tw);
    ct.tiles_y = (a.H + ct.th - 1) / ct.th;
    ct.tiles_x = (a.Wd + ct.tw - 1) / ct.tw;
    m_tiles = (long long)a.F * ct.tiles_y * ct.tiles_x;
    uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.Wd, (uint64_t)a.H, (uint64_t)a.F};
    uint64_t str[3] = {(uint64_t)a.C * 2, (uint64_t)a.C * a.Wd * 2, (uint64_t)a.C * a.Wd * a.H * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)ct.tw, (uint32_t)ct.th, 1};
    if (!make_tmap(L, &tmA, dtype, a.A, 4, dims, str, box, swz)) return;
  } else {
    m_tiles = (a.M + GT_BM - 1) / GT_BM;
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmA, dtype, a.A, 2, dims, str, box, swz)) return;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.K * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    if (!make_tmap(L, &tmB, dtype, a.W, 2, dims, str, box, swz)) return;
  }
  const int kblocks = a.K / BK;
  const int stage_bytes = gt_stage_bytes<BN, BK>();
  const size_t stg_bytes = 16 * (size_t)GT_STG_WORDS * 4 + 16 * 128 * 4;   // + per-warp bias slices   // epilogue transposition buffers
  int stages = (int)((220 * 1024 - stg_bytes - 2048) / stage_bytes);  // one persistent CTA per SM owns the shared memory
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align*/ + (2 * stages + 4) * 8 + 16 + stg_bytes;
  auto kern = gemm_tc_kernel<T, BN, BK, CONV>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    attr_done = true;
  }
  const int n_tiles = a.N / BN;
  const long long total = m_tiles * n_tiles;
  if (total > 0x7fffffffLL) return L.fail(EDV_ERR_ARG, "gemm_tc: too many tiles");
  const int grid = (int)std::min<long long>(total, num_sms());
  (void)kblocks;
  note_gemm(L, a, 2);
  kern<<<grid, GT_THREADS, smem, L.stream>>>(tmA, tmB, a.e, a.M, a.N, a.K, stages, ct, n_tiles, (int)total);
  L.check("gemm_tc");
}

// 64-channel 3x3 convs with halo reuse (conv_halo.cuh)
template <typename T, int BN> void launch_conv_halo(Launch& L, const GemmArgs& a) {
  using namespace tc;
  auto kern = conv3x3_halo_kernel<T, 64, BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch_smem_bytes<64, BN>());
    attr_done = true;
  }
  const int tiles_x = (a.Wd + CH_TW - 1) / CH_TW, tiles_y = (a.H + CH_TH - 1) / CH_TH;
  const long long total = (long long)a.F * tiles_x * tiles_y;
  if (total > 0x7fffffffLL) return L.fail(EDV_ERR_ARG, "conv_halo: too many tiles");
  const int grid = (int)std::min<long long>(total, num_sms());
  note_gemm(L, a, 2);
  kern<<<grid, CH_THREADS, ch_smem_bytes<64, BN>(), L.stream>>>((const T*)a.A, (const T*)a.W, a.e, a.F, a.H, a.Wd, tiles_x,
                                                               tiles_y, (int)total);
  L.check("conv_halo");
}

inline bool conv_halo_ok(const GemmArgs& a) {
  return a.conv && a.stride == 1 && a.C == 64 && (a.N == 64 || a.N == 32) &&
         (a.e.act == ACT_NONE || a.e.act == ACT_RELU || a.e.act == ACT_SIGMOID) && a.e.map == MAP_LINEAR;
}

// 2-SM (cta_group::2) GEMM for the big token GEMMs (gemm_tc2.cuh)
template <typename T, int BN> void launch_gemm_tc2(Launch& L, int dtype, const GemmArgs& a) {
  using namespace tc;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {64u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmA, dtype, a.A, 2, dims, str, box, 128)) return;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.K * 2};
    uint32_t box[2] = {64u, (uint32_t)(BN / 2)};
    if (!make_tmap(L, &tmB, dtype, a.W, 2, dims, str, box, 128)) return;
  }
  const size_t stg_bytes = 16 * (size_t)GT_STG_WORDS * 4 + 16 * 128 * 4;
  const int stage_bytes = gt2_stage_bytes<BN>();
  int stages = (int)((220 * 1024 - stg_bytes - 2048) / stage_bytes);
  if (stages > 8) stages = 8;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * stages + 4) * 8 + 16 + stg_bytes;
  auto kern = gemm_tc2_kernel<T, BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    attr_done = true;
  }
  const int n_tiles = a.N / BN;
  const long long total = (long long)((a.M + 255) / 256) * n_tiles;
  int pairs = (int)std::min<long long>(total, num_sms() / 2);
  note_gemm(L, a, 2);
  kern<<<2 * pairs, GT_THREADS, smem, L.stream>>>(tmA, tmB, a.e, a.M, a.N, a.K, stages, n_tiles, (int)total);
  L.check("gemm_tc2");
}

constexpr int GEMM_2SM_DEFAULT_MIN_M_DECL = 0;
// EDV_GEMM_2SM=<min M> routes linear GEMMs with at least that many rows to the 2-SM kernel (0 = off)
inline int gemm_2sm_min_m() {
  static int v = -1;
  if (v < 0) {
    const char* env = getenv("EDV_GEMM_2SM");
    v = env ? atoi(env) : GEMM_2SM_DEFAULT_MIN_M_DECL;
    if (v < 0) v = 0;
  }
  return v;
}


template <typename T> void launch_gemm_tc(Launch& L, int dtype, const GemmArgs& a) {
  if (!a.conv && gemm_2sm_min_m() > 0 && a.M >= gemm_2sm_min_m() && a.K % 64 == 0 && a.lda == a.K && a.e.act != ACT_GEGLU &&
      a.e.act != ACT_HEAD) {
    if (a.N % 256 == 0) return launch_gemm_tc2<T, 256>(L, dtype, a);
    if (a.N % 192 == 0) return launch_gemm_tc2<T, 192>(L, dtype, a);
  }
  if (conv_halo_ok(a)) {
    if (a.N == 64) launch_conv_halo<T, 64>(L, a);
    else launch_conv_halo<T, 32>(L, a);
    return;
  }
  const bool small_k = a.conv ? (a.C % 64 != 0) : (a.K % 64 != 0);
  const int bk = small_k ? 32 : 64;
  if ((a.conv ? a.C : a.K) % bk != 0) return L.fail(EDV_ERR_ARG, "gemm_tc: K (or conv C) must be a multiple of 32");
  if (a.conv && a.stride != 1) return L.fail(EDV_ERR_ARG, "gemm_tc: conv stride must be 1");
  int bn;
  if (a.e.act == ACT_GEGLU) bn = 128;
  else if (a.e.act == ACT_HEAD) bn = 32;
  else if (bk == 64 && a.N % 256 == 0) bn = 256;   // wide tiles halve the shared-memory traffic per FLOP
  else if (bk == 64 && a.N % 192 == 0) bn = 192;
  else bn = (a.N % 128 == 0) ? 128 : (a.N % 64 == 0) ? 64 : 32;
  if (a.N % bn != 0) return L.fail(EDV_ERR_ARG, "gemm_tc: N must be a multiple of 32");
  if (a.e.act == ACT_HEAD && a.N != 32) return L.fail(EDV_ERR_ARG, "gemm_tc: head epilogue needs N == 32");
  if (a.e.map == MAP_PIXSHUF && (a.e.ps_c % bn != 0 && bn % a.e.ps_c != 0)) bn = 64;
#define EDV_TC_CASE(BN_, BK_)                                                         \
  if (bn == BN_ && bk == BK_) {                                                       \
    if (a.conv) launch_gemm_tc_inst<T, BN_, BK_, true>(L, dtype, a);                  \
    else launch_gemm_tc_inst<T, BN_, BK_, false>(L, dtype, a);                        \
    return;                                                                           \
  }
  EDV_TC_CASE(256, 64)
  EDV_TC_CASE(192, 64)
  EDV_TC_CASE(128, 64)
  EDV_TC_CASE(64, 64)
  EDV_TC_CASE(32, 64)
  EDV_TC_CASE(128, 32)
  EDV_TC_CASE(64, 32)
  EDV_TC_CASE(32, 32)
#undef EDV_TC_CASE
  L.fail(EDV_ERR_ARG, "gemm_tc: no instantiation");
}

// which compile-time epilogue instantiation (gemm_tc.cuh EF_*) matches this epilogue, or -1
inline int epi_kind(const Epi& e) {
  using namespace tc;
  if (e.rowbias || e.map == MAP_PIXSHUF || e.act == ACT_SIGMOID || e.act == ACT_GEGLU || e.act == ACT_HEAD || !e.out) return -1;
  if (e.res2 && (e.res2_f32 || !e.res1)) return -1;
  int f = 0;
  if (e.bias) f |= EF_BIAS;
  if (e.act == ACT_GELU) f |= EF_GELU;
  if (e.act == ACT_RELU) f |= EF_RELU;
  if (e.res1) f |= e.res1_f32 ? EF_RES1_F32 : EF_RES1_T;
  if (e.res2) f |= EF_RES2_T;
  if (e.out_f32) f |= EF_OUT_F32;
  if (e.out_relu) f |= EF_OUT_RELU;
  const int known[] = {EF_BIAS, EF_BIAS | EF_GELU, EF_BIAS | EF_RES1_F32 | EF_OUT_F32, EF_BIAS | EF_OUT_F32, EF_BIAS | EF_RES1_F32,
                       EF_BIAS | EF_RES1_T, EF_BIAS | EF_RES1_T | EF_OUT_RELU, EF_BIAS | EF_RES1_T | EF_RES2_T,
                       EF_BIAS | EF_RES1_T | EF_RES2_T | EF_OUT_RELU, EF_BIAS | EF_RELU, EF_OUT_RELU, 0};
  for (int k : known)
    if (k == f) return f;
  return -1;
}

// dispatch on dtype / engine.  GEGLU and HEAD epilogues exist only in the tcgen05 kernel;
// the CUDA-core path composes them from a plain GEMM plus a small kernel (see engine.cu).
inline void gemm(Launch& L, int dtype, int engine, const GemmArgs& a_in) {
  if (!L.ok()) return;
  GemmArgs a = a_in;
  a.e.kind = epi_kind(a.e);
  if (dtype == EDV_F32) return launch_gemm_simt<float>(L, a);
  if (engine == EDV_ENGINE_SIMT) {
    if (dtype == EDV_BF16) return launch_gemm_simt<bf16>(L, a);
    return launch_gemm_simt<f16>(L, a);
  }
  if (dtype == EDV_BF16) return launch_gemm_tc<bf16>(L, dtype, a);
  return launch_gemm_tc<f16>(L, dtype, a);
}

// ---- elementwise launchers --------------------------------------------------------------------
#define EDV_DISPATCH_T(dtype, ...)                      \
  do {                                                  \
    if ((dtype) == EDV_F32) { using T = float; __VA_ARGS__; } \
    else if ((dtype) == EDV_BF16) { using T = bf16; __VA_ARGS__; } \
    else { using T = f16; __VA_ARGS__; }                \
  } while (0)

inline unsigned nblk(long long n, int per) { return (unsigned)((n + per - 1) / per); }

inline void layernorm(Launch& L, int dtype, const float* x, const float* g, const float* b, void* y, long long Mout,
                      int D, float eps, int grp = 0, int skip = 0, const float* add = nullptr, int add_div = 1,
                      int add_mod = 1) {
  if (!L.ok()) return;
  if (D % 4 != 0 || D > 1024) return L.fail(EDV_ERR_ARG, "layernorm: D must be a multiple of 4 and <= 1024");
  const int maxv = (D + 127) / 128;
  const unsigned blocks = nblk(((Mout + 1) / 2) * 32, 256);   // one warp per two rows
  L.note(0, (double)Mout * D * (4 + dtype_size(dtype)));
  EDV_DISPATCH_T(dtype, {
    if (maxv <= 1) layernorm_kernel<T, 1><<<blocks, 256, 0, L.stream>>>(x, g, b, (T*)y, Mout, D, eps, grp, skip, add, add_div, add_mod);
    else if (maxv <= 2) layernorm_kernel<T, 2><<<blocks, 256, 0, L.stream>>>(x, g, b, (T*)y, Mout, D, eps, grp, skip, add, add_div, add_mod);
    else if (maxv <= 3) layernorm_kernel<T, 3><<<blocks, 256, 0, L.stream>>>(x, g, b, (T*)y, Mout, D, eps, grp, skip, add, add_div, add_mod);
    else if (maxv <= 4) layernorm_kernel<T, 4><<<blocks, 256, 0, L.stream>>>(x, g, b, (T*)y, Mout, D, eps, grp, skip, add, add_div, add_mod);
    else layernorm_kernel<T, 8><<<blocks, 256, 0, L.stream>>>(x, g, b, (T*)y, Mout, D, eps, grp, skip, add, add_div, add_mod);
  });
  L.check("layernorm");
}

constexpr int GN_MAX_SPLIT = 64;   // groupnorm stats scratch: F*32*(1+GN_MAX_SPLIT) float2

inline void groupnorm(Launch& L, int dtype, const void* x, const float* g, const float* b, void* y, float2* stats,
                      int F, int hw, int C, float eps) {
  if (!L.ok()) return;
  if (C % 32 != 0 || C % 8 != 0) return L.fail(EDV_ERR_ARG, "groupnorm: C must be a multiple of 32");
  if (dtype == EDV_F32 || C / 8 > 256) {
    // fp32 path: two-pass (mean, then centred variance) per (frame, group)
    EDV_DISPATCH_T(dtype, {
      L.note(0, (double)F * hw * C * sizeof(T));
      groupnorm_stats_kernel<T><<<dim3(32, F), 256, 0, L.stream>>>((const T*)x, stats, hw, C, eps);
    });
    L.check("groupnorm_stats");
  } else {
    const int per_px = 256 / (C / 8);
    int split = (hw + per_px * 8 - 1) / (per_px * 8);     // ~8 pixels per thread
    if (split < 1) split = 1;
    if (split > GN_MAX_SPLIT) split = GN_MAX_SPLIT;
    float2* part = stats + (size_t)F * 32;                 // caller provides F*32*(1+GN_MAX_SPLIT) entries
    EDV_DISPATCH_T(dtype, {
      L.note(0, (double)F * hw * C * sizeof(T));
      groupnorm_partial_kernel<T><<<dim3(F, split), 256, 0, L.stream>>>((const T*)x, part, hw, C);
    });
    L.check("groupnorm_stats");
    groupnorm_finalize_kernel<<<nblk((long long)F * 32, 256), 256, 0, L.stream>>>(part, stats, F * 32, split, (float)hw * (C / 32), eps);
    L.check("groupnorm_finalize");
  }
  EDV_DISPATCH_T(dtype, {
    long long total8 = (long long)F * hw * C / 8;
    L.note(0, 2.0 * F * hw * C * sizeof(T));
    groupnorm_apply_kernel<T><<<nblk(total8, 256), 256, 0, L.stream>>>((const T*)x, stats, g, b, (T*)y, total8, hw, C);
  });
  L.check("groupnorm_apply");
}

inline void upsample(Launch& L, int dtype, const void* x, void* y, int F, int h, int w, int oh, int ow, int C) {
  if (!L.ok()) return;
  if (C % 8 != 0) return L.fail(EDV_ERR_ARG, "upsample: C must be a multiple of 8");
  const long long rows = (long long)F * oh, per_row = (long long)ow * (C / 8);
  if (rows > 0x7fffffffLL || per_row > 65535LL * 256) return L.fail(EDV_ERR_ARG, "upsample: map too large");
  L.note(0, ((double)F * h * w + (double)F * oh * ow) * C * dtype_size(dtype));
  const dim3 grid((unsigned)rows, (unsigned)((per_row + 255) / 256));
  EDV_DISPATCH_T(dtype, { upsample_nhwc_kernel<T><<<grid, 256, 0, L.stream>>>((const T*)x, (T*)y, F, h, w, oh, ow, C); });
  L.check("upsample");
}

inline void resize_f32(Launch& L, const float* x, float* y, int F, int h, int w, int oh, int ow, int sigmoid = 0) {
  if (!L.ok()) return;
  long long total = (long long)F * oh * ow;
  L.note(0, ((double)F * h * w + (double)F * oh * ow) * 4);
  resize_f32_kernel<<<nblk(total, 256), 256, 0, L.stream>>>(x, y, F, h, w, oh, ow, sigmoid);
  L.check("resize_f32");
}

inline void attention(Launch& L, int dtype, int engine, const void* qkv, void* out, int F, int S, int heads) {
  if (!L.ok()) return;
  // QK^T + PV: 4*S*S*64 per (frame, head); q,k,v read once, o written once
  L.note(4.0 * F * heads * (double)S * S * 64, 4.0 * F * S * heads * 64 * dtype_size(dtype));
  if (dtype != EDV_F32 && engine == EDV_ENGINE_TC) {
    if (dtype == EDV_BF16) tc::launch_attention_tc<bf16>(L, dtype, qkv, out, F, S, heads, &make_tmap);
    else tc::launch_attention_tc<f16>(L, dtype, qkv, out, F, S, heads, &make_tmap);
    return;
  }
  dim3 grid((S + 127) / 128, heads, F);
  EDV_DISPATCH_T(dtype, { spatial_attention_simt_kernel<T><<<grid, 128, 0, L.stream>>>((const T*)qkv, (T*)out, S, heads); });
  L.check("spatial_attention_simt");
}

// fused upsample -> conv3x3 (Cin -> 32) -> ReLU -> 1x1 -> ReLU|sigmoid  (head_fused.cuh)
template <typename T, int CIN>
void launch_head_fused(Launch& L, const void* x, const void* w, const float* bias, const float* head_w, float* out, int F,
                       int H1, int W1, int OH, int OW, float sig_sign) {
  using namespace tc;
  auto kern = head_fused_kernel<T, CIN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hf_smem_bytes<CIN>());
    attr_done = true;
  }
  const int tiles_x = (OW + HF_TW - 1) / HF_TW, tiles_y = (OH + HF_TH - 1) / HF_TH;
  const long long total = (long long)F * tiles_x * tiles_y;
  if (total > 0x7fffffffLL) return L.fail(EDV_ERR_ARG, "head_fused: too many tiles");
  const int grid = (int)std::min<long long>(total, num_sms());
  // algorithmic bytes: the low-resolution map once + one float per output pixel
  L.note(2.0 * F * OH * OW * 32.0 * (9 * CIN + 1), (double)F * H1 * W1 * CIN * sizeof(T) + (double)F * OH * OW * 4);
  kern<<<grid, HF_THREADS, hf_smem_bytes<CIN>(), L.stream>>>((const T*)x, (const T*)w, bias, head_w, out, F, H1, W1, OH, OW,
                                                           sig_sign, tiles_x, tiles_y, (int)total);
  L.check("head_fused");
}

inline bool head_fused_supported(int dtype, int Cin) { return dtype != EDV_F32 && (Cin == 32 || Cin == 128); }

inline void head_fused(Launch& L, int dtype, const void* x, const void* w, const float* bias, const float* head_w,
                       float* out, int F, int H1, int W1, int OH, int OW, int Cin, float sig_sign) {
  if (!L.ok()) return;
  if (!head_fused_supported(dtype, Cin)) return L.fail(EDV_ERR_ARG, "head_fused: 16-bit dtype and Cin in {32,128} only");
#define EDV_HF_CASE(T_, C_) launch_head_fused<T_, C_>(L, x, w, bias, head_w, out, F, H1, W1, OH, OW, sig_sign)
  if (dtype == EDV_BF16) { if (Cin == 32) EDV_HF_CASE(bf16, 32); else EDV_HF_CASE(bf16, 128); }
  else { if (Cin == 32) EDV_HF_CASE(f16, 32); else EDV_HF_CASE(f16, 128); }
#undef EDV_HF_CASE
}

template <typename T, int HD>
void launch_temporal(Launch& L, const void* qkv, void* out, int B, int Tn, int hw, int C, const float* rope) {
  const int heads = 8;
  // shared memory per (position, head): K and V in fp32, q/o in T
  const size_t per = (size_t)Tn * HD * (8 + sizeof(T));
  int hg = heads;
  while (hg > 1 && hg * per > 64 * 1024) hg >>= 1;
  int pb = (int)((64 * 1024) / (hg * per));
  if (pb < 1) pb = 1;
  if (pb > 4) pb = 4;
  while (pb > 1 && 32 * hg * pb > ta_max_threads<HD>()) --pb;
  while (hg > 1 && 32 * hg * pb > ta_max_threads<HD>()) hg >>= 1;
  if (pb > hw) pb = hw;
  const size_t smem = (size_t)pb * hg * per + (size_t)Tn * 16 + 16;
  auto kern = temporal_attention_kernel<T, HD>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  dim3 grid((hw + pb - 1) / pb, B, heads / hg);
  L.note(4.0 * B * hw * heads * (double)Tn * Tn * HD, 4.0 * B * Tn * hw * C * sizeof(T));
  kern<<<grid, 32 * hg * pb, smem, L.stream>>>((const T*)qkv, (T*)out, Tn, hw, C, pb, hg, (const float2*)rope);
  L.check("temporal_attention");
}

template <typename T, int HD, int PB>
void launch_temporal_mma_pb(Launch& L, const void* qkv, void* out, int B, int Tn, int hw, const float* rope) {
  auto kern = tmma::temporal_attention_mma_kernel<T, HD, PB>;
  const size_t smem = PB * tmma::ta_smem_per_pos<HD>();
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  dim3 grid((hw + PB - 1) / PB, B);
  const int C = 8 * HD;
  L.note(4.0 * B * hw * 8 * (double)Tn * Tn * HD, 4.0 * B * Tn * hw * C * sizeof(T));
  kern<<<grid, tmma::TA_WARPS * 32, smem, L.stream>>>((const T*)qkv, (T*)out, Tn, hw, (const float2*)rope);
  L.check("temporal_attention");
}

// 16-bit paths: warp-level tensor-core kernel (attention_temporal_mma.cuh); PB positions per CTA so that the
// staging fits ~96 KB (two CTAs per SM) and every warp gets at least one (position, head) task
template <typename T, int HD>
void launch_temporal_mma(Launch& L, const void* qkv, void* out, int B, int Tn, int hw, const float* rope) {
  constexpr size_t per = tmma::ta_smem_per_pos<HD>();
  if constexpr (4 * per <= 96 * 1024) {
    if (hw >= 4) return launch_temporal_mma_pb<T, HD, 4>(L, qkv, out, B, Tn, hw, rope);
  }
  if constexpr (2 * per <= 96 * 1024) {
    if (hw >= 2) return launch_temporal_mma_pb<T, HD, 2>(L, qkv, out, B, Tn, hw, rope);
  }
  launch_temporal_mma_pb<T, HD, 1>(L, qkv, out, B, Tn, hw, rope);
}

inline bool temporal_simt_forced() {
  static const bool v = [] { const char* e = getenv("EDV_TEMPORAL_SIMT"); return e && e[0] == '1'; }();
  return v;
}

inline void temporal_attention(Launch& L, int dtype, const void* qkv, void* out, int B, int Tn, int hw, int C,
                               const float* rope = nullptr) {
  if (!L.ok()) return;
  if (Tn > 32 || Tn < 1) return L.fail(EDV_ERR_ARG, "temporal attention: T must be in [1,32] (motion_module.py:185-197)");
  if (C % 8 != 0) return L.fail(EDV_ERR_ARG, "temporal attention: C must be a multiple of 8");
  const int hd = C / 8;
  if (dtype != EDV_F32 && !temporal_simt_forced()) {
#define EDV_TM_CASE(HD_)                                                                                   \
    if (hd == HD_) {                                                                                       \
      if (dtype == EDV_BF16) launch_temporal_mma<bf16, HD_>(L, qkv, out, B, Tn, hw, rope);                 \
      else launch_temporal_mma<f16, HD_>(L, qkv, out, B, Tn, hw, rope);                                     \
      return;                                                                                              \
    }
    EDV_TM_CASE(8)
    EDV_TM_CASE(24)
    EDV_TM_CASE(32)
    EDV_TM_CASE(48)
    EDV_TM_CASE(128)
#undef EDV_TM_CASE
  }
#define EDV_TA_CASE(HD_)                                                                   \
  if (hd == HD_) {                                                                         \
    EDV_DISPATCH_T(dtype, { launch_temporal<T, HD_>(L, qkv, out, B, Tn, hw, C, rope); });        \
    return;                                                                                \
  }
  EDV_TA_CASE(8)
  EDV_TA_CASE(24)
  EDV_TA_CASE(32)
  EDV_TA_CASE(48)
  EDV_TA_CASE(128)
#undef EDV_TA_CASE
  L.fail(EDV_ERR_ARG, "temporal attention: unsupported head dim (supported 8,24,32,48,128)");
}

}  // namespace edv
