// Host launchers of the tcgen05 GEMM family (gemm_tc.cuh, gemm_tc2.cuh, conv_halo.cuh).  Included by
// the gemm_tc_*.cu translation units, each of which explicitly instantiates one (dtype, lin|conv) slice.
#pragma once
#include "ops.h"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "gemm_bres.cuh"
#include "gemm_ln.cuh"
#include "conv_halo.cuh"

namespace edv {

// EDV_GEMM_TMA_OUT=0 falls back to the row-segment epilogue everywhere (A/B measurements)
inline bool gemm_tma_out_enabled() {
  static const bool v = [] { const char* e = getenv("EDV_GEMM_TMA_OUT"); return !(e && e[0] == '0'); }();
  return v;
}
// the TMA-store epilogue handles 16-bit row-major outputs with bias / GELU / ReLU only
inline bool gemm_tma_out_ok(const GemmArgs& a) {
  using namespace tc;
  const int k = a.e.kind;
  if (!gemm_tma_out_enabled() || a.conv || k < 0) return false;
  if (k != 0 && k != EF_BIAS && k != (EF_BIAS | EF_GELU) && k != (EF_BIAS | EF_RELU)) return false;
  if (a.e.map != MAP_LINEAR || a.e.out_f32 || !a.e.out) return false;
  if ((a.e.ldo * 2) % 16 != 0 || ((uintptr_t)a.e.out & 15) != 0) return false;
  return true;
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
  // TMA-store epilogue: output tensor map over C [M rows, N cols] (row pitch ldo), box 16 columns x 128 rows, 32B swizzle
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof tmC);
  GemmArgs a2 = a;
  if (!CONV && BN % 64 == 0 && gemm_tma_out_ok(a)) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.e.ldo * 2};
    uint32_t box[2] = {16u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmC, dtype, a.e.out, 2, dims, str, box, 32)) return;
    a2.e.kind |= tc::EF_TMA_OUT;
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
  edv::launch_k(kern, dim3(grid), dim3(GT_THREADS), smem, L.stream, tmA, tmB, tmC, a2.e, a.M, a.N, a.K, stages, ct, n_tiles, (int)total);
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
  edv::launch_k(kern, dim3(grid), dim3(CH_THREADS), ch_smem_bytes<64, BN>(), L.stream, (const T*)a.A, (const T*)a.W, a.e, a.F, a.H, a.Wd, tiles_x,
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


inline int gemm_tc_pick_bk(const GemmArgs& a) { return (a.conv ? (a.C % 64 != 0) : (a.K % 64 != 0)) ? 32 : 64; }

inline int gemm_tc_pick_bn(const GemmArgs& a, int bk) {
  int bn;
  if (a.e.act == ACT_GEGLU) bn = 128;
  else if (a.e.act == ACT_HEAD) bn = 32;
  else if (bk == 64 && a.N % 256 == 0) bn = 256;   // wide tiles halve the shared-memory traffic per FLOP
  else if (bk == 64 && a.N % 192 == 0) bn = 192;
  else bn = (a.N % 128 == 0) ? 128 : (a.N % 64 == 0) ? 64 : 32;
  if (a.e.map == MAP_PIXSHUF && (a.e.ps_c % bn != 0 && bn % a.e.ps_c != 0)) bn = 64;
  return bn;
}

template <typename T, bool CONV> void launch_gemm_tc_any(Launch& L, int dtype, const GemmArgs& a) {
  const int bk = gemm_tc_pick_bk(a);
  if ((a.conv ? a.C : a.K) % bk != 0) return L.fail(EDV_ERR_ARG, "gemm_tc: K (or conv C) must be a multiple of 32");
  if (a.conv && a.stride != 1) return L.fail(EDV_ERR_ARG, "gemm_tc: conv stride must be 1");
  const int bn = gemm_tc_pick_bn(a, bk);
  if (a.N % bn != 0) return L.fail(EDV_ERR_ARG, "gemm_tc: N must be a multiple of 32");
  if (a.e.act == ACT_HEAD && a.N != 32) return L.fail(EDV_ERR_ARG, "gemm_tc: head epilogue needs N == 32");
#define EDV_TC_CASE(BN_, BK_) \
  if (bn == BN_ && bk == BK_) return launch_gemm_tc_inst<T, BN_, BK_, CONV>(L, dtype, a);
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

#ifdef EDV_GEMM_TU_LIN
// B-resident GEMM (gemm_bres.cuh): short K, 16-bit output through the TMA-store epilogue
inline bool gemm_bres_enabled() {
  static const bool v = [] { const char* e = getenv("EDV_GEMM_BRES"); return !(e && e[0] == '0'); }();
  return v;
}
constexpr size_t GB_SMEM_MAX = 232448 - 1024;   // 227 KB opt-in limit minus the alignment slack

template <typename T, int BN> bool launch_gemm_bres(Launch& L, int dtype, const GemmArgs& a) {
  using namespace tc;
  const int kblocks = a.K / 64;
  const size_t b_bytes = (size_t)kblocks * BN * 128;
  const size_t fixed = b_bytes + 4 * GT_OUT_SUB_BYTES + 4 * 64 * 4 + 512;
  if (fixed + 2 * GB_A_STAGE > GB_SMEM_MAX) return false;
  int stages = (int)((GB_SMEM_MAX - fixed) / GB_A_STAGE);
  if (stages > 8) stages = 8;
  CUtensorMap tmA, tmB, tmC;
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {64u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmA, dtype, a.A, 2, dims, str, box, 128)) return true;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.K * 2};
    uint32_t box[2] = {64u, (uint32_t)BN};
    if (!make_tmap(L, &tmB, dtype, a.W, 2, dims, str, box, 128)) return true;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.e.ldo * 2};
    uint32_t box[2] = {16u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmC, dtype, a.e.out, 2, dims, str, box, 32)) return true;
  }
  auto kern = gemm_bres_kernel<T, BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    attr_done = true;
  }
  const int n_tiles = a.N / BN;
  const int m_tiles = (a.M + GT_BM - 1) / GT_BM;
  int grid = num_sms();
  if (grid < n_tiles) return false;
  if ((long long)m_tiles * n_tiles < grid) grid = std::max(n_tiles, m_tiles * n_tiles);
  const size_t smem = fixed + (size_t)stages * GB_A_STAGE + 1024;
  GemmArgs a2 = a;
  a2.e.kind |= EF_TMA_OUT;
  note_gemm(L, a, 2);
  edv::launch_k(kern, dim3(grid), dim3(GB_THREADS), smem, L.stream, tmA, tmB, tmC, a2.e, a.M, a.K, stages, n_tiles, m_tiles);
  L.check("gemm_bres");
  return true;
}

template <typename T>
void launch_gemm_ln(Launch& L, int dtype, const void* A, const void* W, const float* bias, float* x, void* xn, const float* gamma,
                    const float* beta, float eps, int M, int K, int do_ln, long long* tim) {
  using namespace tc;
  CUtensorMap tmA, tmB, tmX, tmXn;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmA, dtype, A, 2, dims, str, box, 128)) return;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)GL_N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64u, (uint32_t)GL_HN};
    if (!make_tmap(L, &tmB, dtype, W, 2, dims, str, box, 128)) return;
  }
  {
    uint64_t dims[2] = {(uint64_t)GL_N, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)GL_N * 4};
    uint32_t box[2] = {16u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmX, EDV_F32, x, 2, dims, str, box, 64)) return;
  }
  {
    uint64_t dims[2] = {(uint64_t)GL_N, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)GL_N * 2};
    uint32_t box[2] = {16u, (uint32_t)GT_BM};
    if (!make_tmap(L, &tmXn, dtype, do_ln ? xn : (const void*)x, 2, dims, str, box, 32)) return;
  }
  auto kern = gemm_ln_kernel<T>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GL_SMEM);
    attr_done = true;
  }
  const int m_tiles = (M + GT_BM - 1) / GT_BM;
  const int grid = 2 * std::min(m_tiles, num_sms() / 2);   // clusters of two CTAs, one 128-row tile per pair at a time
  // algorithmic work: 2 M N K; bytes: A + W + the fp32 stream read and written once + xn written once
  L.note(2.0 * M * GL_N * K, ((double)M * K + (double)GL_N * K) * 2 + (double)M * GL_N * (8 + (do_ln ? 2 : 0)));
  edv::launch_k(kern, dim3(grid), dim3(GL_THREADS), GL_SMEM, L.stream, tmA, tmB, tmX, tmXn, bias, gamma, beta, eps, M, K, m_tiles, do_ln, tim);
  L.check("gemm_ln");
}

#ifdef EDV_GEMM_LN_DEFS
inline bool gemm_ln_enabled() {
  static const bool v = [] { const char* e = getenv("EDV_GEMM_LN"); return !(e && e[0] == '0'); }();
  return v;
}
bool gemm_ln_supported(int dtype, int N, int K) {
  return gemm_ln_enabled() && dtype != EDV_F32 && N == tc::GL_N && K % 64 == 0 && K >= 64;
}
void gemm_ln(Launch& L, int dtype, const void* A, const void* W, const float* bias, float* x, void* xn, const float* gamma,
             const float* beta, float eps, int M, int N, int K, int do_ln, long long* tim) {
  if (!L.ok()) return;
  if (!gemm_ln_supported(dtype, N, K)) return L.fail(EDV_ERR_ARG, "gemm_ln: needs a 16-bit dtype, N == 384 and K % 64 == 0");
  if (dtype == EDV_BF16) launch_gemm_ln<bf16>(L, dtype, A, W, bias, x, xn, gamma, beta, eps, M, K, do_ln, tim);
  else launch_gemm_ln<f16>(L, dtype, A, W, bias, x, xn, gamma, beta, eps, M, K, do_ln, tim);
}
#endif

// linear GEMMs (2-D A operand)
template <typename T> void launch_gemm_tc_lin(Launch& L, int dtype, const GemmArgs& a) {
  if (gemm_bres_enabled() && gemm_tma_out_ok(a) && a.K % 64 == 0 && a.K <= 384 && a.lda == a.K && a.M >= 2048) {
    if (a.N % 192 == 0) { if (launch_gemm_bres<T, 192>(L, dtype, a)) return; }
    else if (a.N % 128 == 0) { if (launch_gemm_bres<T, 128>(L, dtype, a)) return; }
    else if (a.N % 64 == 0) { if (launch_gemm_bres<T, 64>(L, dtype, a)) return; }
  }
  if (gemm_2sm_min_m() > 0 && a.M >= gemm_2sm_min_m() && a.K % 64 == 0 && a.lda == a.K && a.e.act != ACT_GEGLU &&
      a.e.act != ACT_HEAD) {
    if (a.N % 256 == 0) return launch_gemm_tc2<T, 256>(L, dtype, a);
    if (a.N % 192 == 0) return launch_gemm_tc2<T, 192>(L, dtype, a);
  }
  launch_gemm_tc_any<T, false>(L, dtype, a);
}
#endif

#ifdef EDV_GEMM_TU_CONV
// 3x3 convolutions: halo-reuse kernel for the 64-channel maps, TMA implicit GEMM otherwise
template <typename T> void launch_gemm_tc_conv(Launch& L, int dtype, const GemmArgs& a) {
  if (conv_halo_ok(a)) {
    if (a.N == 64) launch_conv_halo<T, 64>(L, a);
    else launch_conv_halo<T, 32>(L, a);
    return;
  }
  launch_gemm_tc_any<T, true>(L, dtype, a);
}
#endif

}  // namespace edv
