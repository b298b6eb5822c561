// 2-SM (cta_group::2) variant of the persistent tcgen05 GEMM for the large token GEMMs of the
// encoder (qkv, attn.proj, fc1, fc2):  C[M,N] = A[M,K] * W[N,K]^T.
//
// Two CTAs of a cluster (an SM pair) compute one 256 x BN tile.  Each CTA stages ITS 128 rows of A
// and HALF of the W tile (BN/2 rows); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), for
// which each SM's tensor core reads its own A rows and both halves of W.  Versus gemm_tc.cuh this
// halves the W bytes every SM pulls through L2 and reads from shared memory per MMA -- ncu and the
// in-kernel timers show the K = 384 GEMMs bound by exactly those two (L2 -> SM at ~12 TB/s,
// SS-MMA at 146 instead of 96 cycles for N = 192) once the epilogue is hidden.
//
// Protocol (per CTA: warp 0 TMA producer, warp 1 TMEM allocator (+ MMA issuer in the leader),
// warps 4..19 the same four epilogue warpgroups as gemm_tc.cuh):
//   full[s]    leader's barrier: both CTAs' TMA loads complete_tx on it (peer bit cleared)
//   empty[s]   one per CTA: tcgen05.commit multicast to both CTAs frees the stage in both
//   tfull[b]   one per CTA: commit multicast hands accumulator b to both epilogues
//   tempty[b]  leader's barrier: one arrive per epilogue warp of BOTH CTAs (remote arrive from the peer)
#pragma once
#include "gemm_tc.cuh"

namespace tc {

constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address (-> even CTA)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int BN> constexpr int gt2_stage_bytes() { return (GT_BM + BN / 2) * 64 * 2; }

template <typename T, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Epi e, int M, int N,
                    int K, int stages, int n_tiles, int total_tiles) {
  constexpr int BK = 64;
  constexpr uint32_t A_BYTES = GT_BM * BK * 2;
  constexpr uint32_t BH_BYTES = (BN / 2) * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + BH_BYTES;
  constexpr uint32_t TMEM_COLS = gt_tmem_cols<BN>();
  extern __shared__ __align__(1024) unsigned char smem2_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem2_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint32_t* stg_base = tmem_slot + 4;
  float* bias_base = reinterpret_cast<float*>(stg_base + 16 * GT_STG_WORDS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int kblocks = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 16);      // the 8 epilogue warps of buffer b x 2 CTAs
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  fence_before_sync();
  cluster_sync_all();                      // barriers of both CTAs initialised before any remote signal
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kc = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int tile_n = tile % n_tiles, tile_m2 = tile / n_tiles;
        for (int kb = 0; kb < kblocks; ++kb, ++kc) {
          const int s = kc % stages;
          mbar_wait(&empty_bar[s], ((kc / stages) & 1) ^ 1);
          unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
          if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
          tma_load_2d_2sm(sa, &tmA, &full_bar[s], kb * BK, tile_m2 * 256 + (int)rank * GT_BM);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, tile_n * BN + (int)rank * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // warp-uniform issue loop in the leader CTA (elect_one_sync, tc_common.cuh)
      constexpr uint32_t idesc = make_idesc<T>(256, BN, 0);
      const uint32_t issuer = elect_one_sync();
      uint32_t kc = 0, it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
        const uint32_t b = it & 1;
        mbar_wait(&tempty_bar[b], ((it >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t acc = tmem_base + b * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++kc) {
          const int s = kc % stages;
          mbar_wait(&full_bar[s], (kc / stages) & 1);
          fence_after_sync();
          const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa, 1024, 16, SWZ_128B);
          const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 1024, 16, SWZ_128B);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              mma_ss_2sm(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            mma_commit_2sm(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (issuer) mma_commit_2sm(&tfull_bar[b]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const uint32_t wg = (warp - 4) >> 2;
    const uint32_t g = wg & 1;
    const int half = wg >> 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint32_t it = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
      if ((it & 1) != g) continue;
      const int tile_n = tile % n_tiles, tile_m2 = tile / n_tiles;
      const long long m = (long long)tile_m2 * 256 + (long long)rank * GT_BM + r;
      const bool valid = m < M;
      const long long orow = epi_row(e, m);
      float* bias_s = nullptr;
      if (e.bias && e.act != ACT_GEGLU && e.act != ACT_HEAD) {
        constexpr int CH0 = (BN >= 64) ? BN / 2 : BN;
        bias_s = bias_base + (warp - 4) * 128;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < (CH0 + 31) / 32; ++k)
          if (lane + 32 * k < CH0) bias_s[lane + 32 * k] = __ldg(e.bias + tile_n * BN + half * CH0 + lane + 32 * k);
        __syncwarp();
      }
      mbar_wait(&tfull_bar[g], (it >> 1) & 1);
      fence_after_sync();
      gt_epilogue<T, BN>(e, tmem_base + ((uint32_t)(q * 32) << 16) + g * BN, valid, m, orow, tile_n * BN, tile_n,
                         stg_base + (warp - 4) * GT_STG_WORDS, lane, half, bias_s);
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[g]);
    }
  }
  fence_before_sync();
  cluster_sync_all();                      // nobody leaves (or frees TMEM) while the pair still uses its memory
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
