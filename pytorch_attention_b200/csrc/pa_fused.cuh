// pa_fused.cuh — ONE launch for the whole ViT attention forward (ViT.py:79-89):
//     phase 1  qkv = x Wqkv^T (+b)            gemm_run  (CTA pairs, cta_group::2)
//     phase 2  softmax(q k^T scale) v          attn_run  (per (image, head) items)
//     phase 3  y = O Wproj^T + b               gemm_run
// Every persistent CTA walks its share of phase 1, then of phase 2, then of phase 3.  There is no grid-wide barrier:
// phases are chained by dependency counters in global memory at tile granularity
//     ctr_qkv[128-row tile of the qkv buffer]  += 1 per epilogue warp set of every stored GEMM tile -> an attention item waits for the tiles covering its image
//     ctr_attn[image]                          += 1 per stored (head, query tile) -> a proj tile waits for the images covering its rows
// so a CTA that runs out of phase-1 tiles starts attention on the first images while others finish the last GEMM tiles:
// no partially filled last wave, no per-kernel prologue/epilogue tail, no launch gaps, intermediates stay in L2.
// Producers publish with  bulk-store completion -> __threadfence -> atomicAdd ; consumers spin with ld.acquire.gpu and a
// proxy fence before their TMA loads.  Work is ordered by image in all three phases, and a phase never waits on a later
// one, so the scheme cannot deadlock as long as all CTAs are co-resident (grid = #SMs, 1 CTA/SM).
#pragma once
#include "pa_attn.cuh"
#include "pa_gemm.cuh"

namespace pa {

struct VitFusedParams {
  GemmParams g1;     // qkv projection
  AttnParams at;     // attention core
  GemmParams g2;     // output projection
};

constexpr int FUSED_ESETS = 4;                   // GEMM phases drain accumulators with warps 4-19 (all softmax warps of the attention phase)
// GEMM phases use 256 x BN pair tiles, BN in {192, 256} chosen per phase by the host's tile cost model (e.g. ViT-B: qkv 256,
// proj 192; ViT-L: qkv 192, proj 256 -- 1024 output columns are 5.33 tiles of 192)
__host__ __device__ constexpr int fused_stages(int bn) { return bn == 256 ? 6 : 5; }

template <int BN1, int BN2>
__host__ __device__ inline int vit_fused_data_bytes(int kb) {
  const int g1 = GemmCfg<BN1, fused_stages(BN1), true>::BAR_OFFSET;
  const int g2 = GemmCfg<BN2, fused_stages(BN2), true>::BAR_OFFSET;
  const int at = 2 * 256 * 128 + 4 * kb * 128 + attn_ostage_bytes(64, true);
  int m = g1 > g2 ? g1 : g2;
  return m > at ? m : at;
}
template <int BN1, int BN2>
__host__ __device__ inline int vit_fused_smem_bytes(int kb) {
  return vit_fused_data_bytes<BN1, BN2>(kb) + (2 * fused_stages(BN1) + 4 + 16 + 2 * fused_stages(BN2) + 4) * 8 + 16 + 1024;
}

template <int BN1, int BN2>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
vit_fused_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmD1, const __grid_constant__ CUtensorMap tmQ,
                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmD2,
                 const VitFusedParams P) {
  constexpr int FUSED_BN1 = BN1, FUSED_ST1 = fused_stages(BN1), FUSED_BN2 = BN2, FUSED_ST2 = fused_stages(BN2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars1 = reinterpret_cast<uint64_t*>(smem + vit_fused_data_bytes<BN1, BN2>(P.at.kb));
  uint64_t* barsA = bars1 + (2 * FUSED_ST1 + 4);
  uint64_t* bars2 = barsA + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars2 + (2 * FUSED_ST2 + 4));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmD1);
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmD2);
  }
  if (warp == 1 && lane == 0) {
    gemm_init_barriers<FUSED_ST1, 2, true, FUSED_ESETS>(bars1);
    attn_init_barriers(barsA);
    gemm_init_barriers<FUSED_ST2, 2, true, FUSED_ESETS>(bars2);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto stamp = [&](int slot) {      // debug: per-CTA phase boundaries on the global timer (ns), comparable across SMs
    if (P.g1.trace != nullptr && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.g1.trace[1536 + blockIdx.x * 4 + slot] = (long long)t;
    }
  };
  stamp(0);

  // ---- phase 1: qkv projection
  gemm_run<FUSED_BN1, FUSED_ST1, 2, true, true, FUSED_ESETS>(tmA1, tmB1, tmD1, P.g1, smem, bars1, tmem_base);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the pair's MMAs read both CTAs' shared memory: nobody reuses it before both are done
  tc_fence_after();
  stamp(1);

  // ---- phase 2: attention items (per CTA; the pairing is irrelevant here)
  attn_run<64, false>(tmQ, tmK, tmV, tmO, P.at, smem, barsA, tmem_base);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  stamp(2);

  // ---- phase 3: output projection
  gemm_run<FUSED_BN2, FUSED_ST2, 2, true, true, FUSED_ESETS>(tmA2, tmB2, tmD2, P.g2, smem, bars2, tmem_base);

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  stamp(3);
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

}  // namespace pa
