// pa_attn.cuh — softmax(Q K^T * scale) V core on tcgen05 + TMEM, head_dim 64, keys per unit <= 256.
//
// Work item = (group g, head h, pair of 128-row query tiles).  All keys of the unit fit one S tile, so the
// row softmax is exact and single pass (no online rescaling).  Persistent CTA, 384 threads:
//   warp 0     TMA producer: Q tile(s), K, V of the next item into a 2-deep smem ring (128B swizzle)
//   warp 1     MMA issuer:   S_s = Q_s K^T  (SS, K-major both)  ->  TMEM slot s;   O_s = P_s V  (A = P from TMEM,
//                            B = V MN-major straight from its natural [key][d] layout)
//   warp 2     TMEM allocator (512 columns = 2 slots x 256)
//   warps 4-7  softmax / epilogue warpgroup of slot 0   (thread <-> query row, TMEM lane)
//   warps 8-11 softmax / epilogue warpgroup of slot 1
// TMEM slot layout (columns): S fp32 [0, kp)   P fp16x2 [0, kp/2) (written in place behind the S reads)
//                             O fp32 [192, 256) (written by the PV MMAs only after the softmax has drained S)
#pragma once
#include "pa_ptx.cuh"

namespace pa {

struct AttnParams {
  int G, H;             // groups (batch entries), heads
  int n_q, n_k;         // query / key rows per group
  int kp;               // keys padded to a multiple of 16 (<= 256) = S tile width = K/V box rows
  int q_tiles;          // ceil(n_q / 128)
  int pairs;            // ceil(q_tiles / 2)
  int items;            // G * H * pairs
  int q_col0, k_col0, v_col0;   // element column of head 0 inside the Q / KV tensor maps
  void* O;              // fp16 output [G][n_q][ldo]
  long long ldo, o_group;
  int o_col0;
  float scale_log2e;    // softmax scale * log2(e)
  uint32_t idesc_s, idesc_o;
};

constexpr int ATTN_THREADS = 384;
constexpr int ATTN_HD = 64;
constexpr int ATTN_SLOT_COLS = 256;
constexpr int ATTN_O_COL = 192;
constexpr int ATTN_Q_BYTES = 128 * ATTN_HD * 2;   // 16 KB per query tile

__host__ __device__ inline int attn_item_bytes(int kp) { return 2 * ATTN_Q_BYTES + 2 * kp * ATTN_HD * 2; }
__host__ __device__ inline int attn_smem_bytes(int kp) { return 2 * attn_item_bytes(kp) + 256 + 1024; }

__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_core_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int item_bytes = attn_item_bytes(p.kp);
  const int kv_bytes = p.kp * ATTN_HD * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * item_bytes);
  uint64_t* item_full = bars;         // [2]  TMA -> MMA
  uint64_t* item_empty = bars + 2;    // [2]  MMA -> TMA
  uint64_t* s_full = bars + 4;        // [2]  MMA -> softmax(slot)
  uint64_t* p_full = bars + 6;        // [2]  softmax(slot) -> MMA
  uint64_t* o_full = bars + 8;        // [2]  MMA -> epilogue(slot)
  uint64_t* slot_empty = bars + 10;   // [2]  epilogue(slot) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&item_full[i], 1);
      mbar_init(&item_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&slot_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int b = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int pr = item % p.pairs;
        const int gh = item / p.pairs;
        const int h = gh % p.H, g = gh / p.H;
        const bool two = (2 * pr + 1) < p.q_tiles;
        uint8_t* buf = smem + b * item_bytes;
        mbar_wait(&item_empty[b], ph ^ 1);
        mbar_expect_tx(&item_full[b], (two ? 2 : 1) * ATTN_Q_BYTES + 2 * kv_bytes);
        tma_load_3d(buf, &tmQ, p.q_col0 + h * ATTN_HD, (2 * pr) * 128, g, &item_full[b]);
        if (two) tma_load_3d(buf + ATTN_Q_BYTES, &tmQ, p.q_col0 + h * ATTN_HD, (2 * pr + 1) * 128, g, &item_full[b]);
        tma_load_3d(buf + 2 * ATTN_Q_BYTES, &tmKV, p.k_col0 + h * ATTN_HD, 0, g, &item_full[b]);
        tma_load_3d(buf + 2 * ATTN_Q_BYTES + kv_bytes, &tmKV, p.v_col0 + h * ATTN_HD, 0, g, &item_full[b]);
        if (++b == 2) { b = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int b = 0;
      uint32_t ph = 0;
      uint32_t slot_ph[2] = {0, 0};
      const int ksteps_o = p.kp / 16;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int pr = item % p.pairs;
        const int nslots = ((2 * pr + 1) < p.q_tiles) ? 2 : 1;
        const uint32_t buf = smem_u32(smem + b * item_bytes);
        mbar_wait(&item_full[b], ph);
        tc_fence_after();
        const uint64_t kdesc = make_sdesc(buf + 2 * ATTN_Q_BYTES, 16, 1024, PA_SWZ_128B);
        for (int s = 0; s < nslots; ++s) {
          mbar_wait(&slot_empty[s], slot_ph[s] ^ 1);
          tc_fence_after();
          const uint64_t qdesc = make_sdesc(buf + s * ATTN_Q_BYTES, 16, 1024, PA_SWZ_128B);
          const uint32_t d = tmem_base + s * ATTN_SLOT_COLS;
#pragma unroll
          for (int k = 0; k < ATTN_HD / 16; ++k) umma_ss(d, qdesc + 2 * k, kdesc + 2 * k, p.idesc_s, k != 0);
          umma_commit(&s_full[s]);
        }
        for (int s = 0; s < nslots; ++s) {
          mbar_wait(&p_full[s], slot_ph[s]);
          tc_fence_after();
          const uint32_t slot = tmem_base + s * ATTN_SLOT_COLS;
          // V is [key][d] with d contiguous: MN-major B operand, 8-key groups are 1024 B apart (SBO)
          const uint64_t vdesc = make_sdesc(buf + 2 * ATTN_Q_BYTES + kv_bytes, 1024, 1024, PA_SWZ_128B);
          for (int k = 0; k < ksteps_o; ++k) {
            // P: 16 fp16 keys = 8 TMEM columns per step; V: 16 keys = 2048 B = 128 x 16 B per step
            umma_ts(slot + ATTN_O_COL, slot + 8 * k, vdesc + 128 * k, p.idesc_o, k != 0);
          }
          umma_commit(&o_full[s]);
          slot_ph[s] ^= 1;
        }
        umma_commit(&item_empty[b]);
        if (++b == 2) { b = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue warpgroups =====================
    const int slot = (warp - 4) >> 2;
    const int q = warp & 3;
    const uint32_t t_slot = tmem_base + slot * ATTN_SLOT_COLS + ((uint32_t)(q * 32) << 16);
    const int nchunks = (p.kp + 31) / 32;
    const float sl2 = p.scale_log2e;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int pr = item % p.pairs;
      const int gh = item / p.pairs;
      const int h = gh % p.H, g = gh / p.H;
      const int qt = 2 * pr + slot;
      if (qt >= p.q_tiles) continue;          // this slot is idle for the item (uniform over the warpgroup)
      const int row = qt * 128 + q * 32 + lane;
      const bool warp_active = (qt * 128 + q * 32) < p.n_q;   // some row of this warp is a real query
      mbar_wait(&s_full[slot], ph);
      tc_fence_after();
      float sum = 1.f;
      if (warp_active) {
        // ---- pass 1: row max over the real keys
        float mx = -INFINITY;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t v[32];
          tmem_ld32(t_slot + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = (c * 32 + i < p.n_k) ? __uint_as_float(v[i]) : -INFINITY;
            mx = fmaxf(mx, x);
          }
        }
        const float mxs = mx * sl2;
        // ---- pass 2: p = exp2(s*scale*log2e - max), row sum, fp16 P written in place
        sum = 0.f;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t v[32];
          tmem_ld32(t_slot + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float x0 = (c * 32 + i < p.n_k) ? __uint_as_float(v[i]) : -INFINITY;
            const float x1 = (c * 32 + i + 1 < p.n_k) ? __uint_as_float(v[i + 1]) : -INFINITY;
            const float e0 = ex2f(fmaf(x0, sl2, -mxs));
            const float e1 = ex2f(fmaf(x1, sl2, -mxs));
            sum += e0 + e1;
            pk[i >> 1] = pack_h2(e0, e1);
          }
          tmem_st16(t_slot + c * 16, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[slot]);

      // ---- epilogue: O / rowsum -> fp16 -> global
      mbar_wait(&o_full[slot], ph);
      tc_fence_after();
      if (warp_active) {
        const float inv = 1.f / sum;
        uint32_t v0[32], v1[32];
        tmem_ld32(t_slot + ATTN_O_COL, v0);
        tmem_ld32(t_slot + ATTN_O_COL + 32, v1);
        tmem_ld_wait();
        if (row < p.n_q) {
          uint16_t* dst = reinterpret_cast<uint16_t*>(p.O) + (long long)g * p.o_group + (long long)row * p.ldo +
                          p.o_col0 + h * ATTN_HD;
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            d4[i] = make_uint4(pack_h2(__uint_as_float(v0[8 * i + 0]) * inv, __uint_as_float(v0[8 * i + 1]) * inv),
                               pack_h2(__uint_as_float(v0[8 * i + 2]) * inv, __uint_as_float(v0[8 * i + 3]) * inv),
                               pack_h2(__uint_as_float(v0[8 * i + 4]) * inv, __uint_as_float(v0[8 * i + 5]) * inv),
                               pack_h2(__uint_as_float(v0[8 * i + 6]) * inv, __uint_as_float(v0[8 * i + 7]) * inv));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            d4[4 + i] =
                make_uint4(pack_h2(__uint_as_float(v1[8 * i + 0]) * inv, __uint_as_float(v1[8 * i + 1]) * inv),
                           pack_h2(__uint_as_float(v1[8 * i + 2]) * inv, __uint_as_float(v1[8 * i + 3]) * inv),
                           pack_h2(__uint_as_float(v1[8 * i + 4]) * inv, __uint_as_float(v1[8 * i + 5]) * inv),
                           pack_h2(__uint_as_float(v1[8 * i + 6]) * inv, __uint_as_float(v1[8 * i + 7]) * inv));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_empty[slot]);
      ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace pa
