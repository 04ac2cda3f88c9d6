// pa_attn_wide.cuh — softmax(Q K^T * scale) V for the head dims the 32/64-wide core (pa_attn.cuh) does not take: every multiple
// of 16 up to 192 (48, 80, 96, 112, 128, 144, 160, 176, 192).
// (The reference's own default ViT.Attention(dim=768, num_heads=4) has 192-wide heads: ViT.py:67, 121-127.)
//
// A head row of HD 16-bit values is wider than one swizzle atom, so Q / K / V tiles live in shared memory as NP = HD / W
// column panels of W elements (W = 64 -> 128-byte swizzle, W = 32 -> 64-byte swizzle, W = 16 -> 32-byte swizzle; 96 = 3 x 32,
// 48 = 3 x 16: the zoo's moat_0 has 48-wide heads, moat.py:144-146), one TMA box per panel; the S MMA accumulates over the panels, the PV MMA writes one W-column group of O per panel.
// TMEM (256 columns, so two CTAs fit on an SM): S fp32 [0, 64) for one block of 64 keys, fp16 P written over it [0, 32),
// O fp32 [64, 64 + HD) -- no aliasing, so HD = 192 fits exactly.  Keys are processed in blocks of 64 with an online softmax;
// one thread owns one query row and keeps the whole S block in registers (64 values): a single TMEM read per score.
// 256 threads: warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4-7 softmax + epilogue.
// Unit of work = (group, head, 128-row query tile); K / V are single-buffered (K is free again as soon as the S MMA of the
// block has retired, V after the PV MMA: both long before the next load is needed).  Output rows go straight from registers
// to global memory (every thread holds whole rows: HD * 2 contiguous bytes).
#pragma once
#include "pa_attn.cuh"

namespace pa {

constexpr int AW_THREADS = 256;
constexpr int AW_KB = 64;         // keys per block
constexpr int AW_O_COL = 64;

template <int HD>
struct AttnWideCfg {
  static constexpr int W = (HD % 64 == 0) ? 64 : (HD % 32 == 0) ? 32 : 16;
  static constexpr int NP = HD / W;
  static constexpr int ROW_BYTES = W * 2;
  static constexpr int SBO = 8 * ROW_BYTES;
  static constexpr uint64_t SWZ = (W == 64) ? PA_SWZ_128B : (W == 32) ? PA_SWZ_64B : PA_SWZ_32B;
  static constexpr int OSTEP = (HD % 32 == 0) ? 32 : 16;          // O columns per TMEM round trip in the rescale / read-out loops
  static constexpr int Q_PANEL = 128 * ROW_BYTES;
  static constexpr int KV_PANEL = AW_KB * ROW_BYTES;
  static constexpr int V_KSTEP = 16 * ROW_BYTES / 16;
  static constexpr int Q_BYTES = NP * Q_PANEL, KV_BYTES = NP * KV_PANEL;
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_BYTES + 128 + 1024;
  static_assert(HD % 16 == 0 && HD >= 16 && AW_O_COL + HD <= 256, "head_dim must be a multiple of 16, at most 192");
};

struct AttnWideParams {
  int G, H, n_q, n_k, nkb, q_tiles, units;
  int q_col0, k_col0, v_col0;
  void* O;
  long long ldo, o_group;
  int o_col0;
  float scale_log2e;
  uint32_t idesc_s, idesc_o;
};

__device__ __forceinline__ void tmem_st32v(uint32_t taddr, const uint32_t (&v)[32]) {
  uint32_t a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = v[i]; b[i] = v[16 + i]; }
  tmem_st16(taddr, a);
  tmem_st16(taddr + 16, b);
}

template <int HD>
__global__ void __launch_bounds__(AW_THREADS, 2)
attn_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnWideParams p) {
  using Cfg = AttnWideCfg<HD>;
  extern __shared__ uint8_t aw_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(aw_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* k_smem = q_smem + Cfg::Q_BYTES;
  uint8_t* v_smem = k_smem + Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + Cfg::KV_BYTES);
  uint64_t* q_full = bars;          // per unit
  uint64_t* q_empty = bars + 1;     // per unit: last S MMA of the unit retired
  uint64_t* o_read = bars + 2;      // per unit: O is in registers (4 warps)
  uint64_t* k_full = bars + 3;      // per block
  uint64_t* k_empty = bars + 4;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;      // 4 warps: P written, O rescaled
  uint64_t* o_full = bars + 9;      // PV of the block retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 10; ++i) mbar_init(&bars[i], (i == 2 || i == 8) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int ui = 0, bc = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++ui) {
      const int qt = u % p.q_tiles, gh = u / p.q_tiles;
      const int h = gh % p.H, g = gh / p.H;
      mbar_wait(q_empty, (ui & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int pn = 0; pn < Cfg::NP; ++pn)
          tma_load_3d(q_smem + pn * Cfg::Q_PANEL, &tmQ, p.q_col0 + h * HD + pn * Cfg::W, qt * 128, g, q_full);
      }
      __syncwarp();
      for (int j = 0; j < p.nkb; ++j, ++bc) {
        mbar_wait(k_empty, (bc & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(k_full, Cfg::KV_BYTES);
#pragma unroll
          for (int pn = 0; pn < Cfg::NP; ++pn)
            tma_load_3d(k_smem + pn * Cfg::KV_PANEL, &tmK, p.k_col0 + h * HD + pn * Cfg::W, j * AW_KB, g, k_full);
        }
        __syncwarp();
        mbar_wait(v_empty, (bc & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(v_full, Cfg::KV_BYTES);
#pragma unroll
          for (int pn = 0; pn < Cfg::NP; ++pn)
            tma_load_3d(v_smem + pn * Cfg::KV_PANEL, &tmV, p.v_col0 + h * HD + pn * Cfg::W, j * AW_KB, g, v_full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t q_base = smem_u32(q_smem), k_base = smem_u32(k_smem), v_base = smem_u32(v_smem);
    int ui = 0, bc = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++ui) {
      mbar_wait(q_full, ui & 1);
      for (int j = 0; j < p.nkb; ++j, ++bc) {
        mbar_wait(k_full, bc & 1);
        tc_fence_after();
        // S = Q K^T over the panels.  It overwrites the P of the previous block: the tensor pipe executes in issue order, so
        // that block's PV (issued before) has read P by then; the softmax warps were done with S before they signalled p_full.
        if (elect_one()) {
#pragma unroll
          for (int pn = 0; pn < Cfg::NP; ++pn) {
            const uint64_t qdesc = make_sdesc(q_base + pn * Cfg::Q_PANEL, 16, Cfg::SBO, Cfg::SWZ);
            const uint64_t kdesc = make_sdesc(k_base + pn * Cfg::KV_PANEL, 16, Cfg::SBO, Cfg::SWZ);
#pragma unroll
            for (int k = 0; k < Cfg::W / 16; ++k) umma_ss(tmem_base, qdesc + 2 * k, kdesc + 2 * k, p.idesc_s, (pn | k) != 0);
          }
          umma_commit(s_full);
          umma_commit(k_empty);
          if (j == p.nkb - 1) umma_commit(q_empty);
        }
        __syncwarp();
        mbar_wait(p_full, bc & 1);
        mbar_wait(v_full, bc & 1);
        if (j == 0 && ui > 0) mbar_wait(o_read, (ui - 1) & 1);     // the previous unit's O has been read out
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int pn = 0; pn < Cfg::NP; ++pn) {
            const uint64_t vdesc = make_sdesc(v_base + pn * Cfg::KV_PANEL, Cfg::SBO, Cfg::SBO, Cfg::SWZ);   // V [key][d]: MN-major B
#pragma unroll
            for (int k = 0; k < AW_KB / 16; ++k)
              umma_ts(tmem_base + AW_O_COL + pn * Cfg::W, tmem_base + 8 * k, vdesc + Cfg::V_KSTEP * k, p.idesc_o, (j | k) != 0);
          }
          umma_commit(o_full);
          umma_commit(v_empty);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: one thread per query row =====================
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = p.scale_log2e;
    int ui = 0, bc = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++ui) {
      const int qt = u % p.q_tiles, gh = u / p.q_tiles;
      const int h = gh % p.H, g = gh / p.H;
      const int row = qt * 128 + trow;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < p.nkb; ++j, ++bc) {
        const int nvalid = min(AW_KB, p.n_k - j * AW_KB);
        mbar_wait(s_full, bc & 1);
        tc_fence_after();
        uint32_t s[64];
        tmem_ld64(t_row, s);
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, (i < nvalid) ? __uint_as_float(s[i]) : -INFINITY);
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2f((m_run - m_new) * sl2);       // 0 on the first block
        const float mxs = m_new * sl2;
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float e0 = 0.f, e1 = 0.f;
          if (i < nvalid) e0 = ex2f(fmaf(__uint_as_float(s[i]), sl2, -mxs));          // nvalid is uniform: no divergence
          if (i + 1 < nvalid) e1 = ex2f(fmaf(__uint_as_float(s[i + 1]), sl2, -mxs));
          sum0 += e0;
          sum1 += e1;
          pk[i >> 1] = pack_h2(e0, e1);
        }
        l_run = l_run * alpha + (sum0 + sum1);
        m_run = m_new;
        if (j > 0) {
          // the previous block's PV must have retired before O is rescaled and its P overwritten
          mbar_wait(o_full, (bc - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < HD; c += Cfg::OSTEP) {
            uint32_t o[Cfg::OSTEP];
            if constexpr (Cfg::OSTEP == 32) tmem_ld32(t_row + AW_O_COL + c, o); else tmem_ld16(t_row + AW_O_COL + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < Cfg::OSTEP; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            if constexpr (Cfg::OSTEP == 32) tmem_st32v(t_row + AW_O_COL + c, o); else tmem_st16(t_row + AW_O_COL + c, o);
          }
        }
        tmem_st32v(t_row, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: O / rowsum -> fp16 -> global, 32 columns at a time
      mbar_wait(o_full, (bc - 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l_run;
      uint16_t* dst = reinterpret_cast<uint16_t*>(p.O) + (long long)g * p.o_group + (long long)row * p.ldo + p.o_col0 + h * HD;
#pragma unroll 1
      for (int c = 0; c < HD; c += Cfg::OSTEP) {
        uint32_t o[Cfg::OSTEP];
        if constexpr (Cfg::OSTEP == 32) tmem_ld32(t_row + AW_O_COL + c, o); else tmem_ld16(t_row + AW_O_COL + c, o);
        tmem_ld_wait();
        if (c + Cfg::OSTEP >= HD) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_read);          // the next unit's first PV may overwrite O
        }
        if (row < p.n_q) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c);
#pragma unroll
          for (int i = 0; i < Cfg::OSTEP / 8; ++i) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(o[8 * i + k]) * inv;
            d4[i] = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

}  // namespace pa
