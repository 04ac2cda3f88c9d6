// pa_attn_win.cuh — cswin.LePEAttention.forward (cswin.py:101-127) as ONE kernel: windowed softmax(q k^T scale) v PLUS the
// locally-enhanced positional encoding lepe = depthwise3x3(v) zero-padded at the window borders (cswin.py:86-99, :121).
//
// Unit of work = (cross-shaped window, head).  The whole window's Q, K and V (<= 512 / 448 / 448 rows of head_dim 16-bit values)
// are gathered ONCE by 5-D TMA boxes straight from the [B, H*W, ld] token matrix (no img2windows copy) and stay resident in
// shared memory while the CTA walks the window's 128-row query tiles and, per tile, its key blocks (online softmax, O rescaled in
// TMEM).  Because V is resident, the epilogue computes the 3 x 3 LePE term for its own tokens from shared memory: the separate
// LePE kernel (380 us per branch at BASELINE config 4) and the read-modify-write of the output it needed are gone.
// Two single-slot CTAs per SM (256 TMEM columns, <= 113 KB shared memory, 384 threads each): warp 0 TMA | warp 1 MMA | warp 2
// TMEM allocator | warps 4-11 softmax + epilogue, two threads per query row (see pa_cosched.cuh for why this shape beats the
// two-slot CTA of attn_core_kernel).  TMEM slot: S fp32 [0, kb) -> fp16 P in place; O fp32 [256 - HD, 256), kb <= 256 - HD.
// Measured and rejected: the 8 query rows past the last full tile of a 392-token window (3 x 128 + 8) on the CUDA cores, one warp
// per row, instead of a fourth tensor-core tile -- 1280 vs 1271 us per branch: a tile whose other 120 rows are inactive costs
// next to nothing, the kernel is bound by the softmax work of the populated rows, not by the number of chains.
#pragma once
#include "pa_attn.cuh"
#include "pa_cosched.cuh"

namespace pa {

struct AttnWinParams {
  AttnParams at;              // geometry, tensor-map columns, scale, instruction descriptors (windowed fields filled)
  const float* lepe_w;        // [9, Cb] fp32: get_v.weight transposed (tap-major)
  const float* lepe_b;        // [Cb]
  int Cb;                     // channels of this branch (= H * HD)
  int q_rows, kv_rows;        // rows of the resident Q / K / V buffers
};

__host__ __device__ inline int attn_win_q_rows(int q_tiles, int nkb, int kb_rows) {
  const int a = q_tiles * 128, b = nkb * kb_rows;
  return ((a > b ? a : b) + 7) / 8 * 8;
}
__host__ __device__ inline int attn_win_kv_rows(int nkb, int kb_rows, int kb) { return ((nkb - 1) * kb_rows + kb + 7) / 8 * 8; }
__host__ __device__ inline int attn_win_data_bytes(int hd, int q_rows, int kv_rows, int Cb) {
  return (q_rows + 2 * kv_rows) * hd * 2 + 2048 /* max / sum exchange */ + 10 * Cb * 4 /* LePE taps + bias */;
}

template <int HD>
__global__ void __launch_bounds__(CS_THREADS, 2)
attn_win_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnWinParams P) {
  using Cfg = AttnCfg<HD>;
  const AttnParams& p = P.at;
  extern __shared__ uint8_t aw_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(aw_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* k_smem = q_smem + P.q_rows * Cfg::ROW_BYTES;
  uint8_t* v_smem = k_smem + P.kv_rows * Cfg::ROW_BYTES;
  float* xch = reinterpret_cast<float*>(v_smem + P.kv_rows * Cfg::ROW_BYTES);     // xmax[2][128], xsum[2][128]
  float* lw = xch + 512;                                                          // [9][Cb] taps, then [Cb] bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(lw + 10 * P.Cb);
  uint64_t* qkv_full = bars;        // per unit: the window's Q, K, V have landed
  uint64_t* qkv_empty = bars + 1;   // per unit: last MMA retired (1) + the eight softmax warps are done with V (8)
  uint64_t* s_full = bars + 2;      // per key block
  uint64_t* p_full = bars + 3;      // per key block, 8 warps
  uint64_t* o_full = bars + 4;      // per key block
  uint64_t* slot_empty = bars + 5;  // per query tile, 8 warps: O is in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // rows never touched by TMA (tile padding past the window) must read as finite zeros; LePE taps + bias of the whole branch
  {
    const int data16 = (P.q_rows + 2 * P.kv_rows) * Cfg::ROW_BYTES / 16;
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < data16; i += CS_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 9 * P.Cb; i += CS_THREADS) lw[i] = __ldg(P.lepe_w + i);
    for (int i = threadIdx.x; i < P.Cb; i += CS_THREADS) lw[9 * P.Cb + i] = __ldg(P.lepe_b + i);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 1 && lane == 0) {
    mbar_init(qkv_full, 1);
    mbar_init(qkv_empty, 9);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(slot_empty, 8);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = p.G * p.H;

  if (warp < 4) {
    cs_regs_control();
    if (warp == 0) {
      // ===================== TMA producer: one window of one head per unit =====================
      int ui = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        const int h = u % p.H, g = u / p.H;
        const int img = g / p.nWin, w = g - img * p.nWin;
        const int c3 = w % p.nJ;                              // window column
        const int c4 = img * (p.nWin / p.nJ) + w / p.nJ;      // image * window rows + window row
        mbar_wait(qkv_empty, (ui & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(qkv_full, 3 * p.nkb * p.kb_rows * Cfg::ROW_BYTES);
          for (int j = 0; j < p.nkb; ++j) {
            const int off = j * p.kb_rows * Cfg::ROW_BYTES;
            tma_load_5d(q_smem + off, &tmQ, p.q_col0 + h * HD, 0, c3, j * p.h_box, c4, qkv_full);
            tma_load_5d(k_smem + off, &tmK, p.k_col0 + h * HD, 0, c3, j * p.h_box, c4, qkv_full);
            tma_load_5d(v_smem + off, &tmV, p.v_col0 + h * HD, 0, c3, j * p.h_box, c4, qkv_full);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const int ksteps_o = p.kb / 16;
      const int h16 = (ksteps_o + 1) / 2;
      const uint32_t q_base = smem_u32(q_smem), k_base = smem_u32(k_smem), v_base = smem_u32(v_smem);
      int ui = 0, bc = 0, tc = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        mbar_wait(qkv_full, ui & 1);
        tc_fence_after();
        for (int qt = 0; qt < p.q_tiles; ++qt, ++tc) {
          const uint64_t qdesc = make_sdesc(q_base + qt * Cfg::Q_TILE_BYTES, 16, Cfg::SBO, Cfg::SWZ);
          for (int j = 0; j < p.nkb; ++j, ++bc) {
            const uint32_t boff = j * p.kb_rows * Cfg::ROW_BYTES;
            const uint64_t kdesc = make_sdesc(k_base + boff, 16, Cfg::SBO, Cfg::SWZ);
            // S overwrites the previous block's P: the tensor pipe executes in issue order, that block's PV was issued before.
            // With a single wide key block S also covers the O columns: then the previous tile's O must have been read out first.
            bool slot_waited = false;
            if (j == 0 && tc > 0 && p.kb > Cfg::O_COL) {
              mbar_wait(slot_empty, (tc - 1) & 1);
              tc_fence_after();
              slot_waited = true;
            }
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) umma_ss(tmem_base, qdesc + 2 * k, kdesc + 2 * k, p.idesc_s, k != 0);
              umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_full, bc & 1);
            if (j == 0 && tc > 0 && !slot_waited) mbar_wait(slot_empty, (tc - 1) & 1);     // the previous tile's O has been read out
            tc_fence_after();
            if (elect_one()) {
              const uint64_t vdesc = make_sdesc(v_base + boff, Cfg::SBO, Cfg::SBO, Cfg::SWZ);   // V [key][d]: MN-major B operand
              for (int k = 0; k < ksteps_o; ++k) {
                const int pcol = (k < h16) ? 8 * k : 16 * h16 + 8 * (k - h16);
                umma_ts(tmem_base + Cfg::O_COL, tmem_base + pcol, vdesc + Cfg::V_KSTEP * k, p.idesc_o, (j | k) != 0);
              }
              umma_commit(o_full);
              if (qt == p.q_tiles - 1 && j == p.nkb - 1) umma_commit(qkv_empty);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    cs_regs_worker();
    // ===================== softmax + epilogue (+ LePE): two threads per query row =====================
    const int sw = warp - 4;
    const int hf = sw >> 2;
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t t_slot = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n16 = p.kb >> 4;
    const int h16 = (n16 + 1) / 2;
    const int c_lo = hf ? h16 * 16 : 0;
    const int nst = hf ? n16 - h16 : h16;
    const uint32_t t_my = t_slot + c_lo;
    const float sl2 = p.scale_log2e;
    float* xmax = xch;
    float* xsum = xch + 256;
    constexpr int OH = HD / 2;
    const uint32_t t_o = t_slot + Cfg::O_COL + hf * OH;
    int ui = 0, bc = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
      const int h = u % p.H, g = u / p.H;
      const int img = g / p.nWin, w = g - img * p.nWin;
      const int wi = w / p.nJ, wj = w - wi * p.nJ;
      const float* wt = lw + h * HD + hf * OH;               // this thread's channels: tap t at wt[t * Cb + i]
      for (int qt = 0; qt < p.q_tiles; ++qt) {
        const int row = qt * 128 + trow;                     // token index inside the window
        const bool warp_active = (qt * 128 + q * 32) < p.n_q;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < p.nkb; ++j, ++bc) {
          const int nvalid = min(p.kb_rows, p.n_k - j * p.kb_rows) - c_lo;
          mbar_wait(s_full, bc & 1);
          tc_fence_after();
          float mx = -INFINITY;
          if (warp_active) {
            int k = 0;
#pragma unroll 1
            for (; k + 1 < nst; k += 2) {
              uint32_t v[32];
              tmem_ld32(t_my + k * 16, v);
              tmem_ld_wait();
              mx = chunk_max<32>(v, nvalid - k * 16, mx);
            }
            if (k < nst) {
              uint32_t v[16];
              tmem_ld16(t_my + k * 16, v);
              tmem_ld_wait();
              mx = chunk_max<16>(v, nvalid - k * 16, mx);
            }
          }
          xmax[hf * 128 + trow] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mx = fmaxf(mx, xmax[(hf ^ 1) * 128 + trow]);        // (rewritten only after the next S, i.e. after every warp's p_full arrival)
          const float m_new = fmaxf(m_run, mx);
          const float alpha = ex2f((m_run - m_new) * sl2);    // 0 on the first block
          if (j > 0) {
            mbar_wait(o_full, (bc - 1) & 1);                  // previous block's PV retired: O may be rescaled, P overwritten
            tc_fence_after();
            if (warp_active) {
#pragma unroll
              for (int c = 0; c < OH; c += 16) {
                uint32_t o[16];
                tmem_ld16(t_o + c, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st16(t_o + c, o);
              }
            }
          }
          if (warp_active) {
            const float mxs = m_new * sl2;
            uint32_t va[16], pk[8];
            float e[16];
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
            for (int k = 0; k < nst; ++k) {
              tmem_ld16(t_my + k * 16, va);
              tmem_ld_wait();
              exp_stage(va, e, nvalid - k * 16, sl2, mxs);
              pack_stage(e, pk, s0, s1);
              tmem_st8(t_my + k * 8, pk);
            }
            tmem_st_wait();
            l_run = l_run * alpha + (s0 + s1);
            m_run = m_new;
          }
          if (j == p.nkb - 1) xsum[hf * 128 + trow] = l_run;   // ordered towards the partner by p_full -> o_full
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / rowsum + LePE -> fp16 -> image position of the token
        mbar_wait(o_full, (bc - 1) & 1);
        tc_fence_after();
        uint32_t v[OH];
        float l_other = 0.f;
        if (warp_active) {
          if (OH == 32) tmem_ld32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          else tmem_ld16(t_o, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld_wait();
          l_other = xsum[(hf ^ 1) * 128 + trow];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(slot_empty);
        if (warp_active && row < p.n_q) {
          const float inv = 1.f / (l_run + l_other);
          const int rr = row / p.W_sp, cc = row - rr * p.W_sp;
          float acc[OH];
#pragma unroll
          for (int i = 0; i < OH; ++i) acc[i] = __uint_as_float(v[i]) * inv + wt[9 * P.Cb + i];      // attention + get_v bias
#pragma unroll
          for (int du = 0; du < 3; ++du) {
            const int r2 = rr + du - 1;
            if (r2 < 0 || r2 >= p.H_sp) continue;             // zero padding at the WINDOW border (cswin.py:93-96)
#pragma unroll
            for (int dv = 0; dv < 3; ++dv) {
              const int c2 = cc + dv - 1;
              if (c2 < 0 || c2 >= p.W_sp) continue;
              const int t2 = r2 * p.W_sp + c2;
              const uint8_t* vrow = v_smem + t2 * Cfg::ROW_BYTES;
              const float* wtap = wt + (du * 3 + dv) * P.Cb;
#pragma unroll
              for (int ch = 0; ch < OH / 8; ++ch) {
                const int cid = hf * (OH / 8) + ch;                                              // 16-byte chunk of the row
                const int phys = (HD == 64) ? (cid ^ (t2 & 7)) : (cid ^ ((t2 >> 1) & 3));        // SW128 / SW64
                const uint4 u4 = *reinterpret_cast<const uint4*>(vrow + (phys << 4));
                const __half2* h2 = reinterpret_cast<const __half2*>(&u4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  acc[8 * ch + 2 * k] = fmaf(__low2float(h2[k]), wtap[8 * ch + 2 * k], acc[8 * ch + 2 * k]);
                  acc[8 * ch + 2 * k + 1] = fmaf(__high2float(h2[k]), wtap[8 * ch + 2 * k + 1], acc[8 * ch + 2 * k + 1]);
                }
              }
            }
          }
          const long long tok = (long long)(wi * p.H_sp + rr) * p.R + wj * p.W_sp + cc;
          uint4* d4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.O) + (long long)img * p.o_group + tok * p.ldo + p.o_col0 +
                                               h * HD + hf * OH);
#pragma unroll
          for (int i = 0; i < OH / 8; ++i)
            d4[i] = make_uint4(pack_h2(acc[8 * i], acc[8 * i + 1]), pack_h2(acc[8 * i + 2], acc[8 * i + 3]),
                               pack_h2(acc[8 * i + 4], acc[8 * i + 5]), pack_h2(acc[8 * i + 6], acc[8 * i + 7]));
        }
        if (qt == p.q_tiles - 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(qkv_empty);               // this warp's last read of the resident V
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
