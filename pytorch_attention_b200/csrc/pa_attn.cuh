// pa_attn.cuh — softmax(Q K^T * scale) V core on tcgen05 + TMEM.
//
// One templated persistent kernel serves every variant:
//   HD        head dim 64 (ViT / PVT / CvT; 128B-swizzled rows) or 32 (CSWin; 64B-swizzled rows)
//   WINDOWED  false: a unit's rows are consecutive tokens of a group (3-D TMA {cols, rows, group});
//             true : a unit is one cross-shaped window of a CSWin image; rows are gathered straight from the
//                    [B, H*W, ld] token matrix by a 5-D TMA box over {chan, col-in-window, window-col, row, image*window-row}
//                    (box {HD, W_sp, 1, h_box, 1}; no img2windows copy),
//                    and the epilogue scatters back to image order, adding onto the LePE term already there.
// Work item = (group/window, head, pair of 128-row query tiles).  Keys are processed in blocks of <= 256-HD
// (a single block may be up to 256 wide): one block -> exact single-pass softmax; several blocks -> online
// softmax with the running O rescaled in TMEM (cheap: HD columns per row).
//
// 640 threads:  warp 0 TMA producer (+ TMEM allocator) | warp 1 MMA issuer | warps 18-19 idle (register allocation granularity) |
//               warps 2-9 softmax+epilogue of slot 0 | warps 10-17 of slot 1.  TWO threads per query row (a warp may
//               touch TMEM lanes 32*(warp%4)..+31, so warps w and w+4 of a slot share a lane quarter): each owns one
//               half of the row's S columns end to end -- partial max (exchanged through smem), exponentials, its half
//               of P written IN PLACE inside its own S columns, partial row sum, its half of the O columns.
//               Four softmax warps per scheduler instead of two: a single warp cannot keep the MUFU busy (measured).
// Softmax passes: max pass with 32-column TMEM loads; exponential pass in 16-column steps, 12 of 16 exponentials on the MUFU
//               and 4 on the FMA pipe (degree-3 polynomial), never overlapping a TMEM load with the step's compute (measured).
// TMEM slot (256 columns): S fp32 [0,kb)  ->  P fp16x2 written in place inside each thread's own S columns (two segments);
//                          O fp32 [256-HD, 256) (aliases the tail of S only in the single-block case, where the
//                          PV MMAs start after the softmax has drained S).
#pragma once
#include "pa_ptx.cuh"

namespace pa {

struct AttnParams {
  int G, H;             // groups (batch entries or windows), heads
  int n_q, n_k;         // query / key rows per unit
  int kb;               // S tile width = key rows per block buffer, multiple of 16
  int kb_rows;          // rows a K/V TMA box delivers per block (== kb unless windowed)
  int nkb;              // key blocks per unit
  int q_tiles, pairs, items;
  int q_col0, k_col0, v_col0;   // element column of head 0 inside the Q / KV tensor maps
  void* O;              // fp16 output
  long long ldo, o_group;
  int o_col0;
  float scale_log2e;    // softmax scale * log2(e)
  uint32_t idesc_s, idesc_o;
  // windowed geometry (CSWin): image R x R, windows H_sp x W_sp, nJ windows per image row, nWin per image
  int R, H_sp, W_sp, nJ, nWin, h_box;
  int add_into_out;     // epilogue adds onto what O already holds (LePE)
  long long* trace;     // debug: clock64 stamps of CTA 0 ([item][16]), or nullptr
  int tma_store;        // non-windowed: output tile goes through smem + TMA store (tmO valid)
  int debug_flags;      // timing experiments only (env PA_ATTN_DEBUG): 1 = no MUFU in pass 2, 2 = no P store, 4 = skip pass 1
  // dependency hooks of the fused single-launch kernels (nullptr = none)
  const int* wait_ctr;  // before loading a unit of group g: wait_ctr[t] >= wait_target for the 128-row tiles t covering its rows
  int wait_target;
  long long wait_rows_per_group;   // rows of the Q/K/V buffer per group (n_q == n_k assumed by the hook)
  int* signal_ctr;      // after a (head, query tile) of group g has been stored completely: signal_ctr[g] += 1
  int cta_shift;        // fused kernels: CTA c walks the item sequence of virtual CTA (c - cta_shift) mod grid, so that the
                        // CTAs holding this phase's remainder items are not the ones holding the other phases' remainders
  // additive score bias before the softmax (cmt.py:100: q k^T * scale + relative_pos): fp32 [H, n_q, n_k], or nullptr.
  // Single-slot kernel only.  rel_mul = 1 / scale, so that (s + r * rel_mul) * scale == s * scale + r.
  const float* rel_pos;
  float rel_mul;
  // per-row score threshold (kvt.KNNAttention, kvt.py:84-87: only the top-k scores of a row take part in the softmax): raw
  // (unscaled) scores below row_thresh[(g * H + h) * n_q + row] count as -inf.  fp32 [G, H, n_q], or nullptr.  Single-slot kernel only.
  const float* row_thresh;
};

#define ATTN_TRACE(seq, slot_) do { if (p.trace != nullptr && blockIdx.x == 0 && (seq) < 32) p.trace[(seq) * 16 + (slot_)] = clock64(); } while (0)

constexpr int ATTN_THREADS = 640;   // 2 control warps + 2 slots x 8 softmax warps (two threads per query row) + 2 idle:
                                    // registers are allocated to warps in groups of four, so 18 warps cost as much as 20
constexpr int ATTN_SLOT_COLS = 256;

template <int HD>
struct AttnCfg {
  static constexpr int ROW_BYTES = HD * 2;                       // 128 (SW128) or 64 (SW64)
  static constexpr int SBO = 8 * ROW_BYTES;                      // 8-row swizzle atom
  static constexpr uint64_t SWZ = (HD == 64) ? PA_SWZ_128B : PA_SWZ_64B;
  static constexpr int Q_TILE_BYTES = 128 * ROW_BYTES;
  static constexpr int O_COL = ATTN_SLOT_COLS - HD;
  static constexpr int V_KSTEP = 16 * ROW_BYTES / 16;            // descriptor advance (16 B units) per 16 keys
};

// Shared memory plan (host and device agree through these helpers)
__host__ __device__ inline int attn_q_rows(bool windowed, int nkb, int kb_rows) {
  const int r = windowed ? nkb * kb_rows : 256;
  return ((r < 512 ? 512 : r) + 7) / 8 * 8;      // windowed: last query tile may start at row 384 -> keep 512 rows mapped
}
__host__ __device__ inline int attn_ostage_bytes(int hd, bool staged) { return (staged ? 2 * 128 * hd * 2 : 0) + 4096; }   // + max/sum exchange
__host__ __device__ inline int attn_smem_bytes(int hd, bool windowed, int nkb, int kb, int kb_rows, bool staged) {
  const int q_rows = windowed ? attn_q_rows(true, nkb, kb_rows) : 256;
  return 2 * q_rows * hd * 2 + 2 * 2 * kb * hd * 2 + attn_ostage_bytes(hd, staged) + 256 + 1024;
}

// ---- per-chunk softmax helpers (one thread = one query row; v = N consecutive S columns of that row)
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3
  return r;
}
template <int N>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[N], int nval, float mx) {
  if (nval >= N) {
#pragma unroll
    for (int i = 0; i < N; i += 2) mx = max3f(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) mx = fmaxf(mx, (i < nval) ? __uint_as_float(v[i]) : -INFINITY);
  }
  return mx;
}
// p = exp2(s * sl2 - mxs) for the first nval columns (0 beyond), packed to fp16 pairs; returns the row-partial sum
template <int N>
__device__ __forceinline__ float chunk_exp(const uint32_t (&v)[N], uint32_t (&pk)[N / 2], int nval, float sl2, float mxs) {
  float s0 = 0.f, s1 = 0.f;
  if (nval >= N) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float e0 = ex2f(fmaf(__uint_as_float(v[i]), sl2, -mxs));
      const float e1 = ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -mxs));
      s0 += e0;
      s1 += e1;
      pk[i >> 1] = pack_h2(e0, e1);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      float e0 = 0.f, e1 = 0.f;
      if (i < nval) e0 = ex2f(fmaf(__uint_as_float(v[i]), sl2, -mxs));          // nval is warp-uniform: no divergence,
      if (i + 1 < nval) e1 = ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -mxs));  // and no MUFU work on padded keys
      s0 += e0;
      s1 += e1;
      pk[i >> 1] = pack_h2(e0, e1);
    }
  }
  return s0 + s1;
}

// exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], max rel. error 7.6e-5 -- below the fp16
// rounding of P): the MUFU unit (16 ex2 / clk / SM) is the limiter of pass 2, so a fraction of each step's exponentials
// is computed here instead.  x <= 0 up to rounding; clamped so the result stays a normal number (2^-125 ~ 0).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;          // 1.5 * 2^23: the nearest integer n lands in the low mantissa bits
  const float f = x - (t - 12582912.f);    // [-0.5, 0.5]
  float q = fmaf(0.0551695712f, f, 0.242606208f);
  q = fmaf(q, f, 0.693260908f);
  q = fmaf(q, f, 0.999928415f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));   // * 2^n
}

// ---- pass 2 of the row softmax, one 16-column step:  p = exp2(s * sl2 - mxs), fp16 P packed in pairs, partial row sums.
// PA_EXP_F16X2=1 (opt-in at compile time): the exponent x is computed in fp32, rounded to a half pair and exponentiated by ONE packed MUFU
// op per two scores (ex2.approx.f16x2); the result IS the packed fp16 P.  The exp pass is MUFU-bound (two softmax warps per
// scheduler, 16 ex2 / clk / SM): half the MUFU instructions.  Accuracy: P is rounded to fp16 either way; rounding x to half adds a
// relative error of ln2 * ulp(x) / 2 <= 3.4e-4 for the terms that matter (x in [-2, 0]) and more only where 2^x < 2^-2^k is
// small in proportion -- measured against the oracle in tests/ (bar 1e-3).
// PA_EXP_F16X2=0 (default): fp32 MUFU ex2 for 12 of 16 scores and a degree-3 polynomial on the FMA pipe for the other 4 (round 1).
// MEASURED (late round 2): parity green with PA_EXP_F16X2=1 (142 GPU tests), speed unchanged or slightly worse (ViT-B attention core
// 28.1 -> 29.2 us, CSWin C4 3669 -> 3722 us): ptxas splits ex2.approx.f16x2 into TWO MUFU.EX2.F16 -- 16 MUFU ops per 16 scores
// against 12 + 4 polynomials -- and MUFU.EX2.F16 issues at the fp32 rate on sm_100a.  Default stays 0.
#ifndef PA_EXP_F16X2
#define PA_EXP_F16X2 0
#endif
// Stage A: e = the exponent (PA_EXP_F16X2) or the exponential (fp32 path) of the first nval columns (-inf / 0 beyond)
__device__ __forceinline__ void exp_stage(const uint32_t (&v)[16], float (&e)[16], int nval, float sl2, float mxs, int dbg = 0) {
#if PA_EXP_F16X2
  if (nval >= 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = fmaf(__uint_as_float(v[i]), sl2, -mxs);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = (i < nval) ? fmaf(__uint_as_float(v[i]), sl2, -mxs) : -INFINITY;   // 2^-inf = 0: padded keys
  }
#else
  if (dbg & 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = fmaf(__uint_as_float(v[i]), sl2, -mxs);
  } else if (nval >= 16) {
    // 4 of 16 on the FMA pipe (measured: 4 -> -5 %, 6 and 8 of 16 -> no gain: the FMA pipe / issue slots fill up)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = fmaf(__uint_as_float(v[i]), sl2, -mxs);
      e[i] = ((i & 3) == 3) ? ex2_poly(x) : ex2f(x);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      e[i] = 0.f;
      if (i < nval) e[i] = ex2f(fmaf(__uint_as_float(v[i]), sl2, -mxs));   // nval is warp-uniform; padded keys never reach the MUFU
    }
  }
#endif
}
// Stage B: fp16 pack (PA_EXP_F16X2: + the packed exponential) and row-partial sums of a step
__device__ __forceinline__ void pack_stage(const float (&e)[16], uint32_t (&pk)[8], float& s0, float& s1) {
#if PA_EXP_F16X2
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    __half2 h = __floats2half2_rn(e[i], e[i + 1]);
    uint32_t r;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<uint32_t*>(&h)));
    pk[i >> 1] = r;
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&r));
    s0 += f.x;
    s1 += f.y;
  }
#else
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    s0 += e[i];
    s1 += e[i + 1];
    pk[i >> 1] = pack_h2(e[i], e[i + 1]);
  }
#endif
}

__device__ __forceinline__ void attn_init_barriers(uint64_t* bars) {     // one thread; bars: 16 mbarriers
  for (int i = 0; i < 2; ++i) {
    mbar_init(&bars[0 + i], 1);    // q_full
    mbar_init(&bars[2 + i], 1);    // q_empty
    mbar_init(&bars[4 + i], 1);    // kv_full
    mbar_init(&bars[6 + i], 1);    // kv_empty
    mbar_init(&bars[8 + i], 1);    // s_full
    mbar_init(&bars[10 + i], 8);   // p_full
    mbar_init(&bars[12 + i], 1);   // o_full
    mbar_init(&bars[14 + i], 8);   // slot_empty
  }
}

// All roles of one CTA over its whole item sequence.  Barriers initialised + visible and TMEM (512 columns) allocated
// before the call; every thread of the CTA calls it.
template <int HD, bool WINDOWED>
__device__ __forceinline__ void attn_run(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                         const CUtensorMap& tmO, const AttnParams& p, uint8_t* smem, uint64_t* bars,
                                         uint32_t tmem_base) {
  using Cfg = AttnCfg<HD>;
  const int q_rows = WINDOWED ? attn_q_rows(true, p.nkb, p.kb_rows) : 256;
  const int q_bytes = q_rows * Cfg::ROW_BYTES;          // one Q buffer
  const int kvb_bytes = p.kb * Cfg::ROW_BYTES;          // one K (or V) block buffer
  uint8_t* q_smem = smem;                               // [2][q_bytes]
  uint8_t* kv_smem = smem + 2 * q_bytes;                // [2 stages][K | V]
  uint8_t* o_smem = kv_smem + 4 * kvb_bytes;            // [2 slots][128 rows x HD fp16] output staging (non-windowed)
  uint64_t* q_full = bars;            // [2] TMA -> MMA
  uint64_t* q_empty = bars + 2;       // [2] MMA -> TMA
  uint64_t* kv_full = bars + 4;       // [2]
  uint64_t* kv_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;        // [2] MMA -> softmax(slot): S block ready
  uint64_t* p_full = bars + 10;       // [2] softmax(slot) -> MMA: P written (and O rescaled)
  uint64_t* o_full = bars + 12;       // [2] MMA -> softmax(slot): PV of the block retired
  uint64_t* slot_empty = bars + 14;   // [2] epilogue(slot) -> MMA

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int pend_sig = -1;                  // fused kernels: group whose output tile is stored but not yet published (slot leader)

  if (warp == 0) {
    // ===================== TMA producer (warp converged; one elected lane issues) =====================
    int qb = 0, st = 0;
    uint32_t qph = 0, kph = 0;
    for (int item = (int)((blockIdx.x + gridDim.x - p.cta_shift) % gridDim.x); item < p.items; item += gridDim.x) {
      const int pr = item % p.pairs;
      const int gh = item / p.pairs;
      const int h = gh % p.H, g = gh / p.H;
      uint8_t* qbuf = q_smem + qb * q_bytes;
      // window coordinates (windowed): g = image * nWin + wi * nJ + wj
      int c3 = 0, c4 = 0;
      if (WINDOWED) {
        const int img = g / p.nWin, w = g - img * p.nWin;
        c3 = w % p.nJ;                                  // window column
        c4 = img * (p.nWin / p.nJ) + w / p.nJ;          // image * nI + window row
      }
      if (!WINDOWED && p.wait_ctr != nullptr) {
        // fused kernels: the q/k/v rows of group g are produced by an earlier phase on other SMs
        if (elect_one()) {
          const long long r0 = (long long)g * p.wait_rows_per_group, r1 = r0 + p.n_q - 1;
          for (int t = (int)(r0 >> 7); t <= (int)(r1 >> 7); ++t) wait_counter_ge(p.wait_ctr + t, p.wait_target);
        }
        __syncwarp();
      }
      mbar_wait(&q_empty[qb], qph ^ 1);
      if (elect_one()) {
        if (WINDOWED) {
          // whole window of Q (nkb boxes, window-token order), the MMA picks its 128-row tiles by offset
          mbar_expect_tx(&q_full[qb], p.nkb * p.kb_rows * Cfg::ROW_BYTES);
          for (int j = 0; j < p.nkb; ++j)
            tma_load_5d(qbuf + j * p.kb_rows * Cfg::ROW_BYTES, &tmQ, p.q_col0 + h * HD, 0, c3, j * p.h_box, c4, &q_full[qb]);
        } else {
          const bool two = (2 * pr + 1) < p.q_tiles;
          mbar_expect_tx(&q_full[qb], (two ? 2 : 1) * Cfg::Q_TILE_BYTES);
          tma_load_3d(qbuf, &tmQ, p.q_col0 + h * HD, (2 * pr) * 128, g, &q_full[qb]);
          if (two) tma_load_3d(qbuf + Cfg::Q_TILE_BYTES, &tmQ, p.q_col0 + h * HD, (2 * pr + 1) * 128, g, &q_full[qb]);
        }
      }
      __syncwarp();
      for (int j = 0; j < p.nkb; ++j) {
        uint8_t* kbuf = kv_smem + st * 2 * kvb_bytes;
        mbar_wait(&kv_empty[st], kph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[st], 2 * p.kb_rows * Cfg::ROW_BYTES);
          if (WINDOWED) {
            tma_load_5d(kbuf, &tmK, p.k_col0 + h * HD, 0, c3, j * p.h_box, c4, &kv_full[st]);
            tma_load_5d(kbuf + kvb_bytes, &tmV, p.v_col0 + h * HD, 0, c3, j * p.h_box, c4, &kv_full[st]);
          } else {
            tma_load_3d(kbuf, &tmK, p.k_col0 + h * HD, j * p.kb, g, &kv_full[st]);
            tma_load_3d(kbuf + kvb_bytes, &tmV, p.v_col0 + h * HD, j * p.kb, g, &kv_full[st]);
          }
        }
        __syncwarp();
        if (++st == 2) { st = 0; kph ^= 1; }
      }
      if (++qb == 2) { qb = 0; qph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp converged; one elected lane issues) =====================
    // The two slots run half a period apart: per (item, key block) the issue order is
    //     S(slot 0)  ->  PV(slot 1) of the PREVIOUS step  ->  S(slot 1)  ->  PV(slot 0)
    // so one warpgroup is in its softmax (TMEM reads + MUFU) while the other waits on its MMA / writes its output,
    // instead of both competing for the same TMEM lane-quarter ports and then idling together.
    int qb = 0, st = 0;
    uint32_t qph = 0, kph = 0;
    uint32_t se_ph0 = 0, se_ph1 = 0;     // slot_empty phase (one completion per item and slot)
    uint32_t pf_ph0 = 0, pf_ph1 = 0;     // p_full phase (one completion per block and slot)
    const int ksteps_o = p.kb / 16;
    const int h16 = (ksteps_o + 1) / 2;         // 16-column steps owned by the first thread of each row
    const uint32_t q_base = smem_u32(q_smem), kv_base = smem_u32(kv_smem);
    // pending PV of slot 1
    bool pend = false;
    uint32_t pend_kbuf = 0;
    int pend_st = 0, pend_qb = 0, pend_acc = 0, pend_lastj = 0;
    int iseq = 0;

    auto issue_pv = [&](int s, uint32_t kbuf, int accumulate_blocks) {
      const uint32_t slot = tmem_base + s * ATTN_SLOT_COLS;
      // V block is [key][d] with d contiguous: MN-major B operand; 8-key groups are SBO bytes apart
      const uint64_t vdesc = make_sdesc(kbuf + kvb_bytes, Cfg::SBO, Cfg::SBO, Cfg::SWZ);
      // P of keys [0, 16*h16) sits at columns 8k; P of the second half of the keys sits inside ITS thread's S columns
      for (int k = 0; k < ksteps_o; ++k) {
        const int pcol = (k < h16) ? 8 * k : 16 * h16 + 8 * (k - h16);
        umma_ts(slot + Cfg::O_COL, slot + pcol, vdesc + Cfg::V_KSTEP * k, p.idesc_o, (accumulate_blocks | k) != 0);
      }
      umma_commit(&o_full[s]);
    };
    // S = Q K^T of a slot is issued in two parts.  The columns below the O accumulator ([0, 256-HD)) hold nothing the
    // previous item still needs once its PV MMA has been ISSUED (the tensor pipe executes in issue order, and the softmax
    // warps finished with S/P before p_full), so that part goes out BEFORE the wait for the slot hand-back and runs under
    // the previous item's O read-out; only the columns that overlap O (16 of ViT's 208) wait for slot_empty.
    const int s_n1 = p.kb < Cfg::O_COL ? p.kb : Cfg::O_COL;
    const uint32_t idesc_s1 = (p.idesc_s & ~(0x3Fu << 17)) | ((uint32_t)(s_n1 >> 3) << 17);
    const uint32_t idesc_s2 = (p.idesc_s & ~(0x3Fu << 17)) | ((uint32_t)((p.kb - s_n1) >> 3) << 17);
    const uint64_t k_off2 = (uint64_t)((s_n1 * Cfg::ROW_BYTES) >> 4);        // K rows s_n1.. (whole swizzle atoms)
    auto issue_s_early = [&](uint32_t d_tmem, uint64_t qdesc, uint64_t kdesc) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(d_tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_s1, k != 0);
      }
      __syncwarp();
    };
    auto issue_s_late = [&](uint32_t d_tmem, uint64_t qdesc, uint64_t kdesc, uint64_t* bar) {
      if (elect_one()) {
        if (p.kb > s_n1) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_ss(d_tmem + s_n1, qdesc + 2 * k, kdesc + k_off2 + 2 * k, idesc_s2, k != 0);
        }
        umma_commit(bar);
      }
      __syncwarp();
    };
    auto flush_pending = [&]() {
      if (!pend) return;
      mbar_wait(&p_full[1], pf_ph1);
      pf_ph1 ^= 1;
      tc_fence_after();
      if (elect_one()) {
        issue_pv(1, pend_kbuf, pend_acc);
        umma_commit(&kv_empty[pend_st]);            // slot 1's PV is the last reader of this K/V stage
        if (pend_lastj) umma_commit(&q_empty[pend_qb]);
      }
      __syncwarp();
      if (lane == 0) ATTN_TRACE(iseq, 4);
      pend = false;
    };

    for (int item = (int)((blockIdx.x + gridDim.x - p.cta_shift) % gridDim.x); item < p.items; item += gridDim.x, ++iseq) {
      const int pr = item % p.pairs;
      const int nslots = ((2 * pr + 1) < p.q_tiles) ? 2 : 1;
      const uint32_t qbuf = q_base + qb * q_bytes;
      mbar_wait(&q_full[qb], qph);
      tc_fence_after();
      if (lane == 0) ATTN_TRACE(iseq, 0);
      for (int j = 0; j < p.nkb; ++j) {
        const uint32_t kbuf = kv_base + st * 2 * kvb_bytes;
        mbar_wait(&kv_full[st], kph);
        tc_fence_after();
        const uint64_t kdesc = make_sdesc(kbuf, 16, Cfg::SBO, Cfg::SWZ);
        // ---- S of slot 0
        {
          const int tile = WINDOWED ? (2 * pr) : 0;
          const uint64_t qdesc = make_sdesc(qbuf + tile * Cfg::Q_TILE_BYTES, 16, Cfg::SBO, Cfg::SWZ);
          issue_s_early(tmem_base, qdesc, kdesc);
          if (j == 0) {
            mbar_wait(&slot_empty[0], se_ph0 ^ 1);
            se_ph0 ^= 1;
            tc_fence_after();
          }
          issue_s_late(tmem_base, qdesc, kdesc, &s_full[0]);
          if (lane == 0) ATTN_TRACE(iseq, 1);
        }
        // ---- PV of slot 1 from the previous step
        flush_pending();
        // ---- S of slot 1
        if (nslots == 2) {
          const int tile = WINDOWED ? (2 * pr + 1) : 1;
          const uint64_t qdesc = make_sdesc(qbuf + tile * Cfg::Q_TILE_BYTES, 16, Cfg::SBO, Cfg::SWZ);
          issue_s_early(tmem_base + ATTN_SLOT_COLS, qdesc, kdesc);
          if (j == 0) {
            mbar_wait(&slot_empty[1], se_ph1 ^ 1);
            se_ph1 ^= 1;
            tc_fence_after();
          }
          issue_s_late(tmem_base + ATTN_SLOT_COLS, qdesc, kdesc, &s_full[1]);
          if (lane == 0) ATTN_TRACE(iseq, 2);
        }
        // ---- PV of slot 0
        {
          mbar_wait(&p_full[0], pf_ph0);
          pf_ph0 ^= 1;
          tc_fence_after();
          if (elect_one()) {
            issue_pv(0, kbuf, j);
            if (nslots == 1) {                       // no slot-1 work: slot 0's PV is the last reader
              umma_commit(&kv_empty[st]);
              if (j == p.nkb - 1) umma_commit(&q_empty[qb]);
            }
          }
          __syncwarp();
          if (lane == 0) ATTN_TRACE(iseq, 3);
        }
        if (nslots == 2) {
          pend = true; pend_kbuf = kbuf; pend_st = st; pend_qb = qb; pend_acc = j; pend_lastj = (j == p.nkb - 1);
        }
        if (++st == 2) { st = 0; kph ^= 1; }
      }
      if (++qb == 2) { qb = 0; qph ^= 1; }
    }
    flush_pending();
  } else if (warp >= 2 && warp < 18) {
    // ===================== softmax + epilogue: 8 warps per slot, two threads per query row =====================
    const int sw = warp - 2;
    const int slot = sw >> 3;
    const int hf = (sw >> 2) & 1;             // which half of the row's columns this thread owns
    const int q = warp & 3;                   // TMEM lane quarter of this warp
    const int trow = q * 32 + lane;           // row inside the 128-row tile
    const uint32_t t_slot = tmem_base + slot * ATTN_SLOT_COLS + ((uint32_t)(q * 32) << 16);
    const int n16 = p.kb >> 4;
    const int h16 = (n16 + 1) / 2;
    const int c_lo = hf ? h16 * 16 : 0;       // first S column of this thread
    const int nst = hf ? n16 - h16 : h16;     // its number of 16-column steps
    const uint32_t t_my = t_slot + c_lo;      // its S columns; its P goes in place at t_my + 8k
    const float sl2 = p.scale_log2e;
    float* xch = reinterpret_cast<float*>(o_smem + attn_ostage_bytes(HD, p.tma_store != 0) - 4096);
    float* xmax = xch + (slot * 2) * 128;     // [2 halves][128 rows]
    float* xsum = xch + 512 + (slot * 2) * 128;
    constexpr int OH = HD / 2;                // O columns per thread
    const uint32_t t_o = t_slot + Cfg::O_COL + hf * OH;
    uint32_t sf_ph = 0, of_ph = 0;
    int iseq = -1;
    const bool tracer = (q == 0 && hf == 0 && lane == 0);
    pend_sig = -1;
    for (int item = (int)((blockIdx.x + gridDim.x - p.cta_shift) % gridDim.x); item < p.items; item += gridDim.x) {
      ++iseq;
      const int pr = item % p.pairs;
      const int gh = item / p.pairs;
      const int h = gh % p.H, g = gh / p.H;
      const int qt = 2 * pr + slot;
      if (qt >= p.q_tiles) continue;          // slot idle for this item (uniform over the slot's warps)
      const int row = qt * 128 + trow;
      const bool warp_active = (qt * 128 + q * 32) < p.n_q;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < p.nkb; ++j) {
        const int nvalid = min(p.kb_rows, p.n_k - j * p.kb_rows) - c_lo;    // real keys among this thread's columns
        mbar_wait(&s_full[slot], sf_ph);
        sf_ph ^= 1;
        tc_fence_after();
        if (tracer && j == 0) ATTN_TRACE(iseq, 5 + 5 * slot);
        // ---- pass 1: partial row max over this thread's columns
        float mx = -INFINITY;
        if (warp_active) {
          int k = (p.debug_flags & 4) ? nst : 0;
#pragma unroll 1
          for (; k + 1 < nst; k += 2) {        // 32 columns per TMEM round trip (only the max is kept: registers are free here)
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
        asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
        mx = fmaxf(mx, xmax[(hf ^ 1) * 128 + trow]);
        if (tracer && j == 0) ATTN_TRACE(iseq, 6 + 5 * slot);
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2f((m_run - m_new) * sl2);      // 0 on the first block (m_run = -inf)
        if (j > 0) {
          // previous block's PV must have retired before O is rescaled / P overwritten
          mbar_wait(&o_full[slot], of_ph);
          of_ph ^= 1;
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
          // ---- pass 2: p = exp2((s - m) * scale*log2e) over this thread's columns; fp16 P in place; partial row sum
          const float mxs = m_new * sl2;
          uint32_t va[16], pk[8];
          float e[16];
          float s0 = 0.f, s1 = 0.f;
          // Every attempt to have the next step's TMEM load in flight while this step computes was slower on B200: a second
          // register buffer (39 -> 52 us, twice) and also re-using `va` as soon as the exponent arguments are formed, with
          // no extra registers (35 -> 47 us).  A tcgen05.ld overlapping this warp's tcgen05.st / MUFU stream loses more than
          // the latency it hides; the other three softmax warps of the scheduler are what covers it.
#pragma unroll 1
          for (int k = 0; k < nst; ++k) {
            tmem_ld16(t_my + k * 16, va);
            tmem_ld_wait();
            exp_stage(va, e, nvalid - k * 16, sl2, mxs, p.debug_flags);
            pack_stage(e, pk, s0, s1);
            tmem_st8(t_my + k * 8, pk);
          }
          tmem_st_wait();
          l_run = l_run * alpha + (s0 + s1);
          m_run = m_new;
        }
        if (j == p.nkb - 1) xsum[hf * 128 + trow] = l_run;     // partial row sum for the partner (ordered by p_full -> o_full)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[slot]);
        if (tracer && j == 0) ATTN_TRACE(iseq, 7 + 5 * slot);
      }

      // ---- epilogue: this thread's half of the O columns / rowsum (+ LePE already in place) -> fp16 -> global
      mbar_wait(&o_full[slot], of_ph);
      of_ph ^= 1;
      tc_fence_after();
      if (tracer) ATTN_TRACE(iseq, 8 + 5 * slot);
      uint32_t v[OH];
      if (warp_active) {
        if (OH == 32) tmem_ld32(t_o, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        else tmem_ld16(t_o, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld_wait();
      }
      // O is in registers: hand the TMEM slot back NOW so the next S MMA overlaps the normalise + store below
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_empty[slot]);
      if (!WINDOWED && p.tma_store) {
        // the staging tile is free once the PREVIOUS item's bulk store has read it (long ago: off the critical path)
        if (sw == slot * 8 && lane == 0) {
          if (p.signal_ctr != nullptr && pend_sig >= 0) {     // previous tile of this slot: stores complete -> publish
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __threadfence();
            atomicAdd(p.signal_ctr + pend_sig, 1);
          } else {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          pend_sig = g;
        }
        asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
      }
      if (warp_active) {
        const float inv = 1.f / (l_run + xsum[(hf ^ 1) * 128 + trow]);
        if (!WINDOWED && p.tma_store) {
          // ---- staged path: rows -> swizzled smem tile -> one TMA store per slot (clips rows >= n_q)
          uint8_t* rowp = o_smem + slot * (128 * HD * 2) + trow * (HD * 2);
#pragma unroll
          for (int i = 0; i < OH / 8; ++i) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * i + k]) * inv;
            const int ch = hf * (OH / 8) + i;                                              // 16-byte chunk of the row
            const int swz = (HD == 64) ? (ch ^ (trow & 7)) : (ch ^ ((trow >> 1) & 3));     // SW128 / SW64 chunk swizzle
            *reinterpret_cast<uint4*>(rowp + (swz << 4)) =
                make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
          }
          fence_proxy_async_smem();
        } else if (row < p.n_q) {
          long long tok;
          int grp;
          if (WINDOWED) {
            const int img = g / p.nWin, w = g - img * p.nWin;
            const int wi = w / p.nJ, wj = w - wi * p.nJ;
            const int r = row / p.W_sp, c = row - r * p.W_sp;
            tok = (long long)(wi * p.H_sp + r) * p.R + wj * p.W_sp + c;
            grp = img;
          } else {
            tok = row;
            grp = g;
          }
          uint16_t* dst = reinterpret_cast<uint16_t*>(p.O) + (long long)grp * p.o_group + tok * p.ldo + p.o_col0 + h * HD + hf * OH;
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int i = 0; i < OH / 8; ++i) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * i + k]) * inv;
            if (p.add_into_out) {
              const uint4 old = d4[i];
              const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                f[2 * k] += __low2float(oh[k]);
                f[2 * k + 1] += __high2float(oh[k]);
              }
            }
            d4[i] = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
          }
        }
      }
      if (!WINDOWED && p.tma_store) {
        // all eight warps of the slot (active or not) meet, then one thread issues the bulk store of the tile
        asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
        if (sw == slot * 8 && lane == 0) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmO)),
                       "r"(smem_u32(o_smem + slot * (128 * HD * 2))), "r"(p.o_col0 + h * HD), "r"(qt * 128), "r"(g)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (tracer) ATTN_TRACE(iseq, 9 + 5 * slot);
    }
  }

  if (!WINDOWED && p.tma_store && (warp == 2 || warp == 10) && lane == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.signal_ctr != nullptr && pend_sig >= 0) { __threadfence(); atomicAdd(p.signal_ctr + pend_sig, 1); }
  }
}

template <int HD, bool WINDOWED>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_core_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const AttnParams p) {
  using Cfg = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_rows = WINDOWED ? attn_q_rows(true, p.nkb, p.kb_rows) : 256;
  const int data_bytes = 2 * q_rows * Cfg::ROW_BYTES + 4 * p.kb * Cfg::ROW_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + data_bytes + attn_ostage_bytes(HD, p.tma_store != 0));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (WINDOWED) {
    // rows never touched by TMA (tile padding beyond the window) must read as finite zeros
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < data_bytes / 16; i += ATTN_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    attn_init_barriers(bars);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  attn_run<HD, WINDOWED>(tmQ, tmK, tmV, tmO, p, smem, bars, tmem_base);

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace pa
