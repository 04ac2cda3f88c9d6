// pa_attn_proj.cuh — attention core AND output projection of the spatial-reduction family (pvt.py:84-89, segformer.py:44-48)
// in one kernel:   y[128-row tile] = concat_h( softmax(Q_h K_h^T scale) V_h ) Wp^T + b
//
// Why: with <= 64 keys per image the PVT attention is HBM-bound as separate launches -- the attention core writes O
// [B*N, C] and the projection GEMM reads it back (at BASELINE config 3: 2 x 134 MB of 940 MB moved per forward, 69 + 62 us).
// Here O never leaves the SM, and it never even reaches shared memory:
//   * per head, O_h = P_h V_h is read from TMEM, normalised, rounded to fp16 and written back to TMEM as 32 packed columns;
//   * the eight heads side by side are a 128 x C fp16 matrix in TMEM -- exactly the A operand of a TS MMA (the same operand
//     form P takes in P V), so the projection is  Y[:, n-chunk] = O_all(TMEM) x Wp[n-chunk, :]^T  with only Wp streamed
//     through shared memory.
//
// One CTA per SM, 640 threads, CTAs in clusters of two; unit = (image, 128-row query tile), all heads:
//   warp 0       TMA producer: per head Q_h tile, K_h, V_h (2-stage rings, the first two heads of the NEXT unit are requested
//                before this unit's projection tiles); the Wp tiles [128 n x 64 k] (4-stage ring) -- each CTA of the pair
//                fetches HALF of every tile and multicasts it to both (first version: every CTA streamed all of Wp per
//                128 rows, 8.4 TB/s of L2 reads at config 3, and the projection phase ran at 3.5 k instead of 2.0 k cycles
//                per 128-column chunk)
//   warp 1       MMA issuer: S(h) -> slot h&1, P V(h) (whichever is ready first), then per 128-column chunk of y the K = C chain
//   warp 2       TMEM allocator (512 columns)
//   warps 4-19   two softmax sets of 8 warps (two threads per query row); set s owns attention slot s: scores are read from
//                TMEM ONCE (<= 32 columns per thread stay in registers between the max and the exp pass), O read-out into the
//                packed O_all columns; afterwards all 16 warps are the projection epilogue (bias, 16-bit, per-warp staging
//                tile, bulk store).  (first version: one set of 8 warps walking the heads in turn -- 2.3 k cycles per head)
// TMEM:  [0, 256)   O_all: head h at columns [32h, 32h + 32), fp16 pairs
//        [256, 512) two 128-column regions: attention slot s = {S fp32 [0, kb) -> P fp16 in place, O_h fp32 [64, 128)}
//                   during the head loop, accumulator of y chunk j (j & 1) during the projection
// Limits: 64-wide heads, at most 64 keys (one 64-column S tile), 128 <= C = heads x 64 <= 512, 16-bit y.
//
// STATUS: opt-in (PA_PVT_FUSED=1), parity-tested, NOT the default -- measured at BASELINE config 3 (B=32, 64x64 tokens, C=512,
// 8 heads, 64 keys; profiles/attn_proj_r02.txt):   two launches (attention 69 us + proj GEMM 62 us) 251 us per forward,
// this kernel 248 us in its first form (one softmax set, every CTA streaming all of Wp), 289-298 us in this form.
// Clock stamps per 128-row unit: head loop 18-25 k cycles, projection 14-16 k (ideal: 4 k MUFU-bound / 8.2 k tensor-bound).
// What the stamps show:
//   * the projection's MMAs EXECUTE at ~170 cycles each (64 at full rate) although they are issued far ahead: a TS MMA reads
//     its A operand from TMEM for every N = 128 chunk -- 4 x 128 KB per unit -- while the epilogue warps drain 256 KB of fp32
//     accumulators from the same TMEM, and the head loop reads S and O (2 x 256 KB) the same way: the TMEM read path
//     (tens of bytes per clock) is the bottleneck of BOTH phases, not the tensor pipe, MUFU or L2;
//   * at 128-row tiles the operand streams are large: per unit 128 KB of Q, 128 KB of K/V and 512 KB of Wp pass through shared
//     memory (the stand-alone GEMM amortises Wp over 256 rows), so the rings must hold latency x rate = ~128 KB in flight.
// A successor needs the accumulators to leave TMEM less often: y chunks of N = 256 (two A passes instead of four) and O_all as
// an SS operand for CTA pairs, or the whole unit on 256 rows (cta_group::2).
#pragma once
#include "pa_attn.cuh"

namespace pa {

constexpr int AP_THREADS = 640;
constexpr int AP_W_STAGES = 7;      // 7 x 16 KB in flight: at 256 tensor cycles per tile a 4-deep ring covered half of the L2 latency
constexpr int AP_Q_BYTES = 16384;                 // 128 rows x 128 B
constexpr int AP_KV_BYTES = 8192;                 // 64 rows x 128 B
constexpr int AP_W_BYTES = 16384;                 // 128 output channels x 64 k
constexpr int AP_OFF_Q = 0;                                         // [2]
constexpr int AP_OFF_K = AP_OFF_Q + 2 * AP_Q_BYTES;                 // [2]
constexpr int AP_OFF_V = AP_OFF_K + 2 * AP_KV_BYTES;                // [2]
constexpr int AP_OFF_W = AP_OFF_V + 2 * AP_KV_BYTES;                // [AP_W_STAGES]
constexpr int AP_OFF_Y = AP_OFF_W + AP_W_STAGES * AP_W_BYTES;       // 16 staging tiles of 32 x 32 16-bit (64-byte swizzle)
constexpr int AP_OFF_X = AP_OFF_Y + 16 * 2048;                      // xmax[2 sets][2][128], xsum[2 sets][2][128] fp32
constexpr int AP_OFF_BAR = AP_OFF_X + 4096;
constexpr int AP_SMEM_BYTES = AP_OFF_BAR + 512 + 1024;              // + barriers + 1024-alignment slack

struct AttnProjParams {
  int G, H, n_q, n_k, kb;          // groups (images), heads, rows per group, keys per group, S tile width (multiple of 16, <= 64)
  int q_tiles, units;
  int C;                           // H * 64 = width of y and K of the projection
  int q_col0, k_col0, v_col0;      // element column of head 0 inside the Q / K / V tensor maps
  float scale_log2e;
  const float* bias;               // [C] fp32 or nullptr
  int out_dtype;                   // 0 fp16, 1 bf16
  uint32_t idesc_s, idesc_o, idesc_y;
  long long* trace;                // debug: clock64 stamps of CTA 0, [unit][16], or nullptr
};
#define AP_TRACE(seq, slot_) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && (seq) < 32) p.trace[(seq) * 16 + (slot_)] = clock64(); } while (0)

__global__ void __launch_bounds__(AP_THREADS, 1)
attn_proj_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmY, const AttnProjParams p) {
  using Cfg = AttnCfg<64>;
  extern __shared__ uint8_t ap_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ap_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AP_OFF_BAR);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* k_full = bars + 4;        // [2]
  uint64_t* k_empty = bars + 6;       // [2]
  uint64_t* v_full = bars + 8;        // [2]
  uint64_t* v_empty = bars + 10;      // [2]
  uint64_t* s_full = bars + 12;       // [2] per slot
  uint64_t* p_full = bars + 14;       // [2] 8 warps of the slot's softmax set
  uint64_t* o_full = bars + 16;       // [2]
  uint64_t* slot_free = bars + 18;    // [2] 8 warps: O_h read out, packed O stored
  uint64_t* oall_full = bars + 20;    // 8 * H arrivals per unit: every head's packed O is in TMEM
  uint64_t* w_full = bars + 21;       // [AP_W_STAGES] both halves of the tile have landed (one from each CTA of the pair)
  uint64_t* w_empty = bars + 21 + AP_W_STAGES;          // [AP_W_STAGES] 2 arrivals: both CTAs' MMAs have read the stage
  uint64_t* y_full = bars + 21 + 2 * AP_W_STAGES;       // [2]
  uint64_t* y_empty = y_full + 2;                       // [2] 16 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_empty + 2);
  float* xmax = reinterpret_cast<float*>(smem + AP_OFF_X);          // [set][half][128]
  float* xsum = xmax + 512;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = p.H;
  const int n_chunks = (p.C + 127) / 128;
  const int crank = (int)cluster_ctarank();
  // units are dealt to the CTA pairs two at a time so that both CTAs of a pair run the same number of rounds (they share the
  // Wp ring); a pair's last round may hold a phantom unit (u >= units): it runs on zero-filled tiles and stores nothing
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int rounds = (p.units + 2 * npairs - 1) / (2 * npairs);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 14; ++i) mbar_init(&bars[i], 1);        // q/k/v full+empty, s_full
    mbar_init(&p_full[0], 8); mbar_init(&p_full[1], 8);
    mbar_init(&o_full[0], 1); mbar_init(&o_full[1], 1);
    mbar_init(&slot_free[0], 8); mbar_init(&slot_free[1], 8);
    mbar_init(oall_full, 8 * H);
    for (int i = 0; i < AP_W_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 2); }
    mbar_init(&y_full[0], 1); mbar_init(&y_full[1], 1);
    mbar_init(&y_empty[0], 16); mbar_init(&y_empty[1], 16);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_oall = tmem_base;
  const uint32_t t_reg = tmem_base + 256;       // region s at t_reg + 128 s

  if (warp == 0) {
    // ===================== TMA producer =====================
    int hc = 0, wc = 0;
    auto load_head = [&](int u, int h) {
      const int qt = u % p.q_tiles, g = u / p.q_tiles;
      const int st = hc & 1;
      const uint32_t ph = (hc >> 1) & 1;
      mbar_wait(&q_empty[st], ph ^ 1);
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[st], AP_Q_BYTES);
        tma_load_3d(smem + AP_OFF_Q + st * AP_Q_BYTES, &tmQ, p.q_col0 + h * 64, qt * 128, g, &q_full[st]);
        mbar_expect_tx(&k_full[st], p.kb * 128);
        tma_load_3d(smem + AP_OFF_K + st * AP_KV_BYTES, &tmK, p.k_col0 + h * 64, 0, g, &k_full[st]);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&v_full[st], p.kb * 128);
        tma_load_3d(smem + AP_OFF_V + st * AP_KV_BYTES, &tmV, p.v_col0 + h * 64, 0, g, &v_full[st]);
      }
      __syncwarp();
      ++hc;
    };
    auto load_w = [&](int t) {            // tile t of the unit: chunk t / H, k-block t % H; this CTA's 64 of its 128 rows
      const int j = t / H, kbk = t - j * H;
      const int st = wc % AP_W_STAGES;
      const uint32_t ph = (wc / AP_W_STAGES) & 1;
      mbar_wait(&w_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&w_full[st], AP_W_BYTES);
        tma_load_3d_mc(smem + AP_OFF_W + st * AP_W_BYTES + crank * (AP_W_BYTES / 2), &tmW, kbk * 64, j * 128 + crank * 64, 0,
                       &w_full[st], 3);
      }
      __syncwarp();
      ++wc;
    };
    const int pre = H < 2 ? H : 2;
    const int wtiles = n_chunks * H;
    const int wpre = wtiles < AP_W_STAGES ? wtiles : AP_W_STAGES;
    if (rounds > 0) for (int h = 0; h < pre; ++h) load_head(pair * 2 + crank, h);
    for (int r = 0; r < rounds; ++r) {
      const int u = (r * npairs + pair) * 2 + crank;
      for (int h = pre; h < H; ++h) load_head(u, h);
      for (int t = 0; t < wpre; ++t) load_w(t);
      if (r + 1 < rounds) for (int h = 0; h < pre; ++h) load_head(((r + 1) * npairs + pair) * 2 + crank, h);
      for (int t = wpre; t < wtiles; ++t) load_w(t);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const int ksteps_o = p.kb / 16;
    const int h16 = (ksteps_o + 1) / 2;
    const uint32_t q_base = smem_u32(smem + AP_OFF_Q), k_base = smem_u32(smem + AP_OFF_K), v_base = smem_u32(smem + AP_OFF_V);
    const uint32_t w_base = smem_u32(smem + AP_OFF_W);
    int hc = 0, wc = 0, yc = 0;
    int su0 = 0, su1 = 0, yu0 = 0, yu1 = 0;  // uses of region 0 / 1 as an attention slot / as a y accumulator so far
    for (int r = 0; r < rounds; ++r) {
      const int hc0 = hc;
      AP_TRACE(r, 0);
      // ---- head loop: S(h) and P V(h) are issued in whatever order their inputs become ready
      int ns = 0, npv = 0;
      while (npv < H) {
        if (ns < H) {
          const int st = hc & 1;
          const uint32_t ph = (hc >> 1) & 1;
          const int su = st ? su1 : su0, yu = st ? yu1 : yu0;
          bool ok = mbar_test_wait(&q_full[st], ph) && mbar_test_wait(&k_full[st], ph);
          if (ok && su > 0) ok = mbar_test_wait(&slot_free[st], (su - 1) & 1);    // previous head in this region read out
          if (ok && yu > 0) ok = mbar_test_wait(&y_empty[st], (yu - 1) & 1);      // previous y chunk in this region drained
          if (ok) {
            tc_fence_after();
            const uint64_t qdesc = make_sdesc(q_base + st * AP_Q_BYTES, 16, Cfg::SBO, Cfg::SWZ);
            const uint64_t kdesc = make_sdesc(k_base + st * AP_KV_BYTES, 16, Cfg::SBO, Cfg::SWZ);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_ss(t_reg + 128 * st, qdesc + 2 * k, kdesc + 2 * k, p.idesc_s, k != 0);
              umma_commit(&s_full[st]);
              umma_commit(&q_empty[st]);
              umma_commit(&k_empty[st]);
            }
            __syncwarp();
            if (st) ++su1; else ++su0;
            ++hc; ++ns;
          }
        }
        if (npv < ns) {
          const int hp = hc0 + npv;
          const int st = hp & 1;
          const uint32_t ph = (hp >> 1) & 1;
          if (mbar_test_wait(&p_full[st], ph) && mbar_test_wait(&v_full[st], ph)) {
            tc_fence_after();
            const uint64_t vdesc = make_sdesc(v_base + st * AP_KV_BYTES, Cfg::SBO, Cfg::SBO, Cfg::SWZ);   // V [key][d]: MN-major B
            if (elect_one()) {
              for (int k = 0; k < ksteps_o; ++k) {
                const int pcol = (k < h16) ? 8 * k : 16 * h16 + 8 * (k - h16);
                umma_ts(t_reg + 128 * st + 64, t_reg + 128 * st + pcol, vdesc + Cfg::V_KSTEP * k, p.idesc_o, k != 0);
              }
              umma_commit(&o_full[st]);
              umma_commit(&v_empty[st]);
            }
            __syncwarp();
            ++npv;
          }
        }
      }
      AP_TRACE(r, 1);
      // ---- projection: y chunk j = O_all (TMEM, fp16) x Wp[128 j .. 128 j + 127, :]^T
      mbar_wait(oall_full, r & 1);
      tc_fence_after();
      AP_TRACE(r, 2);
      for (int j = 0; j < n_chunks; ++j, ++yc) {
        const int rg = yc & 1;
        { const int su = rg ? su1 : su0, yu = rg ? yu1 : yu0;
          if (su > 0) mbar_wait(&slot_free[rg], (su - 1) & 1);
          if (yu > 0) mbar_wait(&y_empty[rg], (yu - 1) & 1); }
        tc_fence_after();
        for (int kbk = 0; kbk < H; ++kbk, ++wc) {
          const int st = wc % AP_W_STAGES;
          const uint32_t ph = (wc / AP_W_STAGES) & 1;
          mbar_wait(&w_full[st], ph);
          tc_fence_after();
          const uint64_t wdesc = make_sdesc(w_base + st * AP_W_BYTES, 16, 1024, PA_SWZ_128B);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ts(t_reg + 128 * rg, t_oall + 32 * kbk + 8 * k, wdesc + 2 * k, p.idesc_y, (kbk | k) != 0);
            umma_commit_mc(&w_empty[st], 3);          // the stage is refilled by BOTH CTAs' producers
            if (kbk == H - 1) umma_commit(&y_full[rg]);
          }
          __syncwarp();
        }
        if (rg) ++yu1; else ++yu0;
        AP_TRACE(r, 3 + (j & 3));
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax sets, O read-out, projection epilogue =====================
    const int set = (warp - 4) >> 3;              // owns attention slot `set`
    const int hf = ((warp - 4) >> 2) & 1;
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int n16 = p.kb >> 4;
    const int h16 = (n16 + 1) / 2;
    const int c_lo = hf ? h16 * 16 : 0;
    const int nst = hf ? n16 - h16 : h16;         // 16-column steps of this thread: 0, 1 or 2
    const int nvalid = min(p.kb, p.n_k) - c_lo;
    const float sl2 = p.scale_log2e;
    uint8_t* wbuf = smem + AP_OFF_Y + (warp - 4) * 2048;
    const uint32_t t_slot = t_reg + 128 * set + lane_off;
    const uint32_t t_my = t_slot + c_lo;
    float* xm = xmax + set * 256;
    float* xs = xsum + set * 256;
    const int cq = set * 2 + hf;                  // epilogue: this warp's 32-column quarter of every 128-column chunk
    const bool tracer = (warp == 4);
    int hc = 0, yc = 0;
    for (int r = 0; r < rounds; ++r) {
      const int u = (r * npairs + pair) * 2 + crank;
      const bool real = u < p.units;
      const int qt = u % p.q_tiles, g = u / p.q_tiles;
      const bool warp_active = (qt * 128 + q * 32) < p.n_q;
      // heads whose running index has this set's parity
      for (int h = 0; h < H; ++h, ++hc) {
        if ((hc & 1) != set) continue;
        const uint32_t ph = (hc >> 1) & 1;
        mbar_wait(&s_full[set], ph);
        tc_fence_after();
        if (tracer && h < 2) AP_TRACE(r, 8);
        // ---- scores: one TMEM read, kept in registers for both passes
        uint32_t v0[16], v1[16];
        float mx = -INFINITY;
        if (warp_active) {
          if (nst > 0) tmem_ld16(t_my, v0);
          if (nst > 1) tmem_ld16(t_my + 16, v1);
          tmem_ld_wait();
          if (nst > 0) mx = chunk_max<16>(v0, nvalid, mx);
          if (nst > 1) mx = chunk_max<16>(v1, nvalid - 16, mx);
        }
        xm[hf * 128 + trow] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + set) : "memory");
        mx = fmaxf(mx, xm[(hf ^ 1) * 128 + trow]);
        float l_run = 0.f;
        if (warp_active) {
          const float mxs = mx * sl2;
          uint32_t pk[8];
          float e[16];
          float s0 = 0.f, s1 = 0.f;
          if (nst > 0) { exp_stage(v0, e, nvalid, sl2, mxs); pack_stage(e, pk, s0, s1); tmem_st8(t_my, pk); }
          if (nst > 1) { exp_stage(v1, e, nvalid - 16, sl2, mxs); pack_stage(e, pk, s0, s1); tmem_st8(t_my + 8, pk); }
          tmem_st_wait();
          l_run = s0 + s1;
        }
        xs[hf * 128 + trow] = l_run;               // ordered towards the partner by p_full -> o_full
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[set]);
        // ---- read-out: O_h / rowsum -> fp16 -> packed columns [32 h + 16 hf, +16) of O_all
        mbar_wait(&o_full[set], ph);
        tc_fence_after();
        if (warp_active) {
          uint32_t v[32], pk[16];
          tmem_ld32(t_slot + 64 + hf * 32, v);
          tmem_ld_wait();
          const float inv = 1.f / (l_run + xs[(hf ^ 1) * 128 + trow]);
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = pack_h2(__uint_as_float(v[2 * k]) * inv, __uint_as_float(v[2 * k + 1]) * inv);
          tmem_st16(t_oall + lane_off + 32 * h + 16 * hf, pk);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&slot_free[set]); mbar_arrive(oall_full); }
      }
      if (tracer) AP_TRACE(r, 9);
      // ---- projection epilogue: this warp drains rows 32q..32q+31, columns [32 cq, 32 cq + 32) of every 128-column chunk
      const int row0 = qt * 128 + q * 32;
      for (int j = 0; j < n_chunks; ++j, ++yc) {
        const int rg = yc & 1;
        const uint32_t ph = (yc >> 1) & 1;
        const int col = j * 128 + cq * 32;
        float4 bq[8];
        if (p.bias != nullptr && col < p.C) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(p.bias + col) + i);
        }
        mbar_wait(&y_full[rg], ph);
        tc_fence_after();
        if (tracer) AP_TRACE(r, 10 + (j & 3));
        uint32_t v[32];
        tmem_ld32(t_reg + 128 * rg + lane_off + cq * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[rg]);
        if (real && col < p.C && row0 < p.n_q) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + bq[i].x);
              v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + bq[i].y);
              v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + bq[i].z);
              v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + bq[i].w);
            }
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile read by the previous store
          __syncwarp();
          uint8_t* rowp = wbuf + lane * 64;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * ch + k]);
            *reinterpret_cast<uint4*>(rowp + ((ch ^ ((lane >> 1) & 3)) << 4)) =
                p.out_dtype == 0 ? make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]))
                                 : make_uint4(pack_bf2(f[0], f[1]), pack_bf2(f[2], f[3]), pack_bf2(f[4], f[5]), pack_bf2(f[6], f[7]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmY)),
                         "r"(smem_u32(wbuf)), "r"(col), "r"(row0), "r"(g)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
#undef AP_TRACE

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still multicast into this CTA's ring / commit to its barriers
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace pa
