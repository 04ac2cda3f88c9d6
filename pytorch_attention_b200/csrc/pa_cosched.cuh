// pa_cosched.cuh — the whole ViT attention forward (ViT.py:79-89) as ONE launch in which the projection GEMMs run UNDER the
// softmax chain on every SM, instead of before / after it (pa_fused.cuh runs the three phases back to back per SM).
//
// Why two CTAs per SM: the attention core needs 256 TMEM columns per 128-row query tile (S fp32 208 columns + O) and the
// GEMM needs 256 for a 256 x 256 pair-tile accumulator, so the SM's 512 columns are split 256 / 256 between two
// co-resident CTAs with different roles, each with <= 113 KB of shared memory and 384 threads:
//
//   role G (one CTA pair per TPC, tcgen05 cta_group::2)      role A (one CTA per SM, cta_group::1)
//     warp 0     TMA producer, 3-stage ring of 64-wide         warp 0     TMA producer: Q tile (double-buffered), K, V
//                k-blocks (A 128 rows + half of B per CTA)      warp 1     MMA issuer: S = Q K^T, O = P V (P from TMEM)
//     warp 1     MMA issuer (leader CTA), M=256 N=256          warp 2     output store + publication (bulk store, counter)
//     warp 2     TMEM allocator                                  warps 4-11 softmax, two threads per query row
//     warps 4-11 epilogue: each warp drains its own 32 rows
//                through a private 2 KB staging tile + its
//                own bulk stores (no block-level barrier)
//
// The tensor pipe of an SM is shared by the two CTAs: while role A sits in its latency chain (TMEM loads, MUFU), the pipe
// executes role G's k-blocks; role G's single accumulator is drained while the pipe serves role A's S / PV MMAs.
// Work streams: role G walks  [qkv tiles ..., proj tiles ...]  (m-group-major, so the first images complete first), role A
// walks (image, head, query tile) units.  They are chained by the same global dependency counters as the sequenced kernel:
//     ctr_qkv[128-row tile]  += 1 per epilogue warp of every stored qkv tile   -> role A waits for the tiles covering its image
//     ctr_attn[image]        += 1 per stored (head, query tile)                -> a proj tile waits for the images it covers
// Every role-G worker finishes all its qkv tiles before its first proj tile and qkv tiles wait for nothing, so the scheme
// cannot deadlock provided every CTA of the grid is resident (2 per SM) -- the host checks cudaOccupancyMaxActiveClusters
// before choosing this kernel and otherwise uses the sequenced kernel / three launches.
// Roles are elected at run time (the hardware's CTA placement is not specified): the first cluster to arrive on a TPC takes
// role G, the second role A; a ticket per role gives the worker index.
#pragma once
#include "pa_attn.cuh"
#include "pa_gemm.cuh"

namespace pa {

constexpr int CS_THREADS = 384;
constexpr int CS_BN = 256;
constexpr int CS_STAGES = 3;
constexpr int CS_STAGE_BYTES = 2 * 16384;                     // A: 128 rows x 64 k  +  B: 128 rows (half of the 256 columns) x 64 k
constexpr int CS_RING_BYTES = CS_STAGES * CS_STAGE_BYTES;     // 96 KB
constexpr int CS_EPI_WARPS = 8;
constexpr int CS_EPI_BYTES = CS_EPI_WARPS * 2048;             // one 32 x 32 16-bit staging tile per epilogue warp
constexpr int CS_G_BAR_OFFSET = CS_RING_BYTES + CS_EPI_BYTES;
constexpr int CS_BAR_BYTES = 128;      // mbarriers + TMEM slot + role words
// Two CTAs per SM leave 115712 B of shared memory per CTA and role G needs 112 KB of 1024-aligned tiles, so there is no room
// for the usual "+1024 alignment slack, barriers behind the tiles" plan (and an __align__(1024) extern array costs 1 KB of
// static shared memory): the kernel asks for tiles + 1024 B, aligns the tiles inside, and puts the barriers into the
// alignment pad when that is >= 128 B and behind the tiles otherwise -- either way inside tiles + 1024.

__host__ __device__ inline int cs_attn_bar_offset(int kb) { return 2 * 16384 + 2 * kb * 128 + 16384 + 2048; }
__host__ __device__ inline int cs_smem_bytes(int kb) {
  const int a = cs_attn_bar_offset(kb), g = CS_G_BAR_OFFSET;
  return a > g ? a : g;
}

// layout of the scheduling words (ints) that follow the dependency counters in the workspace; zeroed before every launch
constexpr int CS_SCHED_TICKET = 0;       // [2]   next worker index per role
constexpr int CS_SCHED_ROLECTR = 2;      // [256] clusters that have arrived on a TPC, keyed by the smaller smid of the pair
constexpr int CS_SCHED_SMID = 258;       // [grid] smid of every CTA
// followed by [grid] (role, worker) per cluster
__host__ __device__ inline int cs_sched_ints(int grid) { return 258 + 2 * grid; }

struct CsGemmPhase {
  int M, N;
  int m_tiles, m_groups, n_tiles, tiles;
  const float* bias;       // fp32 per column or nullptr
  int out_dtype;           // 0 fp16, 1 bf16
  const int* wait_ctr;     // before loading A rows of a tile: wait_ctr[row / wait_rows] >= wait_target (nullptr: none)
  int wait_rows, wait_target;
  int* signal_ctr;         // after a warp's 32 rows of a tile are stored: signal_ctr[m-tile] += 1 (nullptr: none)
  uint32_t idesc;
};

struct CsParams {
  CsGemmPhase g[2];        // qkv projection, output projection (both contract over K)
  void* d[2];              // their outputs, row-major [M, N] 16-bit (contiguous rows)
  int K;
  int lag;                 // tile order: the proj tiles of m-group j follow the qkv tiles of m-group j + lag (see cs_tile)
  int tail;                // the proj tiles of the LAST `tail` m-groups are cut into 256 x 64 quarters (see cs_tile); lag == m_groups only
  int qtail;               // the same for the last `qtail` m-groups of the qkv phase
  AttnParams at;
  int* sched;
  long long* trace;        // debug: per CTA {role, worker, start, first unit ready, end} on the global timer, or nullptr
  int debug;               // timing experiments only (PA_CS_DEBUG; results invalid): 1 = role A idle, 2 = role G idle
  int* probe;              // residency self-test (see launch_vit_cosched): no work, probe[0] counts arrivals, probe[1] the CTAs
                           // that saw the whole grid resident (TMEM allocated, 2 CTAs per SM) before a ~5 ms time-out
};

__device__ __forceinline__ long long cs_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}

// debug timeline of worker 0 (leader CTA): clock64 stamps, G tiles at trace[4096 + seq*8 + slot], A units at trace[8192 + seq*16 + slot]
#define CS_TRACE_G(seq, slot_) do { if (P.trace != nullptr && worker == 0 && crank == 0 && (seq) < 64) P.trace[4096 + (seq) * 8 + (slot_)] = clock64(); } while (0)
#define CS_TRACE_GC(seq, chunk, slot_) do { if (P.trace != nullptr && worker == 0 && crank == 0 && (seq) == 2) P.trace[12288 + (chunk) * 8 + (slot_)] = clock64(); } while (0)
#define CS_TRACE_A(seq, slot_) do { if (P.trace != nullptr && worker == 0 && (seq) < 64) P.trace[8192 + (seq) * 16 + (slot_)] = clock64(); } while (0)

// Register budget: the kernel starts with 80 registers per thread (2 CTAs x 384 threads per SM).  Inside each role the control
// warpgroup (warps 0-3: TMA / MMA / store) hands most of its share to the two worker warpgroups (warps 4-11), whose
// 32-register TMEM loads otherwise spill around every load (measured: the max pass took 4.8 k cycles instead of 1.3 k).
// The instructions sit at the top of warpgroup-uniform branches so that ptxas allocates registers per region.
__device__ __forceinline__ void cs_regs_control() { asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory"); }
__device__ __forceinline__ void cs_regs_worker() { asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory"); }

// Global tile order of role G.  All qkv tiles of an m-group come before that m-group's proj tiles, `lag` m-groups later:
//     qkv(0) .. qkv(lag-1) | qkv(lag) proj(0) | qkv(lag+1) proj(1) | ... | qkv(MG-1) proj(MG-1-lag) | proj(MG-lag) .. proj(MG-1)
// lag = m_groups (the default) is "all qkv tiles, then all proj tiles".  Smaller lags were built to take the proj tiles of the
// last images off the kernel's tail and measured (tools/cosched_lag_sweep.py): they are slower, the smaller the worse -- the
// attention stream trails the qkv production, so an interleaved proj tile usually waits and holds up the qkv tiles queued
// behind it in its worker.  The order is deadlock-free for any lag >= 1: a proj tile waits
// for attention units that need only qkv tiles of m-groups <= its own + 1, all EARLIER in this order, and qkv tiles wait for
// nothing -- by induction over the order every tile's dependencies complete.
// Tail split: the kernel ends with the proj tiles of the last images, which cannot start before the attention stream has
// finished them -- a handful of 256 x 256 tiles (11-13 k cycles each) on a handful of workers while the other seventy idle.
// Those tiles (the last `tail` m-groups of the proj phase) are issued as 256 x 64 quarters instead: four times as many
// workers share the tail, each for a third of the time (a 64-wide tile is shared-memory-port-bound, ~310 cycles per k-block
// against 512 for the full width -- irrelevant when the alternative is idling).
// The qkv phase has the same problem one dependency earlier: 450 tiles on 74 workers are 6 rounds + 6 tiles, the attention of the
// last image cannot start before that seventh round has finished, and everything behind it (its attention units, their proj
// tiles) sits on the kernel's critical path.  `qtail`: the LAST m-groups of the qkv phase as quarters as well (the counters a
// quarter publishes weigh 1, a full tile's 4, so the attention role's target is the same for every m-tile).
struct CsTile { int ph, mg, col0, bn; };
__device__ __forceinline__ CsTile cs_tile(const CsParams& P, int t) {
  const int n1 = P.g[0].n_tiles, n2 = P.g[1].n_tiles, MG = P.g[0].m_groups, D = P.lag;
  constexpr int QPT = CS_BN / 64;                // quarters per tile
  CsTile r;
  r.bn = CS_BN;
  if (D == MG) {
    // [qkv full][qkv quarters][proj full][proj quarters]
    r.ph = 0;
    const int full1 = (MG - P.qtail) * n1;
    if (t < full1) { r.mg = t / n1; r.col0 = (t - r.mg * n1) * CS_BN; return r; }
    t -= full1;
    const int nq1 = n1 * QPT;
    if (t < P.qtail * nq1) { r.mg = MG - P.qtail + t / nq1; r.col0 = (t % nq1) * 64; r.bn = 64; return r; }
    t -= P.qtail * nq1;
    r.ph = 1;
    const int full2 = (MG - P.tail) * n2;
    if (t < full2) { r.mg = t / n2; r.col0 = (t % n2) * CS_BN; return r; }
    t -= full2;
    const int nq2 = n2 * QPT;
    r.mg = MG - P.tail + t / nq2;
    r.col0 = (t % nq2) * 64;
    r.bn = 64;
    return r;
  }
  if (t < D * n1) { r.ph = 0; r.mg = t / n1; r.col0 = (t - r.mg * n1) * CS_BN; return r; }
  t -= D * n1;
  const int per = n1 + n2, s = t / per;
  if (s < MG - D) {
    const int q = t - s * per;
    if (q < n1) { r.ph = 0; r.mg = D + s; r.col0 = q * CS_BN; }
    else { r.ph = 1; r.mg = s; r.col0 = (q - n1) * CS_BN; }
    return r;
  }
  t -= (MG - D) * per;
  r.ph = 1; r.mg = MG - D + t / n2; r.col0 = (t % n2) * CS_BN;
  return r;
}

// ================================================================================================ role G
__device__ __forceinline__ void cs_gemm_role(const CUtensorMap& tmA1, const CUtensorMap& tmB1, const CUtensorMap& tmD1,
                                             const CUtensorMap& tmA2, const CUtensorMap& tmB2, const CUtensorMap& tmD2,
                                             const CUtensorMap& tmB1q, const CUtensorMap& tmB2q,
                                             const CsParams& P, uint8_t* smem, uint64_t* bars, uint32_t tmem_base,
                                             int worker, int nworkers, int crank) {
  uint64_t* full_bar = bars;                 // [3] leader's: both CTAs' TMA bytes
  uint64_t* empty_bar = bars + CS_STAGES;    // [3] own: multicast commit of the leader
  uint64_t* tfull_bar = bars + 2 * CS_STAGES;       // accumulator ready (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 1;             // leader's: 8 epilogue warps x 2 CTAs
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (P.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int T0 = P.g[0].tiles, T = T0 + P.g[1].tiles;

  if (warp < 4) {
  cs_regs_control();
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0, tseq = 0;
    uint32_t phase = 0;
    for (int t = worker; t < T; t += nworkers, ++tseq) {
      const CsTile tl = cs_tile(P, t);
      const int ph = tl.ph, mg = tl.mg;
      const CsGemmPhase& g = P.g[ph];
      const int mt = mg * 2 + crank;
      const CUtensorMap* tA = ph ? &tmA2 : &tmA1;
      const CUtensorMap* tB = tl.bn == 64 ? (ph ? &tmB2q : &tmB1q) : ph ? &tmB2 : &tmB1;      // quarter tiles: 32 B rows per CTA
      const uint32_t stage_tx = 2 * (16384 + (uint32_t)(tl.bn / 2) * 128);    // both CTAs: A 128 x 64 + B (bn / 2) x 64
      if (g.wait_ctr != nullptr) {
        if (elect_one()) {
          const int r0 = mt * GEMM_BLOCK_M, r1 = min(r0 + GEMM_BLOCK_M, g.M) - 1;
          if (r0 < g.M)
            for (int w = r0 / g.wait_rows; w <= r1 / g.wait_rows; ++w) wait_counter_ge(g.wait_ctr + w, g.wait_target);
        }
        __syncwarp();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * CS_STAGE_BYTES;
        if (elect_one()) {
          if (kb == 0) CS_TRACE_G(tseq, 0);
          if (crank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_3d_2sm(sa, tA, kb * GEMM_BLOCK_K, mt * GEMM_BLOCK_M, 0, &full_bar[stage]);
          tma_load_3d_2sm(sa + 16384, tB, kb * GEMM_BLOCK_K, tl.col0 + crank * (tl.bn / 2), 0, &full_bar[stage]);
        }
        __syncwarp();
        if (++stage == CS_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of the pair) =====================
    if (crank == 0) {
      int stage = 0, tseq = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int t = worker; t < T; t += nworkers, ++tseq) {
        const CsTile tlm = cs_tile(P, t);
        const uint32_t idesc = tlm.bn == CS_BN ? P.g[tlm.ph].idesc
                                               : (P.g[tlm.ph].idesc & ~(0x3Fu << 17)) | ((uint32_t)(tlm.bn >> 3) << 17);
        mbar_wait(tempty_bar, acc_phase ^ 1);           // single accumulator: drained by both CTAs' epilogue warps
        acc_phase ^= 1;
        tc_fence_after();
        if (lane == 0) CS_TRACE_G(tseq, 1);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0 && kb == 0) CS_TRACE_G(tseq, 2);
          const uint32_t sa = smem_base + stage * CS_STAGE_BYTES;
          const uint64_t adesc = make_sdesc(sa, 16, 1024, PA_SWZ_128B);
          const uint64_t bdesc = make_sdesc(sa + 16384, 16, 1024, PA_SWZ_128B);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) umma_ss2(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit2_mc(&empty_bar[stage], 3);
            if (kb == num_kb - 1) umma_commit2_mc(tfull_bar, 3);
          }
          __syncwarp();
          if (++stage == CS_STAGES) { stage = 0; phase ^= 1; }
        }
        if (lane == 0) CS_TRACE_G(tseq, 3);
      }
    }
  }
  } else {
    cs_regs_worker();
    // ===================== epilogue: warp (q, half) drains rows 32q..32q+31, 32-column chunks half, half+2, ... ============
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* wbuf = smem + CS_RING_BYTES + (warp - 4) * 2048;
    uint32_t acc_phase = 0;
    int tseq = 0;
    const bool tracer = (warp == 4 && lane == 0);
    for (int t = worker; t < T; t += nworkers, ++tseq) {
      const CsTile tl = cs_tile(P, t);
      const int ph = tl.ph, mg = tl.mg;
      const int nchunks = tl.bn / 64;                 // 32-column chunks of this warp: 4 (full tile) or 1 (quarter tile)
      const CsGemmPhase& g = P.g[ph];
      const int mt = mg * 2 + crank;
      const int row0 = mt * GEMM_BLOCK_M + q * 32;
      const CUtensorMap* tD = ph ? &tmD2 : &tmD1;
      mbar_wait(tfull_bar, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      if (tracer) CS_TRACE_G(tseq, 4);
      // Four 32-column chunks per warp: TMEM -> registers -> the warp's private 2 KB staging tile (64-byte swizzle) -> one bulk
      // store.  With two CTAs of 113 KB on the SM there is no L1 left: the column biases come from L2, so the 8 x 16-byte
      // bias loads of chunk j+1 are issued while chunk j is converted and staged (measured: biased tiles drained in 4.6 k
      // cycles with the loads in line, unbiased ones in 3.0 k).  Measured and rejected: the TMEM load of chunk j+1 in flight
      // during chunk j (no gain), rows stored straight from registers (drain 2.9 k -> 5.5 k cycles); late round 2: all four chunks
      // held as packed registers and staged under the next mainloop (accumulator free 1.5 k cycles earlier, mainloop 0.7 k longer:
      // 87.2 vs 86.2 us, and again 86.7 vs 85.0 us on top of the tail split) -- profiles/cosched_drain_variants_r02.txt.
      const bool rows_ok = row0 < g.M;
      const bool has_bias = g.bias != nullptr;
      float4 bq[8];
      if (has_bias && tl.col0 + half * 32 < g.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(g.bias + tl.col0 + half * 32) + i);
      }
#pragma unroll 1
      for (int j = 0; j < nchunks; ++j) {
        const int c = half * 32 + j * 64;
        const int col = tl.col0 + c;
        uint32_t v[32];
        tmem_ld32(t_base + c, v);
        tmem_ld_wait();
        if (j == nchunks - 1) {
          // this warp's last TMEM read of the tile: hand the accumulator back before converting the chunk
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tempty_bar);
          if (tracer) CS_TRACE_G(tseq, 5);
        }
        if (col < g.N && rows_ok) {
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + bq[i].x);
              v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + bq[i].y);
              v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + bq[i].z);
              v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + bq[i].w);
            }
            if (j + 1 < nchunks && col + 64 < g.N) {
#pragma unroll
              for (int i = 0; i < 8; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(g.bias + col + 64) + i);
            }
          }
          if (lane == 0 && !(P.debug & 4)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile read by the previous store
          __syncwarp();
          uint8_t* rowp = wbuf + lane * 64;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * ch + k]);
            *reinterpret_cast<uint4*>(rowp + ((ch ^ ((lane >> 1) & 3)) << 4)) =
                g.out_dtype == 0 ? make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]))
                                 : make_uint4(pack_bf2(f[0], f[1]), pack_bf2(f[2], f[3]), pack_bf2(f[4], f[5]), pack_bf2(f[6], f[7]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(P.debug & 8)) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(tD)),
                         "r"(smem_u32(wbuf)), "r"(col), "r"(row0), "r"(0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (tracer) CS_TRACE_G(tseq, 6);
      // publish right away: with a single accumulator this warp has nothing to do until the next mainloop has finished
      if (lane == 0 && g.signal_ctr != nullptr && mt < g.m_tiles) signal_counter(g.signal_ctr + mt, tl.bn == CS_BN ? CS_BN / 64 : 1);
      if (tracer) CS_TRACE_G(tseq, 7);
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ================================================================================================ role A
// unit = (image g, head h, 128-row query tile qt); one 256-column TMEM slot: S fp32 [0, kb) -> P fp16 in place; O fp32 [192, 256)
template <bool RELPOS>
__device__ __forceinline__ void cs_attn_role(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                             const CUtensorMap& tmO, const CsParams& P, uint8_t* smem, uint64_t* bars,
                                             uint32_t tmem_base, int worker, int nworkers) {
  using Cfg = AttnCfg<64>;
  const AttnParams& p = P.at;
  const int kbytes = p.kb * 128;
  uint8_t* q_smem = smem;                       // [2][128 rows x 128 B]
  uint8_t* k_smem = smem + 2 * 16384;
  uint8_t* v_smem = k_smem + kbytes;
  uint8_t* o_smem = v_smem + kbytes;            // 128 x 64 fp16 output tile (SW128)
  float* xch = reinterpret_cast<float*>(o_smem + 16384);   // xmax[2][128], xsum[2][128]
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* k_full = bars + 4;
  uint64_t* k_empty = bars + 5;
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 8;
  uint64_t* p_full = bars + 9;      // 8 softmax warps
  uint64_t* o_full = bars + 10;
  uint64_t* slot_empty = bars + 11; // 8 softmax warps: O is in registers
  uint64_t* ost_full = bars + 12;   // 8 softmax warps: output tile staged
  uint64_t* ost_empty = bars + 13;  // store warp: staging tile read by the bulk store
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int units = p.G * p.H * p.q_tiles;

  if (warp < 4) {
  cs_regs_control();
  if (warp == 0) {
    // ===================== TMA producer =====================
    int it = 0;
    for (int u = worker; u < units; u += nworkers, ++it) {
      const int qt = u % p.q_tiles, gh = u / p.q_tiles;
      const int h = gh % p.H, g = gh / p.H;
      const int b = it & 1;
      if (p.wait_ctr != nullptr) {
        if (elect_one()) {
          const long long r0 = (long long)g * p.wait_rows_per_group, r1 = r0 + p.n_k - 1;
          for (int t = (int)(r0 >> 7); t <= (int)(r1 >> 7); ++t) wait_counter_ge(p.wait_ctr + t, p.wait_target);
          if (it == 0 && P.trace != nullptr) P.trace[blockIdx.x * 8 + 3] = cs_globaltimer();
        }
        __syncwarp();
      }
      if (lane == 0) CS_TRACE_A(it, 0);
      mbar_wait(&q_empty[b], ((it >> 1) & 1) ^ 1);
      mbar_wait(k_empty, (it & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&q_full[b], 16384);
        tma_load_3d(q_smem + b * 16384, &tmQ, p.q_col0 + h * 64, qt * 128, g, &q_full[b]);
        mbar_expect_tx(k_full, kbytes);
        tma_load_3d(k_smem, &tmK, p.k_col0 + h * 64, 0, g, k_full);
      }
      __syncwarp();
      mbar_wait(v_empty, (it & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(v_full, kbytes);
        tma_load_3d(v_smem, &tmV, p.v_col0 + h * 64, 0, g, v_full);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const int ksteps_o = p.kb / 16;
    const int h16 = (ksteps_o + 1) / 2;
    const uint32_t q_base = smem_u32(q_smem), k_base = smem_u32(k_smem), v_base = smem_u32(v_smem);
    const uint64_t kdesc = make_sdesc(k_base, 16, Cfg::SBO, Cfg::SWZ);
    const uint64_t vdesc = make_sdesc(v_base, Cfg::SBO, Cfg::SBO, Cfg::SWZ);    // V [key][d]: MN-major B operand
    // S columns below the O accumulator hold nothing the previous unit still needs once its PV has been issued: that part
    // of the next S is issued before the wait for the slot hand-back (see pa_attn.cuh)
    const int s_n1 = p.kb < Cfg::O_COL ? p.kb : Cfg::O_COL;
    const uint32_t idesc_s1 = (p.idesc_s & ~(0x3Fu << 17)) | ((uint32_t)(s_n1 >> 3) << 17);
    const uint32_t idesc_s2 = (p.idesc_s & ~(0x3Fu << 17)) | ((uint32_t)((p.kb - s_n1) >> 3) << 17);
    const uint64_t k_off2 = (uint64_t)((s_n1 * 128) >> 4);
    int it = 0;
    for (int u = worker; u < units; u += nworkers, ++it) {
      const int b = it & 1;
      mbar_wait(&q_full[b], (it >> 1) & 1);
      mbar_wait(k_full, it & 1);
      tc_fence_after();
      const uint64_t qdesc = make_sdesc(q_base + b * 16384, 16, Cfg::SBO, Cfg::SWZ);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base, qdesc + 2 * k, kdesc + 2 * k, idesc_s1, k != 0);
      }
      __syncwarp();
      if (it > 0) {
        mbar_wait(slot_empty, (it - 1) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        if (p.kb > s_n1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + s_n1, qdesc + 2 * k, kdesc + k_off2 + 2 * k, idesc_s2, k != 0);
        }
        umma_commit(s_full);
        umma_commit(k_empty);
        umma_commit(&q_empty[b]);
      }
      __syncwarp();
      if (lane == 0) CS_TRACE_A(it, 1);
      mbar_wait(p_full, it & 1);
      mbar_wait(v_full, it & 1);
      tc_fence_after();
      if (elect_one()) {
        for (int k = 0; k < ksteps_o; ++k) {
          const int pcol = (k < h16) ? 8 * k : 16 * h16 + 8 * (k - h16);
          umma_ts(tmem_base + Cfg::O_COL, tmem_base + pcol, vdesc + Cfg::V_KSTEP * k, p.idesc_o, k != 0);
        }
        umma_commit(o_full);
        umma_commit(v_empty);
      }
      __syncwarp();
      if (lane == 0) CS_TRACE_A(it, 7);
    }
  } else if (warp == 2) {
    // ===================== output store + publication =====================
    int it = 0;
    for (int u = worker; u < units; u += nworkers, ++it) {
      const int qt = u % p.q_tiles, gh = u / p.q_tiles;
      const int h = gh % p.H, g = gh / p.H;
      mbar_wait(ost_full, it & 1);
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmO)),
                     "r"(smem_u32(o_smem)), "r"(p.o_col0 + h * 64), "r"(qt * 128), "r"(g)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(ost_empty);
        if (p.signal_ctr != nullptr) signal_counter(p.signal_ctr + g);
        CS_TRACE_A(it, 8);
      }
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  } else {
    cs_regs_worker();
    // ===================== softmax + epilogue: two threads per query row =====================
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
    const uint32_t t_o = t_slot + Cfg::O_COL + hf * 32;
    const int nvalid = min(p.kb, p.n_k) - c_lo;
    const bool tracer = (warp == 4 && lane == 0);
    int it = 0;
    for (int u = worker; u < units; u += nworkers, ++it) {
      const int qt = u % p.q_tiles;
      const bool warp_active = (qt * 128 + q * 32) < p.n_q;
      // additive score bias (RELPOS, cmt.py:100): this thread's row of rel_pos[h], columns from c_lo on, multiplied by 1 / scale
      // when it is added to the raw scores -- both passes then run unchanged on s + r / scale.  With at most 64 keys (every CMT
      // stage of the zoo has 49) the thread's <= 32 values are requested HERE, before the wait for S, and stay in registers for
      // both passes (first version: 16 scalar loads per step inside each pass, every one an exposed L2 round trip -- the
      // attention of a config-3-sized CMT forward took 565 us instead of 69)
      const float* rp = nullptr;
      float rb0[16], rb1[16];
      float thr = -INFINITY;                  // kvt.KNNAttention: raw scores below the row's k-th largest are masked
      const bool rel_pre = RELPOS && nst <= 2;
      if constexpr (RELPOS) {
        const int gh = u / p.q_tiles, h = gh % p.H;
        const int row = min(qt * 128 + trow, p.n_q - 1);
        if (p.row_thresh != nullptr) thr = __ldg(p.row_thresh + (size_t)gh * p.n_q + row);
        if (p.rel_pos == nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { rb0[i] = 0.f; rb1[i] = 0.f; }
        }
        rp = p.rel_pos != nullptr ? p.rel_pos + ((size_t)h * p.n_q + row) * p.n_k + c_lo : nullptr;
        if (rp != nullptr && rel_pre && warp_active) {
          if ((p.n_k & 3) == 0) {                    // rows 16-byte aligned (c_lo is a multiple of 16 columns)
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 a = (i < nvalid) ? __ldg(reinterpret_cast<const float4*>(rp + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 b = (nst > 1 && 16 + i < nvalid) ? __ldg(reinterpret_cast<const float4*>(rp + 16 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
              rb0[i] = a.x; rb0[i + 1] = a.y; rb0[i + 2] = a.z; rb0[i + 3] = a.w;
              rb1[i] = b.x; rb1[i + 1] = b.y; rb1[i + 2] = b.z; rb1[i + 3] = b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              rb0[i] = (i < nvalid) ? __ldg(rp + i) : 0.f;
              rb1[i] = (nst > 1 && 16 + i < nvalid) ? __ldg(rp + 16 + i) : 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) { rb0[i] *= p.rel_mul; rb1[i] *= p.rel_mul; }
        }
      }
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      if (tracer) CS_TRACE_A(it, 2);
      // ---- pass 1: partial row max
      float mx = -INFINITY;
      if constexpr (RELPOS) {
        if (warp_active) {
#pragma unroll 1
          for (int k = 0; k < nst; ++k) {
            uint32_t v[16];
            const int nv = nvalid - k * 16;
            tmem_ld16(t_my + k * 16, v);
            if (rel_pre) {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __uint_as_float(v[i]);
                v[i] = __float_as_uint(sv >= thr ? sv + (k == 0 ? rb0[i] : rb1[i]) : -INFINITY);
              }
            } else {
              float r[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = (rp != nullptr && i < nv) ? __ldg(rp + k * 16 + i) * p.rel_mul : 0.f;
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __uint_as_float(v[i]);
                v[i] = __float_as_uint(sv >= thr ? sv + r[i] : -INFINITY);
              }
            }
            mx = chunk_max<16>(v, nv, mx);
          }
        }
      } else
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
      mx = fmaxf(mx, xmax[(hf ^ 1) * 128 + trow]);
      if (tracer) CS_TRACE_A(it, 3);
      float l_run = 0.f;
      if (warp_active) {
        // ---- pass 2: exponentials, fp16 P in place, partial row sum
        const float mxs = mx * sl2;
        uint32_t va[16], pk[8];
        float e[16];
        float s0 = 0.f, s1 = 0.f;
        // (measured again in this kernel, two softmax warps per scheduler instead of four: with the TMEM load of step k+1 in
        //  flight during step k the pass takes 3.3 k cycles instead of 2.45 k alone, 4.2 k instead of 2.9 k next to the GEMM role)
        // (also measured: 32 columns per TMEM round trip -- 2.45 k vs 2.34 k cycles, no gain: the pass is bound by the MUFU /
        //  issue rate of the two softmax warps per scheduler, not by the number of round trips)
#pragma unroll 1
        for (int k = 0; k < nst; ++k) {
          tmem_ld16(t_my + k * 16, va);
          if constexpr (RELPOS) {
            if (rel_pre) {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __uint_as_float(va[i]);
                va[i] = __float_as_uint(sv >= thr ? sv + (k == 0 ? rb0[i] : rb1[i]) : -INFINITY);
              }
            } else {
              float r[16];
              const int nv = nvalid - k * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = (rp != nullptr && i < nv) ? __ldg(rp + k * 16 + i) * p.rel_mul : 0.f;
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float sv = __uint_as_float(va[i]);
                va[i] = __float_as_uint(sv >= thr ? sv + r[i] : -INFINITY);
              }
            }
          } else {
            tmem_ld_wait();
          }
          exp_stage(va, e, nvalid - k * 16, sl2, mxs);
          pack_stage(e, pk, s0, s1);
          tmem_st8(t_my + k * 8, pk);
        }
        tmem_st_wait();
        l_run = s0 + s1;
      }
      xsum[hf * 128 + trow] = l_run;            // ordered towards the partner by p_full -> o_full
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tracer) CS_TRACE_A(it, 4);
      // ---- epilogue
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      if (tracer) CS_TRACE_A(it, 5);
      uint32_t v[32];
      float l_other = 0.f;
      if (warp_active) {
        tmem_ld32(t_o, v);
        tmem_ld_wait();
        l_other = xsum[(hf ^ 1) * 128 + trow];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_empty);   // O is in registers: the next S may overwrite the slot
      mbar_wait(ost_empty, (it & 1) ^ 1);       // staging tile free
      if (warp_active) {
        const float inv = 1.f / (l_run + l_other);
        uint8_t* rowp = o_smem + trow * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * i + k]) * inv;
          const int ch = hf * 4 + i;
          *reinterpret_cast<uint4*>(rowp + ((ch ^ (trow & 7)) << 4)) =
              make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(ost_full);
      if (tracer) CS_TRACE_A(it, 6);
    }
  }
}

// ================================================================================================ kernel
__global__ void __launch_bounds__(CS_THREADS, 2)
vit_cosched_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmD1, const __grid_constant__ CUtensorMap tmQ,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmD2,
                   const __grid_constant__ CUtensorMap tmB1q, const __grid_constant__ CUtensorMap tmB2q, const CsParams P) {
  extern __shared__ uint8_t cs_raw[];
  const uint32_t pad = (1024u - (smem_u32(cs_raw) & 1023u)) & 1023u;
  uint8_t* smem = cs_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pad >= CS_BAR_BYTES ? cs_raw : smem + cs_smem_bytes(P.at.kb));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  volatile int* s_info = reinterpret_cast<volatile int*>(bars + 15);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int cluster = blockIdx.x >> 1;
  const int grid = gridDim.x;

  // ---- role election (see the header comment)
  if (threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    __stcg(P.sched + CS_SCHED_SMID + blockIdx.x, (int)smid);
  }
  cluster_sync_all();
  if (threadIdx.x == 0 && crank == 0) {
    const int s0 = __ldcg(P.sched + CS_SCHED_SMID + blockIdx.x), s1 = __ldcg(P.sched + CS_SCHED_SMID + blockIdx.x + 1);
    const int key = (s0 < s1 ? s0 : s1) & 255;
    int role = atomicAdd(P.sched + CS_SCHED_ROLECTR + key, 1) & 1;
    int w = atomicAdd(P.sched + CS_SCHED_TICKET + role, 1);
    if (w >= grid / 4) {
      // more than half of the clusters asked for this role (unexpected placement): the counts per role must stay exact,
      // so this cluster takes the other role -- only the pairing on its SMs, i.e. speed, is affected
      role ^= 1;
      w = atomicAdd(P.sched + CS_SCHED_TICKET + role, 1);
    }
    __stcg(P.sched + CS_SCHED_SMID + grid + 2 * cluster, role);
    __stcg(P.sched + CS_SCHED_SMID + grid + 2 * cluster + 1, w);
  }
  cluster_sync_all();
  if (threadIdx.x == 0) {
    s_info[0] = __ldcg(P.sched + CS_SCHED_SMID + grid + 2 * cluster);
    s_info[1] = __ldcg(P.sched + CS_SCHED_SMID + grid + 2 * cluster + 1);
  }
  __syncthreads();
  const int role = s_info[0];
  const int cworker = s_info[1];          // worker index of the cluster within its role
  const int nclusters_role = grid / 4;    // half of the clusters per role

  if (warp == 0 && lane == 0) {
    if (role == 0) {
      tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmD1);
      tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmD2); tma_prefetch_desc(&tmB1q); tma_prefetch_desc(&tmB2q);
    } else {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    }
  }
  if (warp == 1 && lane == 0) {
    if (role == 0) {
      for (int i = 0; i < CS_STAGES; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[CS_STAGES + i], 1); }
      mbar_init(&bars[2 * CS_STAGES], 1);
      mbar_init(&bars[2 * CS_STAGES + 1], 2 * CS_EPI_WARPS);
    } else {
      for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);      // q_full[2] q_empty[2] k_full k_empty v_full v_empty s_full
      mbar_init(&bars[9], 8);                                   // p_full
      mbar_init(&bars[10], 1);                                  // o_full
      mbar_init(&bars[11], 8);                                  // slot_empty
      mbar_init(&bars[12], 8);                                  // ost_full
      mbar_init(&bars[13], 1);                                  // ost_empty
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (role == 0) { tmem_alloc2(tmem_slot, 256); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (P.trace != nullptr && threadIdx.x == 0) {
    P.trace[blockIdx.x * 8 + 0] = role;
    P.trace[blockIdx.x * 8 + 1] = cworker;
    P.trace[blockIdx.x * 8 + 2] = cs_globaltimer();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    P.trace[blockIdx.x * 8 + 5] = smid;
  }

  if (P.probe != nullptr) {
    if (threadIdx.x == 0) {
      atomicAdd(P.probe, 1);
      const long long t0 = clock64();
      while (clock64() - t0 < 10000000ll) {
        if (*reinterpret_cast<volatile int*>(P.probe) >= grid) { atomicAdd(P.probe + 1, 1); break; }
        __nanosleep(500);
      }
    }
  } else if (role == 0) {
    if (!(P.debug & 2)) cs_gemm_role(tmA1, tmB1, tmD1, tmA2, tmB2, tmD2, tmB1q, tmB2q, P, smem, bars, tmem_base, cworker, nclusters_role, crank);
  } else {
    if (!(P.debug & 1)) cs_attn_role<false>(tmQ, tmK, tmV, tmO, P, smem, bars, tmem_base, cworker * 2 + crank, nclusters_role * 2);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (P.trace != nullptr && threadIdx.x == 0) P.trace[blockIdx.x * 8 + 4] = cs_globaltimer();
  if (warp == 2) {
    if (role == 0) tmem_dealloc2(tmem_base, 256);
    else tmem_dealloc(tmem_base, 256);
  }
}

// The attention role as a kernel of its own (no GEMM role, no dependency counters): two single-slot CTAs per SM instead of the
// two-slot CTA of attn_core_kernel.  Used for 64-wide heads with a single key block (ViT / PVT / CvT three-launch paths).
// RELPOS: the score-modifying variant -- an additive fp32 score bias [H, n_q, n_k] before the softmax (cmt.Attention, cmt.py:100)
// and / or a per-row threshold below which scores are masked (kvt.KNNAttention, kvt.py:84-87).
template <bool RELPOS>
__global__ void __launch_bounds__(CS_THREADS, 2)
attn_single_slot_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const CsParams P) {
  extern __shared__ uint8_t cs_raw[];
  const uint32_t pad = (1024u - (smem_u32(cs_raw) & 1023u)) & 1023u;
  uint8_t* smem = cs_raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pad >= CS_BAR_BYTES ? cs_raw : smem + cs_attn_bar_offset(P.at.kb));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bars[9], 8);
    mbar_init(&bars[10], 1);
    mbar_init(&bars[11], 8);
    mbar_init(&bars[12], 8);
    mbar_init(&bars[13], 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cs_attn_role<RELPOS>(tmQ, tmK, tmV, tmO, P, smem, bars, tmem_base, (int)blockIdx.x, (int)gridDim.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

}  // namespace pa
