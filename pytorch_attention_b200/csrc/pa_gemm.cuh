// pa_gemm.cuh — persistent warp-specialised tcgen05 GEMM for the projection stages.
//
//   D[z][m, n] = sum_k A[z][m, k] * B[z][n, k]  (+ bias) (+ residual)      A, B K-major ("TN", nn.Linear layout)
//
// One CTA per SM, 256 threads:
//   warp 0   TMA producer      (cp.async.bulk.tensor, 128B-swizzled 64-wide K blocks, STAGES-deep ring)
//   warp 1   MMA issuer        (one thread: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue         tcgen05.ld 32x32b -> bias/residual -> convert -> swizzled smem staging -> TMA bulk store
//                              (coalesced, asynchronous; double-buffered TMEM accumulator so it overlaps the next mainloop).
//                              Outputs whose row pitch is not a multiple of 16 B fall back to direct global stores.
// Tiles are distributed round-robin over the persistent grid; rows beyond M / K are zero-filled by TMA
// (3-D tensor maps {K, rows, Z}); the TMA store clips at the tensor edge, so M, N need not divide the tile.
#pragma once
#include "pa_ptx.cuh"

namespace pa {

struct GemmParams {
  int M, N, K, Z;          // per-batch problem, Z batches
  int m_tiles, n_tiles;    // ceil(M/128), ceil(N/BLOCK_N)
  int m_groups;            // ceil(m_tiles / CLUSTER): a cluster works on CLUSTER adjacent m-tiles of one n-tile
  int a_batched, b_batched;  // 1: operand has a batch (z) coordinate, 0: shared across z
  void* D;
  long long ldd;           // row pitch of D in elements
  long long d_batch;       // batch pitch of D in elements
  const float* bias;       // fp32 bias or nullptr
  int bias_mode;           // 0 none, 1 per column n, 2 per row m
  int out_dtype;           // 0 fp16, 1 bf16, 2 fp32
  const void* residual;    // optional [M, N] tensor added in the epilogue (16-bit or fp32), or nullptr
  long long ldr, r_batch;
  int res_dtype;
  int tma_store;           // 1: epilogue goes through smem + TMA store (tmD valid)
  // dependency hooks for the fused single-launch kernels (nullptr = none).  Counters live in global memory.
  const int* wait_ctr;     // before loading A rows [r0, r1] of a tile: wait_ctr[r / wait_rows] >= wait_target for both ends
  int wait_rows, wait_target;
  int* signal_ctr;         // after a CTA's 128-row tile has been stored completely: signal_ctr[m-tile] += 1
  int worker_shift;        // fused kernels: worker w walks the tile sequence of virtual worker (w + worker_shift) mod n
  int balanced;            // PAIR + BLOCK_N 256 only: balanced contiguous partition of 64-column units (see TileWalk)
  int n_units;             // ceil(N / 64)
  uint32_t idesc;
  long long* trace;        // debug: per-tile clock64 stamps of CTA 0 ([tile][8]), or nullptr
  int debug_flags;         // debug experiments (env PA_GEMM_DEBUG): 1 = skip epilogue body
};

// trace slots: 0 kernel start | 1 first MMA of tile issued (tempty ok) | 2 first full barrier of tile passed
//              3 last MMA of tile issued | 4 epilogue: accumulator ready | 5 epilogue: tile drained | 6 producer: first load of tile issued
__device__ __forceinline__ void trace_stamp(const GemmParams& p, int tile_seq, int slot) {
  if (p.trace != nullptr && blockIdx.x == 0 && tile_seq < 64) p.trace[tile_seq * 8 + slot] = clock64();
}
// steps inside the epilogue of tile 2 (warp 4 lane 0): rows 40.. of the trace, 4 stamps per 32-column sub-tile
// (compiled in only with -DPA_TRACE_CHUNKS: the four clock reads per sub-tile cost the stand-alone GEMM 12 %)
__device__ __forceinline__ void trace_chunk(const GemmParams& p, int tile_seq, int chunk, int slot) {
#ifdef PA_TRACE_CHUNKS
  if (p.trace != nullptr && blockIdx.x == 0 && tile_seq == 2 && chunk < 8) p.trace[320 + chunk * 4 + slot] = clock64();
#endif
}

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_CSTAGE_BYTES = 128 * 32 * 4;   // one 128 x 32 output sub-tile in fp32 (16-bit uses half)

template <int BLOCK_N, int STAGES, bool PAIR = false>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;       // cta_group::2: each CTA stages half of the B tile
  static constexpr int B_BYTES = B_ROWS * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int C_OFFSET = STAGES * STAGE_BYTES;             // 2 output staging buffers
  static constexpr int BIAS_OFFSET = C_OFFSET + 2 * GEMM_CSTAGE_BYTES;   // BLOCK_N fp32 column biases of the current tile
  static constexpr int BAR_OFFSET = BIAS_OFFSET + 1024;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;        // + barriers + 1024-alignment slack
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
};

__device__ __forceinline__ void epi_bar_sync(int set = 0) { asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory"); }
template <int ESETS>
__device__ __forceinline__ void epi_bar_sync_all() { asm volatile("bar.sync 5, %0;" ::"n"(128 * ESETS) : "memory"); }   // all epilogue sets

// Tile sequence of one persistent worker (a CTA, or a cluster).  Classic: tiles `first, first+stride, ...` of fixed width.
// Balanced: the (m-group, 64-column unit) grid is cut into equal contiguous ranges, one per worker, and each range is walked
// as tiles of up to four units inside one m-group -- no partially filled last wave (e.g. 450 tiles on 74 pairs = 6.08 waves
// cost 7 rounds in the classic order, 6.08 here).  All three roles of a CTA walk the same sequence.
struct TileWalk {
  long long cur, end;      // balanced: unit range; classic: tile index / count
  int stride, per_z, n_cols_units, block_n, balanced;
  int z, mg, col0, ncols;
  __device__ TileWalk(const GemmParams& p, int worker, int nworkers, int block_n_) {
    balanced = p.balanced; block_n = block_n_;
    if (balanced) {
      n_cols_units = p.n_units;
      per_z = p.m_groups * p.n_units;
      const long long total = (long long)per_z * p.Z;
      cur = total * worker / nworkers;
      end = total * (worker + 1) / nworkers;
      stride = 0;
    } else {
      n_cols_units = p.n_tiles;
      per_z = p.m_groups * p.n_tiles;
      cur = worker; end = (long long)per_z * p.Z; stride = nworkers;
    }
  }
  __device__ bool next() {
    if (cur >= end) return false;
    z = (int)(cur / per_z);
    const int r = (int)(cur - (long long)z * per_z);
    mg = r / n_cols_units;
    const int u = r - mg * n_cols_units;
    if (balanced) {
      int w = n_cols_units - u;
      if (w > 4) w = 4;
      if (w > end - cur) w = (int)(end - cur);
      col0 = u * 64; ncols = w * 64; cur += w;
    } else {
      col0 = u * block_n; ncols = block_n; cur += stride;
    }
    return true;
  }
};

// CLUSTER > 1: the CTAs of a cluster take adjacent m-tiles of the same n-tile; each loads 1/CLUSTER of the B tile and
// TMA-multicasts it to all of them (L2->SM traffic per flop drops from (128+BN) to (128+BN/CLUSTER) rows per k-block).
// A stage may be refilled only after EVERY CTA's MMAs have read it: the consumer release is a multicast commit.
//
// PAIR (requires CLUSTER == 2): tcgen05 cta_group::2.  The two CTAs of a cluster form one 256 x BLOCK_N tile: each stages
// its own 128 rows of A and HALF of the B tile, the leader (rank 0) issues M=256 MMAs that read both CTAs' shared memory
// and write both CTAs' TMEM.  Shared-memory traffic per flop drops by a third versus two independent 128 x BLOCK_N CTAs
// (the 1-CTA kernel is capped at 128/192 = 67 % of the tensor pipe by TMA-write + MMA-read shared-memory bandwidth).
//   full[stage]   leader's barrier, expect_tx = bytes of BOTH CTAs (every TMA load signals the leader's barrier)
//   empty[stage]  each CTA's own barrier, released by the leader's multicast tcgen05.commit
//   tfull[acc]    each CTA's own barrier (multicast commit);  tempty[acc]: leader's, 8 arrivals (4 epilogue warps x 2 CTAs)
template <int STAGES, int CLUSTER, bool PAIR, int ESETS = 1>
__device__ __forceinline__ void gemm_init_barriers(uint64_t* bars) {      // one thread; bars: [2*STAGES + 4] mbarriers
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  for (int i = 0; i < STAGES; ++i) {
    mbar_init(&full_bar[i], 1);
    mbar_init(&empty_bar[i], PAIR ? 1 : CLUSTER);
  }
  for (int i = 0; i < 2; ++i) {
    mbar_init(&tfull_bar[i], 1);
    mbar_init(&tempty_bar[i], (PAIR ? 8 : 4) * ESETS);
  }
}

// The three roles of one CTA (warp 0 producer, warp 1 MMA issuer, warps 4-7 epilogue) over its whole tile sequence.
// Barriers must be initialised and visible (cluster-wide when CLUSTER > 1) and TMEM allocated before the call.
// LEAN: compile out the rarely used epilogue variants (row bias, residual, unaligned-output fallback) -- the fused kernels
// run 20 warps per CTA and have only 96 registers per thread.
// ESETS: 1 = warps 4-7 drain the accumulator; 2 / 4 = warps 8-11 / 8-19 join and the sets take the 32-column sub-tiles round-robin (each
// set has its own staging region, named barrier and bulk-store groups; both sets publish to signal_ctr).
template <int BLOCK_N, int STAGES, int CLUSTER, bool PAIR, bool LEAN = false, int ESETS = 1>
__device__ __forceinline__ void gemm_run(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                                         const GemmParams& p, uint8_t* smem, uint64_t* bars, uint32_t tmem_base) {
  static_assert(!PAIR || CLUSTER == 2, "cta_group::2 needs a cluster of exactly two CTAs");
  using Cfg = GemmCfg<BLOCK_N, STAGES, PAIR>;
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;    // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int crank = (CLUSTER > 1) ? (int)cluster_ctarank() : 0;
  const int ncl = gridDim.x / CLUSTER;
  const int cid = ((int)(blockIdx.x / CLUSTER) + p.worker_shift) % ncl;   // persistent worker index
  constexpr uint16_t CMASK = (uint16_t)((1u << CLUSTER) - 1);
  constexpr int B_SLICE_ROWS = BLOCK_N / CLUSTER;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    int stage = 0, tseq = 0;
    uint32_t phase = 0;
    if (lane == 0) trace_stamp(p, 0, 0);
    TileWalk tw(p, cid, ncl, BLOCK_N);
    for (; tw.next(); ++tseq) {
      const int z = tw.z;
      const int mt = tw.mg * CLUSTER + crank;
      if (p.wait_ctr != nullptr) {
        // fused kernels: the rows of this A tile are produced by an earlier phase on other SMs
        if (elect_one()) {
          const int r0 = mt * GEMM_BLOCK_M, r1 = min(r0 + GEMM_BLOCK_M, p.M) - 1;
          if (r0 < p.M) {
            const int g0 = r0 / p.wait_rows, g1 = r1 / p.wait_rows;
            for (int g = g0; g <= g1; ++g) wait_counter_ge(p.wait_ctr + g, p.wait_target);
          }
        }
        __syncwarp();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        if (elect_one()) {
          if (kb == 0) trace_stamp(p, tseq, 6);
          if (PAIR) {
            const int b_rows = tw.ncols / 2;               // rows of B this CTA stages (half of the tile's columns)
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + b_rows * 128));
            tma_load_3d_2sm(sa, &tmA, kb * GEMM_BLOCK_K, mt * GEMM_BLOCK_M, p.a_batched ? z : 0, &full_bar[stage]);
            if (p.balanced) {                              // B map has 32-row boxes: one per 64-column unit of the tile
              for (int i = 0; i < b_rows; i += 32)
                tma_load_3d_2sm(sb + i * 128, &tmB, kb * GEMM_BLOCK_K, tw.col0 + crank * b_rows + i, p.b_batched ? z : 0, &full_bar[stage]);
            } else {
              tma_load_3d_2sm(sb, &tmB, kb * GEMM_BLOCK_K, tw.col0 + crank * Cfg::B_ROWS, p.b_batched ? z : 0, &full_bar[stage]);
            }
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_3d(sa, &tmA, kb * GEMM_BLOCK_K, mt * GEMM_BLOCK_M, p.a_batched ? z : 0, &full_bar[stage]);
            if (CLUSTER == 1) {
              tma_load_3d(sb, &tmB, kb * GEMM_BLOCK_K, tw.col0, p.b_batched ? z : 0, &full_bar[stage]);
            } else {
              tma_load_3d_mc(sb + crank * B_SLICE_ROWS * 128, &tmB, kb * GEMM_BLOCK_K, tw.col0 + crank * B_SLICE_ROWS,
                             p.b_batched ? z : 0, &full_bar[stage], CMASK);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: leader CTA only; whole warp converged, one elected lane issues) ============
    if (!PAIR || crank == 0) {
      int stage = 0, tseq = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      TileWalk tw(p, cid, ncl, BLOCK_N);
      for (; tw.next(); ++tseq) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        if (lane == 0) trace_stamp(p, tseq, 1);
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        const uint32_t idesc = (p.idesc & ~(0x3Fu << 17)) | ((uint32_t)(tw.ncols >> 3) << 17);   // MMA N = width of this tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0 && kb == 0) trace_stamp(p, tseq, 2);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_sdesc(sa, 16, 1024, PA_SWZ_128B);
          const uint64_t bdesc = make_sdesc(sa + Cfg::A_BYTES, 16, 1024, PA_SWZ_128B);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              // advance 16 elements (32 bytes) along K inside the 128B swizzle atom: +2 in 16-byte units
              if (PAIR) umma_ss2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              else umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            // frees the smem stage (in every CTA that multicasts into it / is read by the pair MMA) when these MMAs retire
            if (PAIR) umma_commit2_mc(&empty_bar[stage], CMASK);
            else if (CLUSTER == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_mc(&empty_bar[stage], CMASK);
            if (kb == num_kb - 1) {            // accumulator complete (PAIR: both CTAs' epilogues)
              if (PAIR) umma_commit2_mc(&tfull_bar[acc], CMASK);
              else umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (lane == 0) trace_stamp(p, tseq, 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * ESETS) {
    // ===================== epilogue (warps 4-7 [+ 8-11]; wider CTAs of the fused kernels leave the rest idle here) ==
    // several sets exist only on the staged-store path: the fused kernels (LEAN) have no other path, the stand-alone kernel is
    // launched with ESETS > 1 only when the host established p.tma_store (see launch_gemm_cfg)
    static_assert(ESETS == 1 || ESETS == 2 || ESETS == 4, "1, 2 or 4 epilogue sets");
    const int eset = (warp - 4) >> 2;       // which set of four warps
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int trow = q * 32 + lane;         // row inside the tile
    const bool leader = (q == 0 && lane == 0);          // one per set: issues and tracks the set's bulk stores
    const bool tracer = (warp == 4 && lane == 0);
    const int elt = (p.out_dtype == 2) ? 4 : 2;
    const int row_bytes = 32 * elt;         // staged sub-tile row: 128 B (fp32, SW128) or 64 B (16-bit, SW64)
    // the 32 KB staging area is split between the ACTIVE sets; a sub-tile is 8 KB (16-bit) or 16 KB (fp32), so with four sets
    // an fp32 output leaves sets 2-3 idle (they still take part in every barrier / arrival / publication)
    const int active_sets = (ESETS == 4 && elt == 4) ? 2 : ESETS;
    const int region = 2 * GEMM_CSTAGE_BYTES / active_sets;
    const int cbuf_stride = (ESETS == 1) ? GEMM_CSTAGE_BYTES : 128 * row_bytes;
    const bool single_buf = (ESETS > 1) && region < 2 * 128 * row_bytes;
    uint8_t* cbuf = smem + Cfg::C_OFFSET + (eset < active_sets ? eset : 0) * region;
    int acc = 0, cb = 0, tseq = 0;
    uint32_t acc_phase = 0;
    int pending_mt = -1;                     // fused kernels: tile whose stores are in flight and not yet published
    TileWalk tw(p, cid, ncl, BLOCK_N);
    for (; tw.next(); ++tseq) {
      const int z = tw.z;
      const int mt = tw.mg * CLUSTER + crank;
      const int row = mt * GEMM_BLOCK_M + trow;
      const int col0 = tw.col0;
      const bool row_ok = row < p.M;
      // 16-bit residual rows with 16-byte aligned pitch: prefetched one sub-tile ahead (issued before the wait for the accumulator)
      const bool res_fast = !LEAN && p.residual != nullptr && p.res_dtype != 2 && (p.ldr % 8) == 0 && (p.r_batch % 8) == 0 &&
                            (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
      uint4 rpre[4];
      if (!LEAN && res_fast && row_ok && eset < active_sets && col0 + eset * 32 + 32 <= p.N) {
        const uint4* rp4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + (long long)z * p.r_batch +
                                                          (long long)row * p.ldr + col0 + eset * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) rpre[i] = __ldg(rp4 + i);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (p.signal_ctr != nullptr && leader && pending_mt >= 0) {
        if (pending_mt < p.m_tiles) signal_counter(p.signal_ctr + pending_mt);   // its stores were issued a whole mainloop ago
      }
      pending_mt = mt;
      if (tracer) trace_stamp(p, tseq, 4);
      const uint32_t t_base = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16);
      const float bias_m = (!LEAN && p.bias_mode == 2 && row_ok) ? p.bias[row] : 0.f;
      const long long d_off = (long long)z * p.d_batch + (long long)row * p.ldd;
      float* sbias = reinterpret_cast<float*>(smem + Cfg::BIAS_OFFSET);
      if (p.bias_mode == 1) {
        // column biases of this tile: one coalesced load into smem, then broadcast reads (was 32 scalar LDGs per chunk)
        const int et = threadIdx.x - 128;
        if (ESETS > 1) epi_bar_sync_all<ESETS>();          // the other sets have finished reading the previous tile's biases
        if (et < 128)
          for (int i = et; i < tw.ncols; i += 128) sbias[i] = (col0 + i < p.N) ? __ldg(p.bias + col0 + i) : 0.f;
        if (ESETS > 1) epi_bar_sync_all<ESETS>();
        else epi_bar_sync();
      }
#pragma unroll 1
      for (int c = (eset < active_sets ? eset * 32 : tw.ncols); c < tw.ncols; c += 32 * active_sets) {
        const int col = col0 + c;
        if (col >= p.N) break;              // uniform over the epilogue warps
        if (p.debug_flags & 1) break;       // experiment: no epilogue work at all (output garbage)
        uint32_t v[32];
        if (tracer) trace_chunk(p, tseq, c >> 5, 0);
        tmem_ld32(t_base + c, v);
        tmem_ld_wait();
        if (tracer) trace_chunk(p, tseq, c >> 5, 1);
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + bias_m;
        if (p.bias_mode == 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + i);
            f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
          }
        }
        if (!LEAN && p.residual != nullptr && row_ok) {
          const long long r_off = (long long)z * p.r_batch + (long long)row * p.ldr + col;
          if (res_fast && col + 32 <= p.N) {
            // this row's 32 residual values were requested one sub-tile ago (rpre): with 227 KB of shared memory there is no
            // L1 left, so each 16-byte load is an L2 round trip -- in line they cost a biased + residual GEMM half its rate
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t w[4] = {rpre[i].x, rpre[i].y, rpre[i].z, rpre[i].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (p.res_dtype == 0) {
                  const __half2 h = *reinterpret_cast<const __half2*>(&w[k]);
                  f[8 * i + 2 * k] += __low2float(h); f[8 * i + 2 * k + 1] += __high2float(h);
                } else {
                  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
                  f[8 * i + 2 * k] += __low2float(h); f[8 * i + 2 * k + 1] += __high2float(h);
                }
              }
            }
            const int ncol = col + 32 * active_sets;
            if (c + 32 * active_sets < tw.ncols && ncol + 32 <= p.N) {
              const uint4* rp4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + r_off + 32 * active_sets);
#pragma unroll
              for (int i = 0; i < 4; ++i) rpre[i] = __ldg(rp4 + i);
            }
          } else if (p.res_dtype == 2) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + r_off;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __ldg(rp + i) : 0.f;
          } else if (p.res_dtype == 0) {
            const __half* hp = reinterpret_cast<const __half*>(p.residual) + r_off;
            for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __half2float(__ldg(hp + i)) : 0.f;
          } else {
            const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_off;
            for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __bfloat162float(__ldg(bp + i)) : 0.f;
          }
        }
        if (LEAN || p.tma_store) {
          // ---- staged path: this 128 x 32 sub-tile -> swizzled smem -> one TMA store
          uint8_t* buf = cbuf + (single_buf ? 0 : cb * cbuf_stride);
          if (leader) {                                    // buffer cb's previous store has been read
            if (single_buf) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          epi_bar_sync(eset);
          if (tracer) trace_chunk(p, tseq, c >> 5, 2);
          uint8_t* rowp = buf + trow * row_bytes;
          if (p.out_dtype == 2) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              *reinterpret_cast<float4*>(rowp + ((ch ^ (trow & 7)) << 4)) = make_float4(f[4 * ch], f[4 * ch + 1], f[4 * ch + 2], f[4 * ch + 3]);
          } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint4 u;
              if (p.out_dtype == 0) {
                u = make_uint4(pack_h2(f[8 * ch], f[8 * ch + 1]), pack_h2(f[8 * ch + 2], f[8 * ch + 3]),
                               pack_h2(f[8 * ch + 4], f[8 * ch + 5]), pack_h2(f[8 * ch + 6], f[8 * ch + 7]));
              } else {
                u = make_uint4(pack_bf2(f[8 * ch], f[8 * ch + 1]), pack_bf2(f[8 * ch + 2], f[8 * ch + 3]),
                               pack_bf2(f[8 * ch + 4], f[8 * ch + 5]), pack_bf2(f[8 * ch + 6], f[8 * ch + 7]));
              }
              *reinterpret_cast<uint4*>(rowp + ((ch ^ ((trow >> 1) & 3)) << 4)) = u;
            }
          }
          fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the TMA (async proxy)
          epi_bar_sync(eset);
          if (tracer) trace_chunk(p, tseq, c >> 5, 3);
          if (leader) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmD)),
                         "r"(smem_u32(buf)), "r"(col), "r"(mt * GEMM_BLOCK_M), "r"(z)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          cb ^= 1;
        } else {
          // ---- unaligned output pitch (e.g. CvT's NCHW rows of 196): stage the warp's 32 x 32 block in smem (padded rows),
          //      then each warp instruction writes ONE row's 32 contiguous elements (coalesced) instead of 32 scattered ones
          float* wbuf = reinterpret_cast<float*>(cbuf) + q * (32 * 33);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 32; ++i) wbuf[lane * 33 + i] = f[i];
          __syncwarp();
          const int row0 = mt * GEMM_BLOCK_M + q * 32;
          for (int r = 0; r < 32; ++r) {
            if (row0 + r >= p.M) break;
            if (col + lane < p.N) {
              const float val = wbuf[r * 33 + lane];
              const long long off = (long long)z * p.d_batch + (long long)(row0 + r) * p.ldd + col + lane;
              if (p.out_dtype == 2) reinterpret_cast<float*>(p.D)[off] = val;
              else if (p.out_dtype == 0) reinterpret_cast<__half*>(p.D)[off] = __float2half_rn(val);
              else reinterpret_cast<__nv_bfloat16*>(p.D)[off] = __float2bfloat16_rn(val);
            }
          }
        }
      }
      if (p.bias_mode == 1 && !p.tma_store) epi_bar_sync();   // nobody may refill the bias buffer while it is being read
      tc_fence_before();
      __syncwarp();
      if (tracer) trace_stamp(p, tseq, 5);
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (leader) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output bytes committed before exit
      if (p.signal_ctr != nullptr && pending_mt >= 0 && pending_mt < p.m_tiles) signal_counter(p.signal_ctr + pending_mt);
    }
  }

}

// ESETS = 2: warps 8-11 join the epilogue (384 threads).  A 256 x 256 x K pair tile's mainloop is K / 64 x 512 cycles; one set of
// four warps needs ~5 k cycles to drain 256 columns, so for K <= 512 (PVT, CSWin, CvT projections) a single set is the limiter.
template <int BLOCK_N, int STAGES, int CLUSTER, bool PAIR = false, int ESETS = 1>
__global__ void __launch_bounds__(128 + 128 * ESETS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, STAGES, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    gemm_init_barriers<STAGES, CLUSTER, PAIR, ESETS>(bars);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) { tmem_alloc2(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();     // peers' barriers must be initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  gemm_run<BLOCK_N, STAGES, CLUSTER, PAIR, false, ESETS>(tmA, tmB, tmD, p, smem, bars, tmem_base);

  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();     // no CTA may exit while a peer can still multicast into it
  tc_fence_after();
  if (warp == 2) {
    if (PAIR) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace pa
