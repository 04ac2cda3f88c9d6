// pa_gemm.cuh — persistent warp-specialised tcgen05 GEMM for the projection stages.
//
//   D[z][m, n] = sum_k A[z][m, k] * B[z][n, k]  (+ bias)      A, B K-major ("TN", nn.Linear layout)
//
// One CTA per SM, 256 threads:
//   warp 0   TMA producer      (cp.async.bulk.tensor, 128B-swizzled 64-wide K blocks, STAGES-deep ring)
//   warp 1   MMA issuer        (one thread: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue         (tcgen05.ld 32x32b -> bias -> convert -> global), double-buffered TMEM accumulator
// Tiles are distributed round-robin over the persistent grid; rows beyond M / K are zero-filled by TMA
// (3-D tensor maps {K, rows, Z}) and masked in the epilogue, so M, N need not divide the tile.
#pragma once
#include "pa_ptx.cuh"

namespace pa {

struct GemmParams {
  int M, N, K, Z;          // per-batch problem, Z batches
  int m_tiles, n_tiles;    // ceil(M/128), ceil(N/BLOCK_N)
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
  uint32_t idesc;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 256;

template <int BLOCK_N, int STAGES>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;        // + barriers + 1024-alignment slack
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;    // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int tiles_per_z = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_per_z * p.Z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int z = tile / tiles_per_z;
        const int r = tile - z * tiles_per_z;
        const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          tma_load_3d(sa, &tmA, kb * GEMM_BLOCK_K, mt * GEMM_BLOCK_M, p.a_batched ? z : 0, &full_bar[stage]);
          tma_load_3d(sb, &tmB, kb * GEMM_BLOCK_K, nt * BLOCK_N, p.b_batched ? z : 0, &full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_sdesc(sa, 16, 1024, PA_SWZ_128B);
          const uint64_t bdesc = make_sdesc(sa + Cfg::A_BYTES, 16, 1024, PA_SWZ_128B);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128B swizzle atom: +2 in 16-byte units
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);       // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int z = tile / tiles_per_z;
      const int r = tile - z * tiles_per_z;
      const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      const int row = mt * GEMM_BLOCK_M + q * 32 + lane;
      const int col0 = nt * BLOCK_N;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16);
      const bool row_ok = row < p.M;
      const float bias_m = (p.bias_mode == 2 && row_ok) ? p.bias[row] : 0.f;
      const long long d_off = (long long)z * p.d_batch + (long long)row * p.ldd;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_base + c, v);
        tmem_ld_wait();
        const int col = col0 + c;
        if (col >= p.N) break;              // warp-uniform
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + bias_m;
        if (p.bias_mode == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __ldg(p.bias + col + i) : 0.f;
        }
        if (p.residual != nullptr && row_ok) {
          const long long r_off = (long long)z * p.r_batch + (long long)row * p.ldr + col;
          if (p.res_dtype == 2) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + r_off;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __ldg(rp + i) : 0.f;
          } else if (p.res_dtype == 0) {
            const __half* rp = reinterpret_cast<const __half*>(p.residual) + r_off;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __half2float(__ldg(rp + i)) : 0.f;
          } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_off;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += (col + i < p.N) ? __bfloat162float(__ldg(rp + i)) : 0.f;
          }
        }
        if (row_ok) {
          const bool full = (col + 32 <= p.N);
          if (p.out_dtype == 2) {
            float* d = reinterpret_cast<float*>(p.D) + d_off + col;
            if (full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(d + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            } else {
              for (int i = 0; i < 32; ++i)
                if (col + i < p.N) d[i] = f[i];
            }
          } else {
            uint32_t h[16];
            if (p.out_dtype == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) h[i] = pack_h2(f[2 * i], f[2 * i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) h[i] = pack_bf2(f[2 * i], f[2 * i + 1]);
            }
            uint16_t* d = reinterpret_cast<uint16_t*>(p.D) + d_off + col;
            if (full && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<uint4*>(d + 2 * i) = make_uint4(h[i], h[i + 1], h[i + 2], h[i + 3]);
            } else {
              for (int i = 0; i < 32; ++i)
                if (col + i < p.N) d[i] = (uint16_t)((i & 1) ? (h[i >> 1] >> 16) : (h[i >> 1] & 0xFFFF));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace pa
