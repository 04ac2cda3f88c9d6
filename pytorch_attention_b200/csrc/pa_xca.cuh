// pa_xca.cuh — cross-covariance attention core (xcit.py:251-262) on the tensor cores.
//
//   q^, k^ = columns of q, k (one column per channel d, over the N tokens) divided by max(L2 norm, 1e-12)      xcit.py:255-256
//   A[d, e] = softmax_e( temperature[h] * sum_n q^[n, d] k^[n, e] )        per head: hd x hd                    xcit.py:258-259
//   O[n, d] = sum_e A[d, e] v[n, e]                                                                              xcit.py:262
//
// The contraction of the first step runs over the TOKENS, so its operands are the q / k tiles exactly as they lie in the
// [B*N, 3C] qkv buffer (token rows, channel columns) used as MN-major operands: no transpose exists anywhere.
// One unit of work = (image, group of 128 adjacent channels = 2 heads of 64 or 4 heads of 32):
//   phase 1   G = Q2^T K2, Sqq = Q2^T Q2, Skk = K2^T K2       three M=128, N=128 MMAs per 16 tokens into 3 x 128 TMEM columns;
//             the diagonal hd x hd blocks of G are the per-head covariances, the diagonals of Sqq / Skk the squared norms
//             (the off-diagonal blocks pair channels of different heads and are ignored: the MMAs are tiny next to the GEMMs)
//   softmax   thread r owns channel row r: norms, temperature, softmax over its head's hd columns -- all in registers;
//             A goes to shared memory as a block-diagonal 128 x 128 fp16 K-major B operand
//   phase 2   O2[n, 0:128] = V2[n, 0:128] . blockdiag(A)^T     one M=128 (tokens), N=128, K=128 MMA chain per 128 tokens
// Tokens are streamed in chunks of 128 rows through a 2-stage TMA ring, so any N works; rows past N are zero-filled by TMA.
// 192 threads: warp 0 TMA producer | warp 1 MMA issuer (+ TMEM allocator) | warps 2-5 softmax rows, then output rows.
#pragma once
#include "pa_ptx.cuh"

namespace pa {

constexpr int XT_THREADS = 192;
constexpr int XT_PANEL = 128 * 128;                 // one 64-column panel of a 128-token chunk: 128 rows x 128 B (SW128)
constexpr int XT_STAGE = 4 * XT_PANEL;              // Q panels 0,1 | K panels 0,1   (phase 2: V panels 0,1 in the first half)
constexpr int XT_BMAT = 2 * XT_PANEL;               // block-diagonal A: two 64-wide k panels of 128 rows
constexpr int XT_SMEM = 2 * XT_STAGE + XT_BMAT + 1024 /* norms */ + 128 /* barriers */ + 1024 /* alignment */;

struct XcaTcParams {
  int B, N, C, H, groups, units, nchunks;
  const float* temperature;
  void* out;                 // [B*N, C] fp16
  uint32_t idesc_mn, idesc_k;
};

template <int HD>
__global__ void __launch_bounds__(XT_THREADS, 1)
xca_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const XcaTcParams p) {
  static_assert(HD == 32 || HD == 64, "XCA core: 32- or 64-wide heads");
  extern __shared__ uint8_t xt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xt_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bmat = smem + 2 * XT_STAGE;
  float* xnk = reinterpret_cast<float*>(bmat + XT_BMAT);        // squared k norms of the unit's 128 channels
  uint64_t* bars = reinterpret_cast<uint64_t*>(bmat + XT_BMAT + 1024);
  uint64_t* full = bars;            // [2] TMA -> MMA
  uint64_t* empty = bars + 2;       // [2] MMA -> TMA
  uint64_t* g_full = bars + 4;      // phase-1 accumulators complete
  uint64_t* b_ready = bars + 5;     // 4 warps: A staged in smem, G / Sqq / Skk read
  uint64_t* o_full = bars + 6;      // O tile complete
  uint64_t* o_empty = bars + 7;     // 4 warps: O tile in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], (i == 5 || i == 7) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_G = tmem_base, T_QQ = tmem_base + 128, T_KK = tmem_base + 256, T_O = tmem_base + 384;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int sc = 0;                       // stage use counter (both phases)
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int g = u % p.groups, b = u / p.groups;
      const int c0 = g * 128;
      for (int ph = 0; ph < 2; ++ph) {
        for (int c = 0; c < p.nchunks; ++c, ++sc) {
          const int s = sc & 1;
          uint8_t* st = smem + s * XT_STAGE;
          mbar_wait(&empty[s], ((sc >> 1) & 1) ^ 1);
          if (elect_one()) {
            if (ph == 0) {
              mbar_expect_tx(&full[s], 4 * XT_PANEL);
              tma_load_3d(st, &tmQKV, c0, c * 128, b, &full[s]);
              tma_load_3d(st + XT_PANEL, &tmQKV, c0 + 64, c * 128, b, &full[s]);
              tma_load_3d(st + 2 * XT_PANEL, &tmQKV, p.C + c0, c * 128, b, &full[s]);
              tma_load_3d(st + 3 * XT_PANEL, &tmQKV, p.C + c0 + 64, c * 128, b, &full[s]);
            } else {
              mbar_expect_tx(&full[s], 2 * XT_PANEL);
              tma_load_3d(st, &tmQKV, 2 * p.C + c0, c * 128, b, &full[s]);
              tma_load_3d(st + XT_PANEL, &tmQKV, 2 * p.C + c0 + 64, c * 128, b, &full[s]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t sbase = smem_u32(smem), bbase = smem_u32(bmat);
    int sc = 0, ui = 0, oc = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++ui) {
      // ---- phase 1: token-contracted Gram matrices (MN-major operands: 64-channel panels XT_PANEL apart = LBO, 8-token groups 1024 B apart = SBO)
      for (int c = 0; c < p.nchunks; ++c, ++sc) {
        const int s = sc & 1;
        mbar_wait(&full[s], (sc >> 1) & 1);
        tc_fence_after();
        const uint32_t st = sbase + s * XT_STAGE;
        const uint64_t qd = make_sdesc(st, XT_PANEL, 1024, PA_SWZ_128B);
        const uint64_t kd = make_sdesc(st + 2 * XT_PANEL, XT_PANEL, 1024, PA_SWZ_128B);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {            // 16 tokens per MMA: + 16 rows x 128 B = 128 sixteen-byte units
            const uint32_t acc = (c | k) != 0;
            umma_ss(T_G, qd + 128 * k, kd + 128 * k, p.idesc_mn, acc);
            umma_ss(T_QQ, qd + 128 * k, qd + 128 * k, p.idesc_mn, acc);
            umma_ss(T_KK, kd + 128 * k, kd + 128 * k, p.idesc_mn, acc);
          }
          umma_commit(&empty[s]);
          if (c == p.nchunks - 1) umma_commit(g_full);
        }
        __syncwarp();
      }
      // ---- phase 2: O tile = V2 chunk (K-major A) x blockdiag(A)^T (K-major B, written by the softmax threads)
      mbar_wait(b_ready, ui & 1);
      tc_fence_after();
      for (int c = 0; c < p.nchunks; ++c, ++sc, ++oc) {
        const int s = sc & 1;
        mbar_wait(&full[s], (sc >> 1) & 1);
        mbar_wait(o_empty, (oc & 1) ^ 1);
        tc_fence_after();
        const uint32_t st = sbase + s * XT_STAGE;
        if (elect_one()) {
#pragma unroll
          for (int pn = 0; pn < 2; ++pn) {
            const uint64_t vd = make_sdesc(st + pn * XT_PANEL, 16, 1024, PA_SWZ_128B);
            const uint64_t bd = make_sdesc(bbase + pn * XT_PANEL, 16, 1024, PA_SWZ_128B);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(T_O, vd + 2 * k, bd + 2 * k, p.idesc_k, (pn | k) != 0);
          }
          umma_commit(&empty[s]);
          umma_commit(o_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== 128 worker threads: channel rows in the softmax, token rows in the epilogue =====================
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                     // row: channel of the group / token of the chunk
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int blk0 = (r / HD) * HD;                  // first channel of this row's head inside the group
    int ui = 0, oc = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++ui) {
      const int g = u % p.groups, b = u / p.groups;
      const int ch = g * 128 + r;                    // global channel of this row
      mbar_wait(g_full, ui & 1);
      tc_fence_after();
      // squared norms: the diagonal entry of this row inside its head's block
      float nq2 = 0.f, nk2 = 0.f;
      {
        uint32_t v[HD];
        if (HD == 64) tmem_ld64(T_QQ + lane_off + blk0, *reinterpret_cast<uint32_t(*)[64]>(&v[0]));
        else tmem_ld32(T_QQ + lane_off + blk0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HD; ++i) nq2 = (i == r - blk0) ? __uint_as_float(v[i]) : nq2;
        if (HD == 64) tmem_ld64(T_KK + lane_off + blk0, *reinterpret_cast<uint32_t(*)[64]>(&v[0]));
        else tmem_ld32(T_KK + lane_off + blk0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HD; ++i) nk2 = (i == r - blk0) ? __uint_as_float(v[i]) : nk2;
      }
      xnk[r] = nk2;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float a[HD];
      {
        uint32_t v[HD];
        if (HD == 64) tmem_ld64(T_G + lane_off + blk0, *reinterpret_cast<uint32_t(*)[64]>(&v[0]));
        else tmem_ld32(T_G + lane_off + blk0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_wait();
        const float temp = (ch < p.C) ? __ldg(p.temperature + ch / HD) : 1.f;
        const float sq = temp * 1.4426950408889634f / fmaxf(sqrtf(nq2), 1e-12f);   // F.normalize: x / max(||x||, eps)
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < HD; ++i) {
          a[i] = __uint_as_float(v[i]) * sq / fmaxf(sqrtf(xnk[blk0 + i]), 1e-12f);   // log2-domain logits
          mx = fmaxf(mx, a[i]);
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < HD; ++i) { a[i] = ex2f(a[i] - mx); sum += a[i]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int i = 0; i < HD; ++i) a[i] *= inv;
      }
      // block-diagonal row r of the B operand: [128 rows][2 panels of 64 k], 16-byte chunk c of a row at (c ^ (row & 7)) (SW128)
      {
        const int pnz = blk0 >> 6;                       // panel that holds this head's columns
        const int choff = (blk0 & 63) >> 3;              // first 16-byte chunk of the head inside that panel (0, or 4 for hd 32)
#pragma unroll
        for (int pn = 0; pn < 2; ++pn) {
          uint8_t* rowp = bmat + pn * XT_PANEL + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 val = make_uint4(0, 0, 0, 0);
            if (pn == pnz) {
              if (HD == 64) {
                val = make_uint4(pack_h2(a[8 * c], a[8 * c + 1]), pack_h2(a[8 * c + 2], a[8 * c + 3]),
                                 pack_h2(a[8 * c + 4], a[8 * c + 5]), pack_h2(a[8 * c + 6], a[8 * c + 7]));
              } else {
                const int j = c & 3;                      // static after unrolling
                if ((c >> 2) * 4 == choff)
                  val = make_uint4(pack_h2(a[(8 * j) % HD], a[(8 * j + 1) % HD]), pack_h2(a[(8 * j + 2) % HD], a[(8 * j + 3) % HD]),
                                   pack_h2(a[(8 * j + 4) % HD], a[(8 * j + 5) % HD]), pack_h2(a[(8 * j + 6) % HD], a[(8 * j + 7) % HD]));
              }
            }
            *reinterpret_cast<uint4*>(rowp + ((c ^ (r & 7)) << 4)) = val;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ready);
      // ---- output tiles: this thread = token row r of the chunk
      for (int c = 0; c < p.nchunks; ++c, ++oc) {
        mbar_wait(o_full, oc & 1);
        tc_fence_after();
        const int tok = c * 128 + r;
        uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.N + tok) * p.C + g * 128;
#pragma unroll 1
        for (int cc = 0; cc < 128; cc += 32) {
          uint32_t o[32];
          tmem_ld32(T_O + lane_off + cc, o);
          tmem_ld_wait();
          if (cc == 96) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
          }
          if (tok < p.N && g * 128 + cc < p.C) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + cc);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              d4[i] = make_uint4(pack_h2(__uint_as_float(o[8 * i]), __uint_as_float(o[8 * i + 1])),
                                 pack_h2(__uint_as_float(o[8 * i + 2]), __uint_as_float(o[8 * i + 3])),
                                 pack_h2(__uint_as_float(o[8 * i + 4]), __uint_as_float(o[8 * i + 5])),
                                 pack_h2(__uint_as_float(o[8 * i + 6]), __uint_as_float(o[8 * i + 7])));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace pa
