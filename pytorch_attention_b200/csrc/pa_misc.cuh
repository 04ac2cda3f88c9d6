// pa_misc.cuh — the HBM-bound pieces around the tensor-core stages: depthwise convolutions with folded eval-mode
// BatchNorm (PVT spatial reduction, CvT projection front-end), LayerNorm (CSWin), LePE (CSWin), the small
// channel-attention cores of XCiT, and a row copy.  All are coalesced / 16-byte vectorised along the channel axis.
#pragma once
#include "pa_ptx.cuh"

namespace pa {

// 16-bit element helpers (dtype: 0 fp16, 1 bf16)
__device__ __forceinline__ float ld16(const void* p, long long i, int dtype) {
  return dtype == 0 ? __half2float(reinterpret_cast<const __half*>(p)[i])
                    : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void unpack8(const uint4& u, int dtype, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dtype == 0) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      f[2 * i] = __low2float(h);
      f[2 * i + 1] = __high2float(h);
    } else {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      f[2 * i] = __low2float(h);
      f[2 * i + 1] = __high2float(h);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int dtype) {
  uint4 u;
  if (dtype == 0) {
    u.x = pack_h2(f[0], f[1]); u.y = pack_h2(f[2], f[3]); u.z = pack_h2(f[4], f[5]); u.w = pack_h2(f[6], f[7]);
  } else {
    u.x = pack_bf2(f[0], f[1]); u.y = pack_bf2(f[2], f[3]); u.z = pack_bf2(f[4], f[5]); u.w = pack_bf2(f[6], f[7]);
  }
  return u;
}

// ------------------------------------------------------------------------------------------------
// PVT spatial reduction (pvt.py:67-71, 77-78): depthwise conv k = stride = sr over the token image, eval BatchNorm
// folded into (scale, shift):  out[b, i*Ws+j, c] = scale[c] * sum_{u,v} w[(u*sr+v), c] x[b, (sr*i+u)*W + sr*j+v, c] + shift[c]
// One thread = 8 consecutive channels of one output token (16-byte loads, coalesced along C).
struct SrParams {
  const void* x; void* out;        // x [B, H*W, C] (dtype), out [B, Hs*Ws, C] fp16
  const float* w;                  // [sr*sr, C] fp32 (transposed depthwise weight)
  const float* scale; const float* shift;   // [C] fp32
  int B, H, W, C, sr, Hs, Ws, dtype;
};
__global__ void sr_conv_bn_kernel(const SrParams p) {
  const int cvec = p.C / 8;
  const long long total = (long long)p.B * p.Hs * p.Ws * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    long long t = idx / cvec;
    const int j = (int)(t % p.Ws); t /= p.Ws;
    const int i = (int)(t % p.Hs);
    const int b = (int)(t / p.Hs);
    const int c0 = cv * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int u = 0; u < p.sr; ++u) {
      const long long rowbase = ((long long)b * p.H * p.W + (long long)(p.sr * i + u) * p.W + p.sr * j) * p.C + c0;
      for (int v = 0; v < p.sr; ++v) {
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.x) + rowbase + (long long)v * p.C));
        float xf[8];
        unpack8(xv, p.dtype, xf);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + (long long)(u * p.sr + v) * p.C + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + (long long)(u * p.sr + v) * p.C + c0 + 4));
        acc[0] = fmaf(xf[0], w0.x, acc[0]); acc[1] = fmaf(xf[1], w0.y, acc[1]);
        acc[2] = fmaf(xf[2], w0.z, acc[2]); acc[3] = fmaf(xf[3], w0.w, acc[3]);
        acc[4] = fmaf(xf[4], w1.x, acc[4]); acc[5] = fmaf(xf[5], w1.y, acc[5]);
        acc[6] = fmaf(xf[6], w1.z, acc[6]); acc[7] = fmaf(xf[7], w1.w, acc[7]);
      }
    }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(acc[k], __ldg(p.scale + c0 + k), __ldg(p.shift + c0 + k));
    const long long ob = ((long long)b * p.Hs * p.Ws + (long long)i * p.Ws + j) * p.C + c0;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + ob) = pack8(o, 0);
  }
}

// (measured and rejected, late round 2: one block per output token with (channel groups) x (tap rows) threads and the sr partial
//  sums meeting in shared memory -- every load of a token in flight at once, 2048 blocks x 512 threads at config 3 -- is slower
//  than the one-thread-per-output loop above: C3 forward 255 us against 248.)
// P2T pooling pyramid (p2t.py:78-86), kernel 1: adaptive_avg_pool2d of the token map at up to four output sizes.  Output
// position (pi, pj) of level l averages rows [floor(pi*H/ph), ceil((pi+1)*H/ph)) x cols [floor(pj*W/pw), ceil((pj+1)*W/pw))
// (ATen's adaptive pooling windows).  One thread = 8 consecutive channels of one pooled token, fp32 out [B, M, C].
struct P2tPoolParams {
  const void* x; float* pooled;
  int B, H, W, C, dtype, n_levels, M;
  int ph[4], pw[4], off[4];        // level sizes and first token index of each level
};
__global__ void p2t_pool_kernel(const P2tPoolParams p) {
  const int cvec = p.C / 8;
  const long long total = (long long)p.B * p.M * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    long long t = idx / cvec;
    const int tok = (int)(t % p.M);
    const int b = (int)(t / p.M);
    int l = 0;
    while (l + 1 < p.n_levels && tok >= p.off[l + 1]) ++l;
    const int r = tok - p.off[l], pi = r / p.pw[l], pj = r - pi * p.pw[l];
    const int h0 = (pi * p.H) / p.ph[l], h1 = ((pi + 1) * p.H + p.ph[l] - 1) / p.ph[l];
    const int w0 = (pj * p.W) / p.pw[l], w1 = ((pj + 1) * p.W + p.pw[l] - 1) / p.pw[l];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int hh = h0; hh < h1; ++hh)
      for (int ww = w0; ww < w1; ++ww) {
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.x) +
                                                              ((long long)b * p.H * p.W + (long long)hh * p.W + ww) * p.C + cv * 8));
        float xf[8];
        unpack8(xv, p.dtype, xf);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += xf[k];
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
    float4* dst = reinterpret_cast<float4*>(p.pooled + ((long long)b * p.M + tok) * p.C + cv * 8);
    dst[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    dst[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}
// kernel 2: pool + depthwise3x3(pool) (+ bias, zero pad inside the level's own map), then LayerNorm over the channels
// (two-pass statistics like layernorm_kernel), fp16 tokens [B, M, C].  One warp per pooled token; the maps are tiny.
struct P2tTokParams {
  const float* pooled; void* out;
  const float* w[4]; const float* bias[4];     // per level [9, C] / [C]
  const float* gamma; const float* beta; float eps;
  int B, C, n_levels, M;
  int ph[4], pw[4], off[4];
};
__device__ __forceinline__ float p2t_token_value(const P2tTokParams& p, const float* base, int l, int pi, int pj, int c) {
  float v = base[((long long)pi * p.pw[l] + pj) * p.C + c];
  float conv = p.bias[l] != nullptr ? __ldg(p.bias[l] + c) : 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int y = pi + dy, x = pj + dx;
      if (y >= 0 && y < p.ph[l] && x >= 0 && x < p.pw[l])
        conv = fmaf(__ldg(p.w[l] + ((dy + 1) * 3 + dx + 1) * p.C + c), base[((long long)y * p.pw[l] + x) * p.C + c], conv);
    }
  return v + conv;                    // pool + l(pool)   (p2t.py:82)
}
__global__ void p2t_tokens_kernel(const P2tTokParams p) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)p.B * p.M) return;
  const int tok = (int)(wid % p.M), b = (int)(wid / p.M);
  int l = 0;
  while (l + 1 < p.n_levels && tok >= p.off[l + 1]) ++l;
  const int r = tok - p.off[l], pi = r / p.pw[l], pj = r - pi * p.pw[l];
  const float* base = p.pooled + ((long long)b * p.M + p.off[l]) * p.C;     // this level's map of image b, [ph*pw, C]
  float s = 0.f;
  for (int c = lane; c < p.C; c += 32) s += p2t_token_value(p, base, l, pi, pj, c);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / p.C;
  float v = 0.f;
  for (int c = lane; c < p.C; c += 32) { const float d = p2t_token_value(p, base, l, pi, pj, c) - mean; v = fmaf(d, d, v); }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / p.C + p.eps);
  __half* orow = reinterpret_cast<__half*>(p.out) + ((long long)b * p.M + tok) * p.C;
  for (int c = lane; c < p.C; c += 32)
    orow[c] = __float2half_rn(fmaf((p2t_token_value(p, base, l, pi, pj, c) - mean) * rstd, __ldg(p.gamma + c), __ldg(p.beta + c)));
}

// kvt.KNNAttention (kvt.py:84-87): per query row the k-th largest raw score q.k -- the threshold below which the attention kernel
// masks scores.  One warp per (image, head, row): lane l holds the scores of keys l, l+32, ... (fp32 dot products of the fp16
// q / k rows in the qkv buffer), then an exact radix select over the order-preserving integer image of the floats: 32 rounds of
// "how many scores are >= candidate" (per-lane count + warp sum).  SIMT on purpose: 2 N hd flops per row are nothing next to the
// projections, and the selection is integer work.  A row whose k-th and (k+1)-th scores differ by less than fp32 rounding may keep
// k-1 or k+1 entries (as two PyTorch backends may).
struct KnnParams {
  const void* qkv;                 // fp16 [G, N, ld]: q at column q_col0 + h*hd, k at k_col0 + h*hd
  float* thresh;                   // [G, H, N]
  int G, H, N, hd, topk;
  long long ld, group;             // row / group pitch in elements
  int q_col0, k_col0;
};
template <int MAXK>                // keys per lane: N <= 32 * MAXK
__global__ void knn_threshold_kernel(const KnnParams p) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)p.G * p.H * p.N) return;
  const int row = (int)(wid % p.N);
  const int h = (int)((wid / p.N) % p.H);
  const int g = (int)(wid / ((long long)p.N * p.H));
  const __half* base = reinterpret_cast<const __half*>(p.qkv) + (long long)g * p.group;
  const __half* qr = base + (long long)row * p.ld + p.q_col0 + h * p.hd;
  uint32_t key[MAXK];
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    const int kidx = lane + 32 * j;
    float s = 0.f;
    if (kidx < p.N) {
      const __half* kr = base + (long long)kidx * p.ld + p.k_col0 + h * p.hd;
      for (int d = 0; d < p.hd; d += 8) {
        const uint4 qa = __ldg(reinterpret_cast<const uint4*>(qr + d)), ka = __ldg(reinterpret_cast<const uint4*>(kr + d));
        const __half2* q2 = reinterpret_cast<const __half2*>(&qa);
        const __half2* k2 = reinterpret_cast<const __half2*>(&ka);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __half22float2(q2[i]), b = __half22float2(k2[i]);
          s = fmaf(a.x, b.x, s);
          s = fmaf(a.y, b.y, s);
        }
      }
      uint32_t u = __float_as_uint(s);
      key[j] = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);       // ascending unsigned order == ascending float order
    } else {
      key[j] = 0u;                                                   // below every real score
    }
  }
  uint32_t res = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = res | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) cnt += (key[j] >= cand) ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (cnt >= p.topk) res = cand;                                   // at least k scores are >= cand: the k-th largest is too
  }
  // The attention kernel's tensor-core scores accumulate in another order than the dot products above: compared against the k-th
  // largest value ITSELF, the k-th entry would fall below its own threshold in half of the rows (measured: 46 % of the rows kept
  // k - 1 entries).  The threshold handed over is the midpoint between the k-th largest score and the next smaller one.
  uint32_t lower = 0u;                                               // largest key below res; 0 = none (real keys are never 0)
#pragma unroll
  for (int j = 0; j < MAXK; ++j) lower = (key[j] < res && key[j] > lower) ? key[j] : lower;
#pragma unroll
  for (int o = 16; o; o >>= 1) { const uint32_t other = __shfl_xor_sync(0xffffffffu, lower, o); lower = other > lower ? other : lower; }
  if (lane == 0) {
    const float kth = __uint_as_float(res ^ ((res >> 31) ? 0x80000000u : 0xFFFFFFFFu));
    const float nxt = __uint_as_float(lower ^ ((lower >> 31) ? 0x80000000u : 0xFFFFFFFFu));
    p.thresh[wid] = lower != 0u ? 0.5f * (kth + nxt) : -INFINITY;    // topk == N: every score is kept
  }
}

// SegFormer spatial reduction (segformer.py:27, 38-39): a DENSE conv with k = stride = sr is a GEMM over non-overlapping
// patches.  This kernel lays the patches out as that GEMM's K-major A operand -- a pure re-partition of x (kernel == stride:
// every element of x moves exactly once):  out[b, i*Ws+j, (u*sr+v)*C + c] = x[b, (sr*i+u)*W + sr*j+v, c]
// One thread = 8 consecutive channels (16 bytes) of one (output token, tap).
struct PatchParams {
  const void* x; void* out;        // x [B, H*W, C], out [B*Hs*Ws, sr*sr*C], same 16-bit dtype
  int B, H, W, C, sr, Hs, Ws;
};
__global__ void sr_patchify_kernel(const PatchParams p) {
  const int cvec = p.C / 8, taps = p.sr * p.sr;
  const long long total = (long long)p.B * p.Hs * p.Ws * taps * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    long long t = idx / cvec;
    const int tap = (int)(t % taps); t /= taps;
    const int j = (int)(t % p.Ws); t /= p.Ws;
    const int i = (int)(t % p.Hs);
    const int b = (int)(t / p.Hs);
    const int u = tap / p.sr, v = tap - u * p.sr;
    const long long src = ((long long)b * p.H * p.W + (long long)(p.sr * i + u) * p.W + p.sr * j + v) * p.C + cv * 8;
    reinterpret_cast<uint4*>(p.out)[idx] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.x) + src));
  }
}

// ------------------------------------------------------------------------------------------------
// CvT front-end (cvt.py:55-57, 66): depthwise ks x ks conv (stride 1, zero pad) + eval BatchNorm on an NCHW map,
// written TOKEN-major [B, H*W, C] fp16 so the 1x1 qkv conv becomes a plain K-major GEMM.
// Block = (32 channels) x (32 pixels): NCHW reads are coalesced along pixels, the smem transpose makes the
// token-major writes coalesced along channels.
struct DwParams {
  const void* x; void* out;        // x [B, C, H, W] (dtype), out [B, H*W, C] fp16
  const float* w;                  // [C, ks*ks] fp32
  const float* scale; const float* shift;   // folded conv-bias + BN: y = scale*conv + shift
  int B, C, H, W, ks, dtype;
};
__global__ void dwconv_bn_to_tokens_kernel(const DwParams p) {
  __shared__ float tile[32][33];
  const int HW = p.H * p.W;
  const int pix0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8 threads
  const int pad = (p.ks - 1) / 2;
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc, pix = pix0 + tx;
    float acc = 0.f;
    if (c < p.C && pix < HW) {
      const int r = pix / p.W, q = pix % p.W;
      const long long base = ((long long)b * p.C + c) * HW;
      for (int u = 0; u < p.ks; ++u) {
        const int rr = r + u - pad;
        if (rr < 0 || rr >= p.H) continue;
        for (int v = 0; v < p.ks; ++v) {
          const int qq = q + v - pad;
          if (qq < 0 || qq >= p.W) continue;
          acc = fmaf(ld16(p.x, base + (long long)rr * p.W + qq, p.dtype), __ldg(p.w + (long long)c * p.ks * p.ks + u * p.ks + v), acc);
        }
      }
      acc = fmaf(acc, __ldg(p.scale + c), __ldg(p.shift + c));
    }
    tile[cc][tx] = acc;
  }
  __syncthreads();
  for (int pp = ty; pp < 32; pp += 8) {
    const int pix = pix0 + pp, c = c0 + tx;
    if (pix < HW && c < p.C)
      reinterpret_cast<__half*>(p.out)[((long long)b * HW + pix) * p.C + c] = __float2half_rn(tile[tx][pp]);
  }
}

// ks = 3 fast path: a block owns (image, 32 channels, band of DW3_RB rows).  The band plus its halo rows is staged once in
// shared memory as fp32 (per-channel stride padded to an odd word count: lane = channel reads are conflict-free), every
// thread keeps its channel's nine tap weights and the folded BN in registers, and a warp writes the 32 channels of one
// token as one 64-byte segment.  (The generic kernel above issues one scalar 2-byte load per tap: 70 us for 19 MB of traffic.)
constexpr int DW3_RB = 4;
__global__ void __launch_bounds__(256) dwconv3_bn_to_tokens_kernel(const DwParams p) {
  extern __shared__ float dw_tile[];                  // [32 channels][stride]
  const int W = p.W, H = p.H, HW = H * W;
  const int b = blockIdx.x, c0 = blockIdx.y * 32, row0 = blockIdx.z * DW3_RB;
  const int nrows = DW3_RB + 2;
  const int stride = (nrows * W) | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // stage rows row0-1 .. row0+RB (zero outside the image): lanes run along the contiguous pixels of one channel
  for (int cc = warp; cc < 32; cc += 8) {
    const int c = c0 + cc;
    const long long base = ((long long)b * p.C + c) * HW;
    for (int i = lane; i < nrows * W; i += 32) {
      const int rr = row0 - 1 + i / W;
      float v = 0.f;
      if (c < p.C && rr >= 0 && rr < H) v = ld16(p.x, base + (long long)rr * W + (i % W), p.dtype);
      dw_tile[cc * stride + i] = v;
    }
  }
  const int c = c0 + lane;
  float w[9], sc = 0.f, sh = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = (c < p.C) ? __ldg(p.w + (long long)c * 9 + t) : 0.f;
  if (c < p.C) { sc = __ldg(p.scale + c); sh = __ldg(p.shift + c); }
  __syncthreads();
  const float* tc = dw_tile + lane * stride;
  const int npix = min(DW3_RB, H - row0) * W;
  for (int pp = warp; pp < npix; pp += 8) {
    const int r = pp / W, q = pp - r * W;              // r: row inside the band; staged row index r + 1
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int qq = q + v - 1;
        if (qq >= 0 && qq < W) acc = fmaf(tc[(r + u) * W + qq], w[u * 3 + v], acc);     // rows outside the image were staged as zeros
      }
    }
    if (c < p.C)
      reinterpret_cast<__half*>(p.out)[((long long)b * HW + (long long)(row0 + r) * W + q) * p.C + c] = __float2half_rn(fmaf(acc, sc, sh));
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel axis (CSWinBlock.norm1, cswin.py:184): one warp per token row, fp32 statistics,
// two-pass (mean, then centred variance) like ATen's layer_norm.  Output fp16.
struct LnParams {
  const void* x; void* out; const float* gamma; const float* beta;
  long long rows; int C, dtype; float eps;
};
// Rows of at most NJ x 256 channels (NJ = 1..4): the row is read ONCE, as NJ 16-byte loads per lane that are all issued before
// anything is reduced, and kept in NJ x 8 registers.  One instantiation per NJ keeps the register count at what the width needs
// (the generic kernel below reserves 4 x 8 values and ran at 41 % of the warp slots / 39 % of the DRAM rate at C = 512).
template <int NJ>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= p.rows) return;
  const uint16_t* xr = reinterpret_cast<const uint16_t*>(p.x) + row * p.C;
  uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + row * p.C;
  uint4 raw[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = (j * 32 + lane) * 8;
    raw[j] = (c < p.C) ? __ldg(reinterpret_cast<const uint4*>(xr + c)) : make_uint4(0, 0, 0, 0);
  }
  float f[NJ][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    unpack8(raw[j], p.dtype, f[j]);
    if ((j * 32 + lane) * 8 < p.C) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[j][k];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / p.C;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if ((j * 32 + lane) * 8 < p.C) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v = fmaf(f[j][k] - mean, f[j][k] - mean, v);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / p.C + p.eps);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = (j * 32 + lane) * 8;
    if (c < p.C) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o8[k] = fmaf((f[j][k] - mean) * rstd, gg[k], bb[k]);
      *reinterpret_cast<uint4*>(orow + c) = pack8(o8, 0);
    }
  }
}

__global__ void layernorm_kernel(const LnParams p) {
  // one warp per row; the row is read ONCE into registers (up to 4 x 8 elements per lane = C <= 1024), longer rows
  // fall back to re-reading.  Statistics in fp32, two-pass (mean, then centred variance) like ATen's layer_norm.
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= p.rows) return;
  const uint16_t* xr = reinterpret_cast<const uint16_t*>(p.x) + row * p.C;
  uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + row * p.C;
  if (p.C <= 1024) {
    float f[4][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j * 32 + lane) * 8;
      if (c < p.C) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(xr + c)), p.dtype, f[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += f[j][k];
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / p.C;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j * 32 + lane) * 8;
      if (c < p.C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v = fmaf(f[j][k] - mean, f[j][k] - mean, v);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / p.C + p.eps);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (j * 32 + lane) * 8;
      if (c < p.C) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o8[k] = fmaf((f[j][k] - mean) * rstd, gg[k], bb[k]);
        *reinterpret_cast<uint4*>(orow + c) = pack8(o8, 0);
      }
    }
    return;
  }
  float s = 0.f;
  for (int c = lane * 8; c < p.C; c += 256) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(xr + c)), p.dtype, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / p.C;
  float v = 0.f;
  for (int c = lane * 8; c < p.C; c += 256) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(xr + c)), p.dtype, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) v = fmaf(f[k] - mean, f[k] - mean, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / p.C + p.eps);
  for (int c = lane * 8; c < p.C; c += 256) {
    float f[8], o8[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(xr + c)), p.dtype, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) o8[k] = fmaf((f[k] - mean) * rstd, __ldg(p.gamma + c + k), __ldg(p.beta + c + k));
    *reinterpret_cast<uint4*>(orow + c) = pack8(o8, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// ClassAttention core (xcit.py:180-185) for one (batch, head): only the CLS (token 0) query attends.
struct ClsParams {
  const void* qkv; void* out;    // qkv [B, N, 3C] fp16; out [B, C] fp16
  int B, N, C, H; float scale;
};
template <int HD>
__global__ void __launch_bounds__(256) class_attn_core_kernel(const ClsParams p) {
  constexpr int NP = 256 / HD;     // token-strided partial sums per output channel
  extern __shared__ float sc[];    // [N] scores
  __shared__ float red[8];
  __shared__ float q0[HD];
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ld = 3LL * p.C;
  const __half* base = reinterpret_cast<const __half*>(p.qkv) + (long long)b * p.N * ld + h * HD;
  if (tid < HD) q0[tid] = __half2float(base[tid]);
  __syncthreads();
  float mx = -INFINITY;
  for (int n = tid; n < p.N; n += 256) {
    // the token's k row of this head is HD * 2 contiguous, 16-byte aligned bytes: HD / 8 vector loads (was HD scalar ones)
    const uint4* kr4 = reinterpret_cast<const uint4*>(base + (long long)n * ld + p.C);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
      const uint4 u = __ldg(kr4 + j);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s = fmaf(q0[8 * j + 2 * k], __low2float(h2[k]), s);
        s = fmaf(q0[8 * j + 2 * k + 1], __high2float(h2[k]), s);
      }
    }
    s *= p.scale;
    sc[n] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int n = tid; n < p.N; n += 256) { const float e = __expf(sc[n] - mx); sc[n] = e; sum += e; }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.f / sum;
  // cls[d] = sum_n a[n] v[n][d]: thread (part = tid/HD, d = tid%HD) strides over tokens, then reduce the NP parts
  __shared__ float part[NP][HD];
  {
    const int d = tid % HD, pt = tid / HD;
    float o = 0.f;
    for (int n = pt; n < p.N; n += NP) o = fmaf(sc[n], __half2float(base[(long long)n * ld + 2 * p.C + d]), o);
    part[pt][d] = o;
  }
  __syncthreads();
  if (tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) o += part[i][tid];
    reinterpret_cast<__half*>(p.out)[(long long)b * p.C + h * HD + tid] = __float2half_rn(o * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// 16-byte vectorised copy (ClassAttention pass-through of the patch tokens, xcit.py:187)
__global__ void copy16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __ldg(src + i);
}

// ------------------------------------------------------------------------------------------------
// fp32 -> fp16 / bf16 cast of an activation tensor (opt-in fp32-input mode of the drop-ins: the reference's forward takes
// fp32 tensors, ViT.py:79): 32 bytes in, 16 bytes out per thread step, grid sized in multiples of the SM count
__global__ void cast_f32_to_16_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, long long n8, int to_bf16,
                                      const float* __restrict__ tail_src, uint16_t* __restrict__ tail_dst, int tail) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);
    dst[i] = to_bf16 ? make_uint4(pack_bf2(a.x, a.y), pack_bf2(a.z, a.w), pack_bf2(b.x, b.y), pack_bf2(b.z, b.w))
                     : make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
    const float v = tail_src[threadIdx.x];
    tail_dst[threadIdx.x] = to_bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
  }
}

// ------------------------------------------------------------------------------------------------
// LePE (cswin.py:86-99): depthwise 3x3 over each cross-shaped window of V, zero padded AT WINDOW BORDERS,
// written (fp16) into the attention output buffer at the image position of the token; the attention epilogue
// then adds its softmax(QK^T)V on top.  v: column slice [v_col0, v_col0+Cb) of the [B, L, ldv] qkv buffer.
struct LepeParams {
  const void* v; void* out;       // out [B, L, ldo] fp16 at column o_col0
  const float* w; const float* bias;   // [9, Cb] fp32 (transposed), [Cb]
  long long ldv, ldo; int v_col0, o_col0;
  int B, R, Cb, H_sp, W_sp;       // R = resolution (H = W = R)
};
__global__ void lepe_kernel(const LepeParams p) {
  const int cvec = p.Cb / 8;
  const long long total = (long long)p.B * p.R * p.R * cvec;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvec);
    long long t = idx / cvec;
    const int col = (int)(t % p.R); t /= p.R;
    const int row = (int)(t % p.R);
    const int b = (int)(t / p.R);
    const int c0 = cv * 8;
    const int r_in = row % p.H_sp, c_in = col % p.W_sp;     // position inside the window
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = __ldg(p.bias + c0 + k);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int rr = r_in + u - 1;
      if (rr < 0 || rr >= p.H_sp) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int cc = c_in + v - 1;
        if (cc < 0 || cc >= p.W_sp) continue;
        const long long tok = (long long)b * p.R * p.R + (long long)(row + u - 1) * p.R + (col + v - 1);
        float xf[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.v) + tok * p.ldv + p.v_col0 + c0)), 0, xf);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + (u * 3 + v) * p.Cb + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + (u * 3 + v) * p.Cb + c0 + 4));
        acc[0] = fmaf(xf[0], w0.x, acc[0]); acc[1] = fmaf(xf[1], w0.y, acc[1]);
        acc[2] = fmaf(xf[2], w0.z, acc[2]); acc[3] = fmaf(xf[3], w0.w, acc[3]);
        acc[4] = fmaf(xf[4], w1.x, acc[4]); acc[5] = fmaf(xf[5], w1.y, acc[5]);
        acc[6] = fmaf(xf[6], w1.z, acc[6]); acc[7] = fmaf(xf[7], w1.w, acc[7]);
      }
    }
    const long long tok = (long long)b * p.R * p.R + (long long)row * p.R + col;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + tok * p.ldo + p.o_col0 + c0) = pack8(acc, 0);
  }
}

// Tiled variant: a block owns LEPE_RB image rows x 64 channels of one image; the RB+2 input rows are staged in shared
// memory once (each V element comes from L2 (RB+2)/RB times instead of 9 times), window borders are applied on the fly.
constexpr int LEPE_RB = 4;
__global__ void __launch_bounds__(256) lepe_tiled_kernel(const LepeParams p) {
  extern __shared__ uint4 lepe_smem[];                 // [(RB+2) rows][R tokens][8 x 16 B]
  const int R = p.R;
  const int cchunks = p.Cb / 64;
  const int cc = blockIdx.x % cchunks;
  const int rb = (blockIdx.x / cchunks) % ((R + LEPE_RB - 1) / LEPE_RB);
  const int b = blockIdx.x / (cchunks * ((R + LEPE_RB - 1) / LEPE_RB));
  const int row0 = rb * LEPE_RB;
  const int c0 = cc * 64;
  const int nrows = LEPE_RB + 2;
  for (int i = threadIdx.x; i < nrows * R * 8; i += 256) {
    const int cv = i & 7;
    const int col = (i >> 3) % R;
    const int rr = (i >> 3) / R;
    const int row = row0 + rr - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (row >= 0 && row < R) {
      const long long tok = (long long)b * R * R + (long long)row * R + col;
      val = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.v) + tok * p.ldv + p.v_col0 + c0 + cv * 8));
    }
    lepe_smem[i] = val;
  }
  __syncthreads();
  // a thread keeps its channel group for the whole loop (the stride 256 is a multiple of 8): the 9 x 8 tap weights and the bias
  // live in registers instead of being fetched per output (was 18 + 2 16-byte loads per output item, LSU-bound)
  const int cv = threadIdx.x & 7;
  const int ch = c0 + cv * 8;
  float w[9][8], bias8[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + t * p.Cb + ch));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + t * p.Cb + ch + 4));
    w[t][0] = w0.x; w[t][1] = w0.y; w[t][2] = w0.z; w[t][3] = w0.w;
    w[t][4] = w1.x; w[t][5] = w1.y; w[t][6] = w1.z; w[t][7] = w1.w;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) bias8[k] = __ldg(p.bias + ch + k);
  for (int i = threadIdx.x; i < LEPE_RB * R * 8; i += 256) {
    const int col = (i >> 3) % R;
    const int rr = (i >> 3) / R;
    const int row = row0 + rr;
    if (row >= R) continue;
    const int r_in = row % p.H_sp, c_in = col % p.W_sp;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = bias8[k];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      if (r_in + u - 1 < 0 || r_in + u - 1 >= p.H_sp) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        if (c_in + v - 1 < 0 || c_in + v - 1 >= p.W_sp) continue;
        float xf[8];
        unpack8(lepe_smem[((rr + u) * R + (col + v - 1)) * 8 + cv], 0, xf);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(xf[k], w[u * 3 + v][k], acc[k]);
      }
    }
    const long long tok = (long long)b * R * R + (long long)row * R + col;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + tok * p.ldo + p.o_col0 + ch) = pack8(acc, 0);
  }
}

}  // namespace pa
