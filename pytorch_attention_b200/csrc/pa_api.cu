// pa_api.cu — C-ABI entry points of libpa_b200.so (see include/pa_b200.h).
#include "pa_attn.cuh"
#include "pa_gemm.cuh"
#include "pa_host.cuh"

#include <math.h>
#include <stdlib.h>

using namespace pa;

// ------------------------------------------------------------------------------------------------ GEMM
namespace {

template <int BN, int ST>
int launch_gemm_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, ST>;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    PA_CUDA_OK(cudaFuncSetAttribute(gemm_tn_kernel<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[dev & 63] = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.Z;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tn_kernel<BN, ST><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, p);
  PA_CUDA_OK(cudaGetLastError());
  launch_counter()++;
  return PA_OK;
}

int pick_block_n(int M, int N, int K, int Z) {
  const char* env = getenv("PA_GEMM_BN");
  if (env) {
    int v = atoi(env);
    if (v == 64 || v == 96 || v == 128 || v == 192 || v == 256) return v;
  }
  const int cands[5] = {256, 192, 128, 96, 64};
  const int sms = num_sms();
  const int m_tiles = (M + 127) / 128;
  const int num_kb = (K + 63) / 64;
  double best = 1e30;
  int best_bn = 128;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    const long long tiles = (long long)m_tiles * ((N + bn - 1) / bn) * Z;
    const long long waves = (tiles + sms - 1) / sms;
    // MMA cycles per tile + fixed per-tile overhead; narrow tiles pay extra L2 traffic per flop
    double per_tile = 2.0 * num_kb * bn + 700.0;
    if (bn < 128) per_tile *= 1.0 + 0.10 * (128.0 / bn - 1.0);
    const double cost = waves * per_tile;
    if (cost < best) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

int gemm_impl(const pa_gemm_args* a, cudaStream_t st) {
  if (!a) return fail(PA_ERR_NULL, "pa_gemm_tn: args is NULL");
  if (!a->A || !a->B || !a->D) return fail(PA_ERR_NULL, "pa_gemm_tn: A/B/D must be non-NULL");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0 || a->Z <= 0) return fail(PA_ERR_BAD_SHAPE, "pa_gemm_tn: M,N,K,Z must be positive");
  if (a->K % 8 != 0) return fail(PA_ERR_BAD_SHAPE, "pa_gemm_tn: K=%d must be a multiple of 8", a->K);
  if (a->a_dtype > 1 || a->b_dtype > 1 || a->a_dtype < 0 || a->b_dtype < 0) return fail(PA_ERR_UNSUPPORTED, "pa_gemm_tn: operands must be fp16/bf16");
  if (a->out_dtype < 0 || a->out_dtype > 2) return fail(PA_ERR_UNSUPPORTED, "pa_gemm_tn: bad out_dtype");
  if (a->bias_mode != 0 && !a->bias) return fail(PA_ERR_NULL, "pa_gemm_tn: bias_mode set but bias is NULL");
  int rc = current_device_check();
  if (rc) return rc;

  int bn = a->block_n ? a->block_n : pick_block_n(a->M, a->N, a->K, a->Z);
  CUtensorMap tmA, tmB;
  {
    const int za = a->a_batch ? a->Z : 1;
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->M, (uint64_t)za};
    uint64_t str[2] = {(uint64_t)a->lda * 2, (uint64_t)(a->a_batch ? a->a_batch : (long long)a->lda * a->M) * 2};
    uint32_t box[3] = {64, 128, 1};
    rc = make_tmap_16b(&tmA, a->a_dtype, a->A, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const int zb = a->b_batch ? a->Z : 1;
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->N, (uint64_t)zb};
    uint64_t str[2] = {(uint64_t)a->ldb * 2, (uint64_t)(a->b_batch ? a->b_batch : (long long)a->ldb * a->N) * 2};
    uint32_t box[3] = {64, (uint32_t)bn, 1};
    rc = make_tmap_16b(&tmB, a->b_dtype, a->B, 3, dims, str, box);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.Z = a->Z;
  p.m_tiles = (a->M + 127) / 128;
  p.n_tiles = (a->N + bn - 1) / bn;
  p.a_batched = a->a_batch != 0;
  p.b_batched = a->b_batch != 0;
  p.D = a->D; p.ldd = a->ldd; p.d_batch = a->d_batch;
  p.bias = a->bias; p.bias_mode = a->bias_mode; p.out_dtype = a->out_dtype;
  p.idesc = make_idesc(128, bn, a->a_dtype, a->b_dtype, 0, 0);
  switch (bn) {
    case 256: return launch_gemm_cfg<256, 4>(tmA, tmB, p, st);
    case 192: return launch_gemm_cfg<192, 5>(tmA, tmB, p, st);
    case 128: return launch_gemm_cfg<128, 6>(tmA, tmB, p, st);
    case 96:  return launch_gemm_cfg<96, 7>(tmA, tmB, p, st);
    case 64:  return launch_gemm_cfg<64, 8>(tmA, tmB, p, st);
    default: return fail(PA_ERR_UNSUPPORTED, "pa_gemm_tn: block_n %d not in {64,96,128,192,256}", bn);
  }
}

// ------------------------------------------------------------------------------------------------ attention core
int attn_impl(const pa_attn_args* a, cudaStream_t st) {
  if (!a) return fail(PA_ERR_NULL, "pa_attn_core: args is NULL");
  if (!a->q || !a->kv || !a->o) return fail(PA_ERR_NULL, "pa_attn_core: q/kv/o must be non-NULL");
  if (a->G <= 0 || a->H <= 0 || a->n_q <= 0 || a->n_k <= 0) return fail(PA_ERR_BAD_SHAPE, "pa_attn_core: G,H,n_q,n_k must be positive");
  if (a->n_k > 256) return fail(PA_ERR_UNSUPPORTED, "pa_attn_core: n_k=%d > 256 keys per unit not supported by this kernel", a->n_k);
  if (!(a->scale > 0.f)) return fail(PA_ERR_UNSUPPORTED, "pa_attn_core: scale must be > 0");
  if (a->ldo % 8 || a->o_col0 % 8 || a->o_group % 8 || (reinterpret_cast<uintptr_t>(a->o) & 15))
    return fail(PA_ERR_MISALIGNED, "pa_attn_core: output pitch/offset must be multiples of 8 elements");
  int rc = current_device_check();
  if (rc) return rc;

  AttnParams p;
  p.G = a->G; p.H = a->H; p.n_q = a->n_q; p.n_k = a->n_k;
  p.kp = (a->n_k + 15) / 16 * 16;
  p.q_tiles = (a->n_q + 127) / 128;
  p.pairs = (p.q_tiles + 1) / 2;
  p.items = a->G * a->H * p.pairs;
  p.q_col0 = a->q_col0; p.k_col0 = a->k_col0; p.v_col0 = a->v_col0;
  p.O = a->o; p.ldo = a->ldo; p.o_group = a->o_group; p.o_col0 = a->o_col0;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.idesc_s = make_idesc(128, p.kp, PA_F16, PA_F16, 0, 0);
  p.idesc_o = make_idesc(128, ATTN_HD, PA_F16, PA_F16, 0, 1);

  CUtensorMap tmQ, tmKV;
  {
    uint64_t dims[3] = {(uint64_t)a->ldq, (uint64_t)a->n_q, (uint64_t)a->G};
    uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->q_group * 2};
    uint32_t box[3] = {ATTN_HD, 128, 1};
    rc = make_tmap_16b(&tmQ, PA_DTYPE_F16, a->q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->ldkv, (uint64_t)a->n_k, (uint64_t)a->G};
    uint64_t str[2] = {(uint64_t)a->ldkv * 2, (uint64_t)a->kv_group * 2};
    uint32_t box[3] = {ATTN_HD, (uint32_t)p.kp, 1};
    rc = make_tmap_16b(&tmKV, PA_DTYPE_F16, a->kv, 3, dims, str, box);
    if (rc) return rc;
  }
  const int smem = attn_smem_bytes(p.kp);
  static int attr_smem[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_smem[dev & 63] < smem) {
    PA_CUDA_OK(cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(256)));
    attr_smem[dev & 63] = attn_smem_bytes(256);
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_core_kernel<<<grid, ATTN_THREADS, smem, st>>>(tmQ, tmKV, p);
  PA_CUDA_OK(cudaGetLastError());
  launch_counter()++;
  return PA_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

// ------------------------------------------------------------------------------------------------ exports
extern "C" {

int pa_version(void) { return PA_VERSION; }
const char* pa_last_error(void) { return err_buf(); }
unsigned long long pa_launch_count(void) { return launch_counter().load(); }

int pa_device_check(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(PA_ERR_DEVICE, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= n) return fail(PA_ERR_DEVICE, "device %d out of range (count %d)", device, n);
  int major = 0;
  PA_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(PA_ERR_DEVICE, "device %d is compute capability %d.x; sm_100 (B200) required", device, major);
  return PA_OK;
}

int pa_gemm_tn(const pa_gemm_args* a, void* stream) { return gemm_impl(a, (cudaStream_t)stream); }
int pa_attn_core(const pa_attn_args* a, void* stream) { return attn_impl(a, (cudaStream_t)stream); }

// ---------------------------------------------------------------- ViT
static int vit_check(const pa_vit_args* a) {
  if (!a) return fail(PA_ERR_NULL, "pa_vit: args is NULL");
  if (a->B <= 0 || a->N <= 0 || a->C <= 0 || a->H <= 0) return fail(PA_ERR_BAD_SHAPE, "pa_vit: B,N,C,H must be positive");
  if (a->C % a->H != 0) return fail(PA_ERR_BAD_SHAPE, "pa_vit: dim %d not divisible by num_heads %d (ViT.py:70)", a->C, a->H);
  if (a->C / a->H != 64) return fail(PA_ERR_UNSUPPORTED, "pa_vit: head_dim %d unsupported (64 only)", a->C / a->H);
  if (a->N > 256) return fail(PA_ERR_UNSUPPORTED, "pa_vit: N=%d tokens > 256 not supported yet", a->N);
  if (a->dtype != PA_DTYPE_F16 && a->dtype != PA_DTYPE_BF16) return fail(PA_ERR_UNSUPPORTED, "pa_vit: dtype must be fp16/bf16");
  return PA_OK;
}

size_t pa_vit_workspace_bytes(const pa_vit_args* a) {
  if (vit_check(a)) return 0;
  const size_t rows = (size_t)a->B * a->N;
  return align_up(rows * 3 * a->C * 2, 1024) + align_up(rows * a->C * 2, 1024) + 1024;
}

int pa_vit_fwd(const pa_vit_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = vit_check(a);
  if (rc) return rc;
  if (!a->x || !a->qkv_weight || !a->proj_weight || !a->y) return fail(PA_ERR_NULL, "pa_vit_fwd: x/qkv_weight/proj_weight/y must be non-NULL");
  const size_t need = pa_vit_workspace_bytes(a);
  if (!workspace || workspace_bytes < need) return fail(PA_ERR_WORKSPACE, "pa_vit_fwd: workspace %zu B < required %zu B", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)a->B * a->N;
  const int C = a->C;
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  void* qkv = ws;
  void* obuf = ws + align_up((size_t)rows * 3 * C * 2, 1024);

  // 1. qkv[B*N, 3C] = x Wqkv^T (+b)          (ViT.py:81)
  pa_gemm_args g1 = {};
  g1.a_dtype = a->dtype; g1.b_dtype = a->dtype; g1.out_dtype = PA_DTYPE_F16;
  g1.M = (int)rows; g1.N = 3 * C; g1.K = C; g1.Z = 1;
  g1.A = a->x; g1.lda = C; g1.B = a->qkv_weight; g1.ldb = C;
  g1.D = qkv; g1.ldd = 3 * C;
  g1.bias = a->qkv_bias; g1.bias_mode = a->qkv_bias ? 1 : 0;
  rc = gemm_impl(&g1, st);
  if (rc) return rc;
  // 2. per (b,h): softmax(q k^T scale) v       (ViT.py:83-86), O written as [B*N, C] with column h*64+d
  pa_attn_args at = {};
  at.G = a->B; at.H = a->H; at.n_q = a->N; at.n_k = a->N;
  at.q = qkv; at.ldq = 3 * C; at.q_group = (long long)a->N * 3 * C; at.q_col0 = 0;
  at.kv = qkv; at.ldkv = 3 * C; at.kv_group = (long long)a->N * 3 * C; at.k_col0 = C; at.v_col0 = 2 * C;
  at.o = obuf; at.ldo = C; at.o_group = (long long)a->N * C; at.o_col0 = 0;
  at.scale = a->scale;
  rc = attn_impl(&at, st);
  if (rc) return rc;
  // 3. y = O Wproj^T + b                       (ViT.py:87)
  pa_gemm_args g2 = {};
  g2.a_dtype = PA_DTYPE_F16; g2.b_dtype = PA_DTYPE_F16; g2.out_dtype = a->out_dtype;
  g2.M = (int)rows; g2.N = C; g2.K = C; g2.Z = 1;
  g2.A = obuf; g2.lda = C; g2.B = a->proj_weight; g2.ldb = C;
  g2.D = a->y; g2.ldd = C;
  g2.bias = a->proj_bias; g2.bias_mode = a->proj_bias ? 1 : 0;
  return gemm_impl(&g2, st);
}

}  // extern "C"
