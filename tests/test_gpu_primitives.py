"""GPU parity of the two building blocks behind the C ABI (pa_gemm_tn, pa_attn_core) against fp32 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from pytorch_attention_b200 import ops
    return ops


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (300, 200, 192, 64), (394, 2304, 768, 0), (1000, 768, 768, 96),
                                      (257, 520, 320, 192), (129, 72, 64, 256)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_gemm_tn_vs_fp32(M, N, K, bn, dt):
    ops = _ops()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").to(dt)
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dt)
    bias = torch.randn(N, device="cuda")
    D = ops.gemm_tn(A, B, bias=bias, out_dtype=torch.float32, block_n=bn)
    ref = A.float() @ B.float().t() + bias
    # fp32 accumulation of exactly representable products: only summation-order differences remain
    assert (D - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


def test_gemm_pattern_bit_exact():
    """Integer-valued operands: every product and sum is exact, so the result must be bit-exact (index path of
    TMA swizzle <-> UMMA descriptors <-> TMEM epilogue)."""
    ops = _ops()
    M, N, K = 256, 192, 128
    A = torch.randint(-4, 5, (M, K), device="cuda").half()
    B = torch.randint(-4, 5, (N, K), device="cuda").half()
    D = ops.gemm_tn(A, B, out_dtype=torch.float32)
    assert torch.equal(D, A.float() @ B.float().t())


def test_gemm_batched_rowbias():
    ops = _ops()
    torch.manual_seed(1)
    Z, M, N, K = 3, 384, 196, 384
    W = (torch.randn(M, K, device="cuda") / K ** 0.5).half()
    O = torch.randn(Z, N, K, device="cuda").half()
    bias = torch.randn(M, device="cuda")
    D = ops.gemm_tn(W, O, bias=bias, bias_mode=2, out_dtype=torch.float32)
    ref = torch.einsum("mk,znk->zmn", W.float(), O.float()) + bias[None, :, None]
    assert (D - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def _attn_ref(q, k, v, scale):
    s = torch.einsum("gnhd,gmhd->ghnm", q.float(), k.float()) * scale
    return torch.einsum("ghnm,gmhd->gnhd", s.softmax(-1), v.float())


@pytest.mark.parametrize("G,N,H", [(1, 64, 1), (2, 197, 2), (2, 256, 3), (3, 50, 2), (1, 1, 1), (2, 129, 1)])
def test_attn_core_self(G, N, H):
    ops = _ops()
    torch.manual_seed(G * 1000 + N)
    C = H * 64
    qkv = torch.randn(G, N, 3 * C, device="cuda").half()
    o = ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C)
    q, k, v = qkv.reshape(G, N, 3, H, 64).unbind(2)
    ref = _attn_ref(q, k, v, 0.125).reshape(G, N, C)
    rel = ((o.float() - ref).norm() / ref.norm()).item()
    assert rel < 6e-4, rel      # fp16 P and fp16 O storage; fp32 accumulate


def test_attn_core_cross_many_query_tiles():
    ops = _ops()
    torch.manual_seed(5)
    G, Nq, Nk, H = 2, 1000, 64, 4
    C = H * 64
    q = torch.randn(G, Nq, C, device="cuda").half()
    kv = torch.randn(G, Nk, 2 * C, device="cuda").half()
    o = ops.attn_core(q, kv, H, 0.125, 0, 0, C)
    ref = _attn_ref(q.reshape(G, Nq, H, 64), kv[..., :C].reshape(G, Nk, H, 64), kv[..., C:].reshape(G, Nk, H, 64), 0.125)
    rel = ((o.float() - ref.reshape(G, Nq, C)).norm() / ref.norm()).item()
    assert rel < 6e-4, rel


def test_attn_uniform_when_keys_are_zero():
    """K = 0 makes every softmax row uniform: O = mean of V rows, exactly representable here."""
    ops = _ops()
    N, C = 64, 64
    qkv = torch.zeros(1, N, 3 * C, dtype=torch.float16, device="cuda")
    V = ((torch.arange(N)[:, None] % 8) + torch.arange(64)[None, :] / 64.0).half().cuda()
    qkv[0, :, 2 * C:] = V
    o = ops.attn_core(qkv, qkv, 1, 0.125, 0, C, 2 * C)
    assert torch.equal(o[0].float(), V.float().mean(0, keepdim=True).expand(N, -1).half().float())


@pytest.mark.parametrize("cluster", [1, 2, 4, -2])
@pytest.mark.parametrize("bn", [128, 256])
def test_gemm_cluster_modes_bit_exact(cluster, bn):
    """Every scheduling mode of the GEMM (single CTA, TMA-multicast clusters of 2 / 4, cta_group::2 pairs) must give
    the same bits on integer-valued operands (index path: multicast slices, peer-CTA halves of B, remote barriers)."""
    ops = _ops()
    torch.manual_seed(7)
    M, N, K = 1000, 520, 192
    A = torch.randint(-3, 4, (M, K), device="cuda").half()
    B = torch.randint(-3, 4, (N, K), device="cuda").half()
    D = ops.gemm_tn(A, B, out_dtype=torch.float32, block_n=bn, cluster=cluster)
    assert torch.equal(D, A.float() @ B.float().t())


@pytest.mark.parametrize("res_dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_gemm_residual_and_bias(res_dtype):
    ops = _ops()
    torch.manual_seed(8)
    M, N, K = 700, 512, 256
    A = torch.randn(M, K, device="cuda").half()
    B = (torch.randn(N, K, device="cuda") / 16).half()
    bias = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda").to(res_dtype)
    D = ops.gemm_tn(A, B, bias=bias, residual=R, out_dtype=torch.float32)
    ref = A.float() @ B.float().t() + bias + R.float()
    assert (D - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()


def test_gemm_unaligned_output_pitch():
    """Output rows of 196 fp16 (392 B, not 16-byte aligned: CvT's NCHW maps) take the smem-staged fallback store."""
    ops = _ops()
    torch.manual_seed(9)
    Z, M, N, K = 2, 192, 196, 128
    W = torch.randint(-3, 4, (M, K), device="cuda").half()
    O = torch.randint(-3, 4, (Z, N, K), device="cuda").half()
    D = ops.gemm_tn(W, O, out_dtype=torch.float16)
    assert D.stride(1) == 196
    assert torch.equal(D.float(), torch.einsum("mk,znk->zmn", W.float(), O.float()))


def test_gemm_balanced_walk_env(monkeypatch):
    """Opt-in balanced 64-column-unit tile walk (variable-width tiles, 32-row B boxes) is bit-exact too."""
    ops = _ops()
    from pytorch_attention_b200 import _lib
    monkeypatch.setenv("PA_GEMM_BALANCED", "1")
    _lib.reload_env()                      # the switches are cached by the library
    try:
        torch.manual_seed(10)
        M, N, K = 3000, 840, 320
        A = torch.randint(-3, 4, (M, K), device="cuda").half()
        B = torch.randint(-3, 4, (N, K), device="cuda").half()
        D = ops.gemm_tn(A, B, out_dtype=torch.float32)
        assert torch.equal(D, A.float() @ B.float().t())
    finally:
        monkeypatch.delenv("PA_GEMM_BALANCED")
        _lib.reload_env()
