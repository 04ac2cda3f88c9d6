#!/usr/bin/env python
"""GPU bring-up harness: runs each probe in its own subprocess (a trapped kernel kills only that probe) and
appends JSON lines to gpurun_out/bringup.jsonl.   python tools/gpu_bringup.py [probe ...]"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = os.path.join(OUT, "bringup.jsonl")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(LOG, "a") as f:
        f.write(line + "\n")


def time_cuda(fn, iters=20, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


# ------------------------------------------------------------------ probes
def probe_gemm_pattern():
    """One-hot A reveals which B element each (m,n) accumulates: checks TMA swizzle <-> UMMA descriptor agreement."""
    import torch
    from pytorch_attention_b200 import ops
    for bn in (64, 128, 256):
        M, N, K = 128, bn, 64
        A = torch.zeros(M, K, dtype=torch.float16, device="cuda")
        A[torch.arange(M), torch.arange(M) % K] = 1
        Bm = ((torch.arange(N)[:, None] % 32) * 64 + torch.arange(K)[None, :]).to(torch.float16).cuda()
        D = ops.gemm_tn(A, Bm, out_dtype=torch.float32, block_n=bn)
        ref = A.float() @ Bm.float().t()
        torch.cuda.synchronize()
        bad = (D != ref).sum().item()
        emit(probe="gemm_pattern", bn=bn, mismatches=bad, sample_got=D[:3, :6].tolist(), sample_ref=ref[:3, :6].tolist())


def probe_gemm():
    import torch
    from pytorch_attention_b200 import ops
    torch.manual_seed(0)
    cases = [
        # M, N, K, bn, dtype, bias, out
        (128, 128, 64, 128, torch.float16, False, torch.float32),
        (128, 256, 768, 256, torch.float16, False, torch.float32),
        (300, 200, 192, 64, torch.float16, True, torch.float16),
        (394, 2304, 768, 0, torch.float16, False, torch.float16),
        (394, 2304, 768, 0, torch.bfloat16, True, torch.float16),
        (12608, 2304, 768, 256, torch.float16, False, torch.float16),
        (12608, 2304, 768, 192, torch.float16, False, torch.float16),
        (12608, 2304, 768, 128, torch.float16, False, torch.float16),
        (12608, 768, 768, 0, torch.float16, True, torch.float16),
        (12608, 768, 768, 64, torch.float16, True, torch.float32),
        (12608, 768, 768, 96, torch.float16, True, torch.bfloat16),
        (131072, 512, 512, 0, torch.float16, True, torch.float16),
    ]
    for (M, N, K, bn, dt, has_bias, odt) in cases:
        A = torch.randn(M, K, device="cuda").to(dt)
        Bm = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dt)
        bias = torch.randn(N, device="cuda") if has_bias else None
        D = ops.gemm_tn(A, Bm, bias=bias, out_dtype=odt, block_n=bn)
        ref = A.float() @ Bm.float().t()
        if has_bias:
            ref = ref + bias
        torch.cuda.synchronize()
        err = (D.float() - ref).abs().max().item()
        rel = ((D.float() - ref).norm() / ref.norm()).item()
        us = time_cuda(lambda: ops.gemm_tn(A, Bm, bias=bias, out=D, block_n=bn))
        emit(probe="gemm", M=M, N=N, K=K, bn=bn, dtype=str(dt), out=str(odt), bias=has_bias, max_abs=err, rel_fro=rel,
             us=us, tflops=2.0 * M * N * K / us / 1e6)


def probe_gemm_sweep():
    """block_n x cluster sweep on the projection shapes of the BASELINE configs (correctness + time)."""
    import torch
    from pytorch_attention_b200 import ops
    torch.manual_seed(0)
    shapes = [(12608, 2304, 768), (12608, 768, 768), (12608, 3072, 1024), (131072, 512, 512), (401408, 1536, 512)]
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device="cuda").half()
        Bm = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        ref = (A[:512].float() @ Bm.float().t())
        for bn in (128, 192, 256):
            for cl in (1, -2):
                D = torch.empty(M, N, device="cuda", dtype=torch.float16)
                try:
                    ops.gemm_tn(A, Bm, out=D, block_n=bn, cluster=cl)
                    torch.cuda.synchronize()
                    err = (D[:512].float() - ref).abs().max().item()
                    err2 = (D[-300:].float() - A[-300:].float() @ Bm.float().t()).abs().max().item()
                    us = time_cuda(lambda: ops.gemm_tn(A, Bm, out=D, block_n=bn, cluster=cl), iters=10)
                    emit(probe="gemm_sweep", M=M, N=N, K=K, bn=bn, cluster=cl, max_abs=max(err, err2), us=round(us, 2),
                         tflops=round(2.0 * M * N * K / us / 1e6, 1))
                except Exception as e:  # noqa: BLE001
                    emit(probe="gemm_sweep", M=M, N=N, K=K, bn=bn, cluster=cl, error=str(e)[:200])
        del A, Bm


def probe_gemm_batched():
    import torch
    from pytorch_attention_b200 import ops
    torch.manual_seed(1)
    Z, M, N, K = 3, 384, 196, 384     # CvT-style: y[b] = W[C,C] . O[b][HW,C]^T, bias per row
    W = (torch.randn(M, K, device="cuda") / K ** 0.5).half()
    O = torch.randn(Z, N, K, device="cuda").half()
    bias = torch.randn(M, device="cuda")
    D = ops.gemm_tn(W, O, bias=bias, bias_mode=2, out_dtype=torch.float16)
    ref = torch.einsum("mk,znk->zmn", W.float(), O.float()) + bias[None, :, None]
    torch.cuda.synchronize()
    emit(probe="gemm_batched", max_abs=(D.float() - ref).abs().max().item(), rel_fro=((D.float() - ref).norm() / ref.norm()).item())


def _attn_ref(qkv, B, N, H, scale):
    import torch
    C = H * 64
    q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B, N, C)


def probe_attn_uniform():
    """K = 0 -> P uniform -> O = mean_j V[j]: isolates the P(TMEM) x V(MN-major smem) MMA from the QK^T MMA."""
    import torch
    from pytorch_attention_b200 import ops
    B, N, H = 1, 64, 1
    C = 64
    qkv = torch.zeros(B, N, 3 * C, dtype=torch.float16, device="cuda")
    V = ((torch.arange(N)[:, None] % 8) + torch.arange(64)[None, :] / 64.0).half()
    qkv[0, :, 2 * C:] = V.cuda()
    o = ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C)
    torch.cuda.synchronize()
    ref = V.float().mean(0)
    emit(probe="attn_uniform", max_abs=(o[0].float().cpu() - ref[None]).abs().max().item(), got=o[0, :2, :8].tolist(), ref=ref[:8].tolist())
    # one-hot P: q.k large for j == i  -> O[i] ~= V[i]
    qkv = torch.zeros(B, N, 3 * C, dtype=torch.float16, device="cuda")
    eye = torch.eye(64, dtype=torch.float16, device="cuda") * 16
    qkv[0, :, :C] = eye
    qkv[0, :, C:2 * C] = eye
    Vr = torch.randn(N, 64).half()
    qkv[0, :, 2 * C:] = Vr.cuda()
    o = ops.attn_core(qkv, qkv, H, 1.0, 0, C, 2 * C)
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, B, N, H, 1.0)
    emit(probe="attn_onehot", max_abs=(o.float() - ref).abs().max().item(), got=o[0, 1, :6].tolist(), ref=ref[0, 1, :6].tolist())


def probe_attn():
    import torch
    from pytorch_attention_b200 import ops
    torch.manual_seed(2)
    for (B, N, H) in [(1, 64, 1), (1, 128, 1), (2, 197, 2), (2, 256, 3), (3, 50, 2), (64, 197, 12)]:
        C = H * 64
        qkv = torch.randn(B, N, 3 * C, device="cuda").half()
        o = ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C)
        ref = _attn_ref(qkv, B, N, H, 0.125)
        torch.cuda.synchronize()
        err = (o.float() - ref).abs().max().item()
        rel = ((o.float() - ref).norm() / ref.norm()).item()
        us = time_cuda(lambda: ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C, out=o))
        emit(probe="attn", B=B, N=N, H=H, max_abs=err, rel_fro=rel, us=us)
    # PVT-like: many query tiles, 64 keys
    B, Nq, Nk, H = 2, 4096, 64, 8
    C = H * 64
    q = torch.randn(B, Nq, C, device="cuda").half()
    kv = torch.randn(B, Nk, 2 * C, device="cuda").half()
    o = ops.attn_core(q, kv, H, 0.125, 0, 0, C)
    qq = q.float().reshape(B, Nq, H, 64).permute(0, 2, 1, 3)
    kk = kv[..., :C].float().reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    vv = kv[..., C:].float().reshape(B, Nk, H, 64).permute(0, 2, 1, 3)
    ref = (((qq @ kk.transpose(-1, -2)) * 0.125).softmax(-1) @ vv).transpose(1, 2).reshape(B, Nq, C)
    torch.cuda.synchronize()
    emit(probe="attn_pvt_like", max_abs=(o.float() - ref).abs().max().item(), rel_fro=((o.float() - ref).norm() / ref.norm()).item())


def probe_vit():
    import torch
    import pytorch_attention_b200 as pa
    def vit_attention(x, wq, bq, wp, bp, H):        # plain fp32 check inside the tool (oracle/ is for tests, smoke and bench only)
        B, N, C = x.shape
        qkv = torch.nn.functional.linear(x, wq, bq).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        a = ((qkv[0] @ qkv[1].transpose(-1, -2)) * (C // H) ** -0.5).softmax(-1)
        return torch.nn.functional.linear((a @ qkv[2]).transpose(1, 2).reshape(B, N, C), wp, bp)
    torch.manual_seed(3)
    for (B, N, C, H, dt, odt, iters) in [(2, 197, 768, 12, torch.float16, torch.float16, 20),
                                         (2, 197, 768, 12, torch.float16, torch.float32, 20),
                                         (2, 197, 768, 12, torch.bfloat16, torch.bfloat16, 20),
                                         (64, 197, 768, 12, torch.float16, torch.float16, 50),
                                         (64, 197, 768, 12, torch.bfloat16, torch.float16, 50),
                                         (64, 197, 1024, 16, torch.float16, torch.float16, 50)]:
        m = pa.ViTAttention(C, H).eval()
        with torch.no_grad():
            for p_ in m.parameters():
                p_.copy_(p_.to(dt).float())
        x = torch.randn(B, N, C).to(dt)
        sd = {k: v.float() for k, v in m.state_dict().items()}
        if B <= 8:
            ref = vit_attention(x.float(), sd["qkv.weight"], sd.get("qkv.bias"), sd["proj.weight"], sd["proj.bias"], H)
        else:
            ref = None
        m = m.cuda()
        m.out_dtype = odt
        xg = x.cuda()
        with torch.no_grad():
            y = m(xg)
            torch.cuda.synchronize()
            if ref is None:   # torch GPU fp32 as a stand-in at the large shape
                mm = torch.nn.functional
                qkv = mm.linear(xg.float(), m.qkv.weight.float()).reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
                a = ((qkv[0] @ qkv[1].transpose(-1, -2)) * m.scale).softmax(-1)
                ref = mm.linear((a @ qkv[2]).transpose(1, 2).reshape(B, N, C), m.proj.weight.float(), m.proj.bias.float()).cpu()
            yc = y.float().cpu()
            rel = ((yc - ref).norm() / ref.norm()).item()
            relmax = ((yc - ref).abs().max() / ref.abs().max()).item()
            us = time_cuda(lambda: m(xg), iters=iters)
        flops = B * (8 * N * C * C + 4 * N * N * C)
        emit(probe="vit", B=B, N=N, C=C, H=H, dtype=str(dt), out=str(odt), rel_fro=rel, rel_max=relmax, us=us,
             tokens_per_s=B * N / us * 1e6, tflops=flops / us / 1e6)


PROBES = {
    "gemm_pattern": probe_gemm_pattern,
    "gemm": probe_gemm,
    "gemm_sweep": probe_gemm_sweep,
    "gemm_batched": probe_gemm_batched,
    "attn_uniform": probe_attn_uniform,
    "attn": probe_attn,
    "vit": probe_vit,
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        PROBES[sys.argv[2]]()
        return
    names = sys.argv[1:] or list(PROBES)
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=240, capture_output=True, text=True)
            tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:]).strip()
            emit(probe=n, status="exit", rc=r.returncode, secs=round(time.time() - t0, 1), tail=tail if r.returncode else "")
            if r.returncode == 0:
                sys.stdout.write(r.stdout)
        except subprocess.TimeoutExpired as e:
            emit(probe=n, status="timeout", secs=round(time.time() - t0, 1), tail=str(e.stdout)[-2000:] if e.stdout else "")


if __name__ == "__main__":
    main()
