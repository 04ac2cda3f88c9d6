"""Experiment: fraction of pass-2 exponentials computed on the FMA pipe (PA_ATTN_DEBUG bits 8/16/32) -- time + error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import _lib
from pytorch_attention_b200 import ops

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

torch.manual_seed(0)
B, N, C, H = 64, 197, 768, 12
qkv = torch.randn(B, N, 3 * C, device="cuda").half()
out = torch.empty(B, N, C, device="cuda", dtype=torch.float16)
q, k, v = [t.view(B, N, H, 64).permute(0, 2, 1, 3).float() for t in qkv.split(C, dim=2)]
ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(B, N, C)
for dbg in (0,):
    os.environ["PA_ATTN_DEBUG"] = str(dbg)
    _lib.reload_env()
    fn = lambda: ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C, out=out)
    us = timed(fn)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"PA_ATTN_DEBUG={dbg:2d}: {us:7.2f} us   rel-Fro error vs fp32 softmax {err:.3e}")
