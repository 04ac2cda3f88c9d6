"""A/B: attention core alone, two-slot CTA (attn_core_kernel) vs two single-slot CTAs per SM (attn_single_slot_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import _lib, ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_configs import graph_time
for name, (G, H, nq, nk, ld_q, ld_kv) in {"ViT-B": (64, 12, 197, 197, 2304, 2304), "PVT C3": (32, 8, 4096, 64, 512, 1024), "CvT 14x14": (64, 6, 196, 196, 1152, 1152)}.items():
    q = torch.randn(G, nq, ld_q, device="cuda").half()
    kv = q if ld_q == ld_kv and nq == nk else torch.randn(G, nk, ld_kv, device="cuda").half()
    C = H * 64
    kc, vc = (C, 2 * C) if kv is q else (0, C)
    out = {}
    for mode in ("two-slot", "single-slot x2"):
        if mode == "two-slot": os.environ["PA_ATTN_TWO_SLOT"] = "1"
        else: os.environ.pop("PA_ATTN_TWO_SLOT", None)
        _lib.reload_env()
        o = ops.attn_core(q, kv, H, 0.125, 0, kc, vc)
        us = graph_time(lambda: ops.attn_core(q, kv, H, 0.125, 0, kc, vc), 50)
        out[mode] = (us, o)
    same = torch.equal(out["two-slot"][1], out["single-slot x2"][1])
    print(f"{name}: two-slot {out['two-slot'][0]:.1f} us | single-slot x2 {out['single-slot x2'][0]:.1f} us | bit-equal {same}", flush=True)
