"""Fused ViT kernel: time per forward for each (qkv, proj) tile-width pair, ViT-B and ViT-L shapes (CUDA-graph replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
os.environ["PA_VIT_COSCHED"] = "0"   # sweep the sequenced kernel
from pytorch_attention_b200 import _lib
import pytorch_attention_b200 as pa

def graph_time(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

torch.manual_seed(0)
for C, H in ((768, 12), (1024, 16)):
    m = pa.ViTAttention(C, H).eval().half().cuda()
    x = torch.randn(64, 197, C, device="cuda").half()
    with torch.no_grad():
        for bn1 in (256, 192):
            for bn2 in (256, 192):
                os.environ["PA_FUSED_BN1"], os.environ["PA_FUSED_BN2"] = str(bn1), str(bn2)
                _lib.reload_env()
                print(f"C={C}: qkv tiles 256x{bn1}, proj tiles 256x{bn2}: {graph_time(lambda: m(x)):7.2f} us")
        os.environ.pop("PA_FUSED_BN1"); os.environ.pop("PA_FUSED_BN2"); _lib.reload_env()
        print(f"C={C}: host cost model's choice: {graph_time(lambda: m(x)):7.2f} us")
