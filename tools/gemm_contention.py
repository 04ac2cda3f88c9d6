"""Does a GEMM worker (CTA pair) run faster when fewer workers are active?  If the per-tile time falls as workers are
removed, the tile loop is limited by a shared resource (L2 / fabric bandwidth), not by the SM's tensor pipe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import _lib
from pytorch_attention_b200 import ops

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

torch.manual_seed(0)
for (M, N, K, bn) in [(12608, 2304, 768, 256), (12608, 768, 768, 192)]:
    A = torch.randn(M, K, device="cuda").half()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    D = torch.empty(M, N, device="cuda", dtype=torch.float16)
    tiles = ((M + 255) // 256) * ((N + bn - 1) // bn)
    print(f"M={M} N={N} K={K} bn={bn} pair tiles={tiles}; ideal MMA time per tile = {2*bn*(K//64)} cycles")
    for w in (74, 64, 56, 48, 37, 24, 12, 4, 1):
        os.environ["PA_GEMM_MAXWORKERS"] = str(w)
        _lib.reload_env()
        us = timed(lambda: ops.gemm_tn(A, B, out=D, block_n=bn, cluster=-2))
        waves = -(-tiles // w)
        print(f"  workers {w:3d}: {us:8.1f} us   waves {waves:4d}   us/wave {us/waves:6.2f}   L2->SM {tiles*(256+bn)*K*2/us/1e6:7.2f} TB/s")
    os.environ.pop("PA_GEMM_MAXWORKERS"); _lib.reload_env()
