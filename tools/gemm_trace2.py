"""Stand-alone pair GEMM: per-tile timeline of CTA 0 (clock64) plus device time, for the ViT-B qkv / proj and the PVT q shapes."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
for (M, N, K, bias) in [(12608, 2304, 768, False), (12608, 768, 768, True), (131072, 512, 512, False)]:
    A = torch.randn(M, K, device="cuda").half()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bv = torch.randn(N, device="cuda") if bias else None
    D = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.gemm_tn(A, B, bias=bv, out=D)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.gemm_tn(A, B, bias=bv, out=D)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    tr = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    ops.gemm_tn(A, B, bias=bv, out=D)
    torch.cuda.synchronize()
    lib.pa_debug_set_gemm_trace(None)
    t = tr.cpu().view(64, 8)
    t0 = int(t[0, 0])
    print(f"--- M={M} N={N} K={K} bias={bias}: {us:.1f} us warm ({2.0 * M * N * K / us / 1e6:.0f} TFLOP/s)")
    print("tile  prod_first  mma_start  first_full  mma_issued   epi_ready  epi_done   | mma_span  epi_span")
    last = 0
    for i in range(64):
        if int(t[i, 1]) == 0:
            break
        r = [int(t[i, s]) - t0 for s in (6, 1, 2, 3, 4, 5)]
        last = r[5]
        print(f"{i:4d}  {r[0]:10d} {r[1]:10d} {r[2]:11d} {r[3]:11d} {r[4]:11d} {r[5]:9d}   | {r[3]-r[1]:8d} {r[5]-r[4]:9d}")
    print(f"   CTA 0 busy for {last} cycles; kernel {us:.1f} us -> {last / us / 1e3:.2f} GHz if CTA 0 spans the kernel")
