"""Per-tile timeline of CTA 0 of the projection GEMM (clock64 stamps written by the kernel's debug hook)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
for (M, N, K, bn, cl) in [(12608, 2304, 768, 256, 1), (12608, 2304, 768, 256, -2), (12608, 768, 768, 192, 1)]:
    A = torch.randn(M, K, device="cuda").half()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    D = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.gemm_tn(A, B, out=D, block_n=bn, cluster=cl)
    tr = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    ops.gemm_tn(A, B, out=D, block_n=bn, cluster=cl)
    torch.cuda.synchronize()
    lib.pa_debug_set_gemm_trace(None)
    t = tr.cpu().view(64, 8)
    t0 = int(t[0, 0])
    print(f"--- M={M} N={N} K={K} bn={bn} cluster={cl}   (cycles since kernel start of CTA 0)")
    print("tile  prod_first  mma_start  first_full  mma_issued   epi_ready  epi_done   | mma_span  epi_span")
    for i in range(64):
        if int(t[i, 1]) == 0:
            break
        r = [int(t[i, s]) - t0 for s in (6, 1, 2, 3, 4, 5)]
        print(f"{i:4d}  {r[0]:10d} {r[1]:10d} {r[2]:11d} {r[3]:11d} {r[4]:11d} {r[5]:9d}   | {r[3]-r[1]:8d} {r[5]-r[4]:9d}")
    ck = t.reshape(-1)[320:352].view(8, 4)
    print("epilogue of tile 2, per 32-column sub-tile (warp 4 lane 0): loop_top  tmem_ld_done  staging_free(bar1)  staged(bar2)")
    for c in range(8):
        if int(ck[c, 0]) == 0: continue
        a, b, c2, d = [int(x) - t0 for x in ck[c]]
        print(f"   chunk {c}: {a:8d}  ld +{b-a:4d}  bar1 +{c2-b:4d}  store+fence+bar2 +{d-c2:4d}")
    kt = t[32:40].reshape(-1)[:64].view(16, 4)
    print("MMA thread, tile 2, per k-block: before_wait  after_wait  after_4_mma_issue  after_commit   (deltas vs previous k-block's after_commit)")
    prev = None
    for kb in range(16):
        if int(kt[kb, 0]) == 0:
            break
        a, b, c, d = [int(x) - t0 for x in kt[kb]]
        print(f"   kb {kb:2d}: {a:8d} {b:8d} {c:8d} {d:8d}   wait={b-a:5d} issue={c-b:5d} commit={d-c:4d}" + (f" gap_from_prev={a-prev:4d}" if prev else ""))
        prev = d
