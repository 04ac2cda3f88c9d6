"""Clock stamps of CTA 0 of attn_proj_kernel (attention core + projection in one kernel) on the PVT C3 geometry.
usage: attn_proj_trace.py [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
from pytorch_attention_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = pa.pvt.Attention(512, 8, sr_ratio=8).eval().half().cuda()
x = torch.randn(B, 4096, 512, device="cuda").half()
with torch.no_grad():
    for fused in ("1", "0"):
        os.environ["PA_PVT_FUSED"] = fused; _lib.reload_env()
        for _ in range(3):
            m(x, 64, 64)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            m(x, 64, 64)
        b.record(); torch.cuda.synchronize()
        print(f"PA_PVT_FUSED={fused}: {a.elapsed_time(b) / 10 * 1e3:.1f} us per forward (eager)")
    os.environ["PA_PVT_FUSED"] = "1"; _lib.reload_env()
    tr = torch.zeros(16384, dtype=torch.int64, device="cuda")
    lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    m(x, 64, 64)
    torch.cuda.synchronize()
    lib.pa_debug_set_gemm_trace(None)
t = tr.cpu()[:32 * 16].view(32, 16)
# the GEMMs of the same forward write stamps too (rows of 8): the fused kernel is the last launch, its rows win
print("unit: mma[start lastPV oall_ok chunk0..3 issued] | epi[first_s ro_done y0..y3 full]   (cycles from unit 0 start)")
t0 = int(t[0, 0])
for u in range(8):
    r = [int(v) - t0 if int(v) else 0 for v in t[u]]
    print(f" {u}: {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[3]:7d} {r[4]:7d} {r[5]:7d} {r[6]:7d} || {r[8]:7d} {r[9]:7d} | {r[10]:7d} {r[11]:7d} {r[12]:7d} {r[13]:7d}")
