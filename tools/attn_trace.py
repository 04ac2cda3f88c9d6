"""Per-item timeline of CTA 0 of the attention core (clock64 stamps written by the kernel's debug hook)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
B, N, H = 64, 197, 12
C = H * 64
qkv = torch.randn(B, N, 3 * C, device="cuda").half()
o = torch.empty(B, N, C, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C, out=o)
tr = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attn_core(qkv, qkv, H, 0.125, 0, C, 2 * C, out=o)
torch.cuda.synchronize()
lib.pa_debug_set_gemm_trace(None)
t = tr.cpu().view(32, 16)
t0 = int(t[0, 0])
names = ["mma:q_ok", "S0_iss", "S1_iss", "PV0_iss", "PV1_iss", "w0:S_rdy", "w0:p1", "w0:p2", "w0:O_rdy", "w0:done",
         "w1:S_rdy", "w1:p1", "w1:p2", "w1:O_rdy", "w1:done"]
print("item " + " ".join(f"{n:>9s}" for n in names))
for i in range(32):
    if int(t[i, 1]) == 0:
        break
    print(f"{i:4d} " + " ".join(f"{int(t[i, s]) - t0:9d}" for s in range(15)))
