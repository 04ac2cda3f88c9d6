"""Phase timeline of CTA 0 of the fused single-launch ViT kernel (clock64 stamps per tile/item, globaltimer per phase)."""
import sys, os, ctypes
os.environ["PA_VIT_FUSED"] = "1"
os.environ["PA_VIT_COSCHED"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
from pytorch_attention_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
m = pa.ViTAttention(768, 12).eval().half().cuda()
x = torch.randn(64, 197, 768, device="cuda").half()
with torch.no_grad():
    for _ in range(3):
        m(x)
    tr = torch.zeros(4096, dtype=torch.int64, device="cuda")
    lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    m(x)
    torch.cuda.synchronize()
    lib.pa_debug_set_gemm_trace(None)
t = tr.cpu()
g1 = t[:512].view(64, 8); at = t[512:1024].view(32, 16); g2 = t[1024:1536].view(64, 8)
t0 = int(g1[0, 0])
def rel(v): return int(v) - t0 if int(v) else None
print("phase 1 (qkv GEMM) tiles: mma_start / epi_done")
for i in range(64):
    if int(g1[i, 1]) == 0: break
    print(f"  tile {i}: mma_start {rel(g1[i,1])}  first_full {rel(g1[i,2])}  mma_issued {rel(g1[i,3])}  epi_ready {rel(g1[i,4])}  epi_done {rel(g1[i,5])}")
ck = g1.reshape(-1)[320:352].view(8, 4)
print("phase 1 epilogue of tile 2, per 32-column sub-tile (warp 4 lane 0): loop_top  tmem_ld_done  staging_free(bar1)  staged(bar2)")
for c in range(8):
    if int(ck[c, 0]) == 0: continue
    a, b, c2, d = [int(x) - t0 for x in ck[c]]
    print(f"   chunk {c}: {a:8d}  ld +{b-a:4d}  bar1 +{c2-b:4d}  store+fence+bar2 +{d-c2:4d}")
print("phase 2 (attention) items: q_ok / S0_iss / w0 done / w1 done")
for i in range(32):
    if int(at[i, 1]) == 0: break
    print(f"  item {i}: q_ok {rel(at[i,0])}  S0 {rel(at[i,1])}  w0:S_rdy {rel(at[i,5])} w0:done {rel(at[i,9])}  w1:done {rel(at[i,14])}")
print("phase 3 (proj GEMM) tiles")
t3 = int(g2[0, 0])
for i in range(64):
    if int(g2[i, 1]) == 0: break
    print(f"  tile {i}: prod_first {rel(g2[i,6])} mma_start {rel(g2[i,1])}  first_full {rel(g2[i,2])}  mma_issued {rel(g2[i,3])}  epi_done {rel(g2[i,5])}")
ph = t[1536:1536 + 148 * 4].view(148, 4).double()
base = ph[:, 0].min()
ph = (ph - base) / 1000.0
for k, name in enumerate(["start", "end phase 1 (qkv)", "end phase 2 (attention)", "end phase 3 (proj)"]):
    col = ph[:, k]
    print(f"{name:26s} min {col.min():7.1f} us  median {col.median():7.1f}  max {col.max():7.1f}  (argmax CTA {int(col.argmax())})")
