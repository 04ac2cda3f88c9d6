"""One forward of each non-headline config (for an ncu launch list: which kernel costs what)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
torch.manual_seed(0)
dev = "cuda"
def go(mod, x, call):
    mod = mod.eval().half().cuda()
    with torch.no_grad():
        for _ in range(2):
            call(mod, x)
    torch.cuda.synchronize()
go(pa.pvt.Attention(512, 8, sr_ratio=8), torch.randn(32, 4096, 512, device=dev).half(), lambda m, x: m(x, 64, 64))
go(pa.cswin.CSWinBlock(512, 56, 16, split_size=7, qkv_bias=True), torch.randn(128, 3136, 512, device=dev).half(), lambda m, x: m.attention_half(x))
go(pa.xcit.XCA(768, 12), torch.randn(64, 196, 768, device=dev).half(), lambda m, x: m(x))
go(pa.cvt.Attention(384, 6), torch.randn(64, 384, 14, 14, device=dev).half(), lambda m, x: m(x))
go(pa.xcit.ClassAttention(768, 12), torch.randn(64, 197, 768, device=dev).half(), lambda m, x: m(x))
go(pa.cswin.CSWinBlock(512, 7, 16, split_size=7, qkv_bias=True, last_stage=True), torch.randn(128, 49, 512, device=dev).half(), lambda m, x: m.attention_half(x))
go(pa.vit.Attention(768, 4), torch.randn(64, 197, 768, device=dev).half(), lambda m, x: m(x))
# round-2 siblings and block halves
go(pa.segformer.Attention(512, 8, sr_ratio=8), torch.randn(32, 4096, 512, device=dev).half(), lambda m, x: m(x, 64, 64))
_rel = torch.randn(8, 4096, 64, device=dev)
go(pa.cmt.Attention(512, 8, sr_ratio=8), torch.randn(32, 4096, 512, device=dev).half(), lambda m, x: m(x, 64, 64, _rel))
go(pa.pvt.Block(512, 8, sr_ratio=8), torch.randn(32, 4096, 512, device=dev).half(), lambda m, x: m.attention_half(x, 64, 64))
go(pa.xcit.XCABlockAttentionHalf(768, 12, eta=1.0), torch.randn(64, 196, 768, device=dev).half(), lambda m, x: m(x))
go(pa.bvit.Broad_Attention(768, 12, 64), torch.randn(64, 197, 768, device=dev).half(), lambda m, x: m(x))
import os
from pytorch_attention_b200 import _lib
os.environ["PA_PVT_FUSED"] = "1"; _lib.reload_env()
go(pa.pvt.Attention(512, 8, sr_ratio=8), torch.randn(32, 4096, 512, device=dev).half(), lambda m, x: m(x, 64, 64))
