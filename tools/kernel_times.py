"""Per-kernel device time of one forward of each non-headline configuration, from torch.profiler (CUPTI activity records:
warm caches, back-to-back launches -- unlike an ncu launch list, which serialises and flushes).  usage: kernel_times.py [c3|c4|c4p|cvt|xca|cls|vitd ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import pytorch_attention_b200 as pa

dev = "cuda"
CFG = {
    "c3": lambda: (pa.pvt.Attention(512, 8, sr_ratio=8), torch.randn(32, 4096, 512, device=dev).half(), lambda m, x: m(x, 64, 64)),
    "c4": lambda: (pa.cswin.CSWinBlock(512, 56, 16, split_size=7, qkv_bias=True), torch.randn(128, 3136, 512, device=dev).half(), lambda m, x: m.attention_half(x)),
    "c4p": lambda: (pa.cswin.CSWinBlock(512, 7, 16, split_size=7, qkv_bias=True, last_stage=True), torch.randn(128, 49, 512, device=dev).half(), lambda m, x: m.attention_half(x)),
    "cvt": lambda: (pa.cvt.Attention(384, 6), torch.randn(64, 384, 14, 14, device=dev).half(), lambda m, x: m(x)),
    "xca": lambda: (pa.xcit.XCA(768, 12), torch.randn(64, 196, 768, device=dev).half(), lambda m, x: m(x)),
    "cls": lambda: (pa.xcit.ClassAttention(768, 12), torch.randn(64, 197, 768, device=dev).half(), lambda m, x: m(x)),
    "vitd": lambda: (pa.vit.Attention(768, 4), torch.randn(64, 197, 768, device=dev).half(), lambda m, x: m(x)),
}
for name in (sys.argv[1:] or list(CFG)):
    torch.manual_seed(0)
    mod, x, call = CFG[name]()
    mod = mod.eval().half().cuda()
    with torch.no_grad():
        for _ in range(3):
            call(mod, x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                call(mod, x)
            torch.cuda.synchronize()
    rows = [(e.key, e.self_device_time_total / e.count, e.count) for e in prof.key_averages() if e.self_device_time_total > 0]
    tot = sum(t * c for _, t, c in rows) / 5
    print(f"== {name}: {tot:.1f} us of kernel time per forward")
    for k, t, c in sorted(rows, key=lambda r: -r[1] * r[2]):
        print(f"   {t:9.1f} us x{c // 5:2d}  {k[:110]}")
