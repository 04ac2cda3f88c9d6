"""Timeline of the co-scheduled ViT kernel (worker 0 of each role) under the PA_CS_DEBUG experiments:
0 = normal, 1 = role A idle (GEMM stream alone), 2 = role G idle (attention stream alone on stale qkv)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
from pytorch_attention_b200 import _lib
lib = _lib.load()
B, C, H, N = [int(v) for v in (sys.argv[2:6] if len(sys.argv) >= 6 else (64, 768, 12, 197))]
torch.manual_seed(0)
m = pa.ViTAttention(C, H).eval().half().cuda()
x = torch.randn(B, N, C, device="cuda").half()
for dbg in [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2")]:
    os.environ["PA_VIT_COSCHED"] = "1"; os.environ["PA_CS_DEBUG"] = str(dbg); _lib.reload_env()
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            m(x)
        b.record(); torch.cuda.synchronize()
        tr = torch.zeros(16384, dtype=torch.int64, device="cuda")
        lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
        m(x)
        torch.cuda.synchronize()
        lib.pa_debug_set_gemm_trace(None)
    print(f"==== PA_CS_DEBUG={dbg}: {a.elapsed_time(b) / 20 * 1e3:.1f} us per forward (eager, L2-warm)")
    t = tr.cpu()
    ct = t[:296 * 8].view(296, 8)
    base = int(ct[:, 2].min())
    for role in (0, 1):
        sel = ct[ct[:, 0] == role]
        en = (sel[:, 4] - base).double() / 1e3
        print(f"  role {'GA'[role]}: end min {en.min():.1f} median {en.median():.1f} max {en.max():.1f} us")
    tg = t[4096:4096 + 64 * 8].view(64, 8); ta = t[8192:8192 + 64 * 16].view(64, 16)
    if int(tg[0, 1]):
        c0 = int(tg[0, 0])
        print("  G worker 0: prod_first | mma_start first_full mma_issued | epi_ready epi_tmem_done epi_stored published   [mainloop, drain]")
        for i in range(64):
            if int(tg[i, 1]) == 0: break
            r = [int(v) - c0 for v in tg[i, :8]]
            print("   tile %2d: %7d | %7d %7d %7d | %7d %7d %7d %7d   [%5d, %5d]" % (i, *r, r[3] - r[2], r[5] - r[4]))
    tc = t[12288:12288 + 32].view(4, 8)
    if int(tc[0, 0]):
        print("  G worker 0 tile 2, warp 4, per 32-column chunk: top->ld_done  cvt  wait_read  sts+fence  tma_issue | next top")
        for j in range(4):
            r = [int(v) for v in tc[j, :6]]
            nxt = int(tc[j + 1, 0]) - r[5] if j < 3 else 0
            print("   chunk %d: %5d %5d %5d %5d %5d | %5d" % (j, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], nxt))
    if int(ta[0, 1]):
        a0 = int(ta[0, 0])
        print("  A worker 0: deps_ok | S_issued s_full pass1 pass2 | PV_issued o_full staged published   [S, pass1, pass2, PV, out]")
        for i in range(64):
            if int(ta[i, 1]) == 0: break
            r = [int(v) - a0 for v in ta[i, :9]]
            print("   unit %2d: %7d | %7d %7d %7d %7d | %7d %7d %7d %7d   [%4d %5d %5d %5d %5d]" % (i, r[0], r[1], r[2], r[3], r[4], r[7], r[5], r[6], r[8], r[2]-r[1], r[3]-r[2], r[4]-r[3], r[5]-r[4], r[6]-r[5]))
os.environ.pop("PA_CS_DEBUG"); os.environ.pop("PA_VIT_COSCHED"); _lib.reload_env()
