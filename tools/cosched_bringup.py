"""Bring-up of the co-scheduled single-launch ViT kernel (pa_cosched.cuh): bit-exactness against the three-launch path,
timing against the sequenced kernel, and the per-CTA role / timeline trace.  Own fp32 check, no oracle import."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
from pytorch_attention_b200 import _lib

lib = _lib.load()


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    _lib.reload_env()


def fresh(B, C, H, N, bias=False, seed=0):
    torch.manual_seed(seed)
    m = pa.ViTAttention(C, H, qkv_bias=bias).eval().half().cuda()
    x = torch.randn(B, N, C, device="cuda").half()
    return m, x


def timed(m, x, steps=200):
    xs = [torch.randn_like(x) for _ in range(8)]
    with torch.no_grad():
        for i in range(5):
            m(xs[i % 8])
        gs = []
        for i in range(8):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                m(xs[i])
            gs.append(g)
        for i in range(16):
            gs[i % 8].replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            gs[i % 8].replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3


def main():
    out = {}
    shapes = [(2, 768, 12, 197, False), (5, 128, 2, 197, True), (3, 256, 4, 64, False), (64, 768, 12, 197, False),
              (16, 1024, 16, 197, False), (7, 384, 6, 200, False), (1, 128, 2, 1, True)]
    ok_all = True
    for (B, C, H, N, bias) in shapes:
        m, x = fresh(B, C, H, N, bias)
        with torch.no_grad():
            setenv(PA_VIT_FUSED=0, PA_VIT_COSCHED=0)
            y3 = m(x)
            setenv(PA_VIT_FUSED=None, PA_VIT_COSCHED=1)
            n0 = _lib.launch_count()
            y1 = m(x)
            torch.cuda.synchronize()
            one = _lib.launch_count() - n0 == 1
            same = bool(torch.equal(y1, y3))
            rep = all(bool(torch.equal(m(x), y1)) for _ in range(5))
            # fp32 check of the same math
            w = m.qkv.weight.float(); wp = m.proj.weight.float()
            qkv = x.float() @ w.t() + (m.qkv.bias.float() if bias else 0)
            q, k, v = qkv.view(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
            att = torch.softmax((q @ k.transpose(-1, -2)) * m.scale, -1)
            ref = (att @ v).transpose(1, 2).reshape(B, N, C) @ wp.t() + m.proj.bias.float()
            err = ((y1.float() - ref).norm() / ref.norm()).item()
        print(f"shape B={B} C={C} H={H} N={N} bias={bias}: one_launch={one} bit_equal_3launch={same} repeat_equal={rep} rel_err={err:.2e}", flush=True)
        ok_all &= one and same and rep and err < 1e-3
    out["correct"] = ok_all

    m, x = fresh(64, 768, 12, 197)
    setenv(PA_VIT_FUSED=None, PA_VIT_COSCHED=1)
    t_cs = timed(m, x)
    setenv(PA_VIT_FUSED=1, PA_VIT_COSCHED=0)
    t_seq = timed(m, x)
    setenv(PA_VIT_FUSED=0, PA_VIT_COSCHED=0)
    t_3 = timed(m, x)
    print(f"ViT-B B=64: co-scheduled {t_cs:.1f} us | sequenced fused {t_seq:.1f} us | three launches {t_3:.1f} us", flush=True)
    out.update(vitb_cosched_us=t_cs, vitb_seq_us=t_seq, vitb_three_us=t_3)
    m2, x2 = fresh(64, 1024, 16, 197)
    setenv(PA_VIT_FUSED=None, PA_VIT_COSCHED=1)
    t_cs_l = timed(m2, x2)
    setenv(PA_VIT_FUSED=1, PA_VIT_COSCHED=0)
    t_seq_l = timed(m2, x2)
    print(f"ViT-L shard B=64: co-scheduled {t_cs_l:.1f} us | sequenced fused {t_seq_l:.1f} us", flush=True)
    out.update(vitl_cosched_us=t_cs_l, vitl_seq_us=t_seq_l)

    # ---- timeline
    setenv(PA_VIT_FUSED=None, PA_VIT_COSCHED=1)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        tr = torch.zeros(16384, dtype=torch.int64, device="cuda")
        lib.pa_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
        m(x)
        torch.cuda.synchronize()
        lib.pa_debug_set_gemm_trace(None)
    t = tr.cpu()[:296 * 8].view(296, 8)
    base = int(t[:, 2].min())
    for role in (0, 1):
        sel = t[t[:, 0] == role]
        st = (sel[:, 2] - base).double() / 1e3
        en = (sel[:, 4] - base).double() / 1e3
        line = f"role {'G' if role == 0 else 'A'}: {sel.shape[0]} CTAs, workers {sorted(set(sel[:, 1].tolist()))[:3]}..; start median {st.median():.1f} us; end min {en.min():.1f} median {en.median():.1f} max {en.max():.1f} us"
        if role == 1:
            fr = (sel[:, 3] - base).double() / 1e3
            line += f"; first unit ready median {fr.median():.1f} max {fr.max():.1f} us"
        print(line, flush=True)
    tg = tr.cpu()[4096:4096 + 64 * 8].view(64, 8)
    ta = tr.cpu()[8192:8192 + 64 * 16].view(64, 16)
    c0 = int(tg[0, 0])
    print("role G worker 0 (cycles since its first load): prod_first | mma_start first_full mma_issued | epi_ready epi_tmem_done epi_stored published")
    for i in range(64):
        if int(tg[i, 1]) == 0:
            break
        print("  tile %2d: %7d | %7d %7d %7d | %7d %7d %7d %7d" % ((i,) + tuple(int(v) - c0 for v in tg[i, :8])))
    a0 = int(ta[0, 0])
    print("role A worker 0 (cycles since its first unit's deps): deps_ok | S_issued s_full pass1 pass2 | PV_issued o_full staged published")
    for i in range(64):
        if int(ta[i, 1]) == 0:
            break
        r = [int(v) - a0 for v in ta[i, :9]]
        print("  unit %2d: %7d | %7d %7d %7d %7d | %7d %7d %7d %7d" % (i, r[0], r[1], r[2], r[3], r[4], r[7], r[5], r[6], r[8]))
    per_sm = {}
    for r in t.tolist():
        per_sm.setdefault(r[5], []).append(r[0])
    mixed = sum(1 for v in per_sm.values() if sorted(v) == [0, 1])
    print(f"SMs hosting exactly one G and one A CTA: {mixed} of {len(per_sm)}", flush=True)
    out["sms_mixed"] = mixed
    setenv(PA_VIT_FUSED=None, PA_VIT_COSCHED=None)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/cosched_bringup.json", "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
