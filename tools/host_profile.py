"""Host-side cost of one drop-in forward (python + ctypes + C ABI + launch), by cProfile and by direct C-ABI timing."""
import os, sys, time, cProfile, pstats, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
from pytorch_attention_b200 import _lib as L, ops

mod = pa.ViTAttention(768, 12).eval().half().cuda()
x = torch.randn(64, 197, 768, device="cuda").half()
with torch.no_grad():
    for _ in range(5): mod(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): mod(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"module forward, host submit time: {(t1 - t0) / 300 * 1e6:.1f} us")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300): mod(x)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
    # the C ABI alone, arguments prepared once
    wq, bq, wp, bp = mod._staged(x.dtype)
    y = torch.empty_like(x)
    a = L.VitArgs()
    a.dtype, a.out_dtype = 0, 0
    a.B, a.N, a.C, a.H = 64, 197, 768, 12
    a.scale = 0.125
    a.x, a.qkv_weight, a.qkv_bias = ops._ptr(x), ops._ptr(wq), ops._ptr(bq)
    a.proj_weight, a.proj_bias, a.y = ops._ptr(wp), ops._ptr(bp), ops._ptr(y)
    lib = L.load()
    need = lib.pa_vit_workspace_bytes(ctypes.byref(a))
    ws = ops.workspace(need, x.device)
    st = ops.stream_ptr(x.device)
    wsp, n = ops._ptr(ws), ws.numel()
    for _ in range(5): lib.pa_vit_fwd(ctypes.byref(a), wsp, n, st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): lib.pa_vit_fwd(ctypes.byref(a), wsp, n, st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"pa_vit_fwd alone (ctypes call, fused launch), host time: {(t1 - t0) / 300 * 1e6:.1f} us")
    os.environ["PA_VIT_FUSED"] = "0"
    lib.pa_reload_env()
    for _ in range(5): lib.pa_vit_fwd(ctypes.byref(a), wsp, n, st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): lib.pa_vit_fwd(ctypes.byref(a), wsp, n, st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"pa_vit_fwd alone, three launches, host time: {(t1 - t0) / 300 * 1e6:.1f} us")
