"""Functional wrappers over the C ABI working on torch CUDA tensors (device memory + streams are torch's;
all arithmetic happens in libpa_b200.so).  No fallback path exists: CPU tensors raise."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_DT = {torch.float16: L.PA_DTYPE_F16, torch.bfloat16: L.PA_DTYPE_BF16, torch.float32: L.PA_DTYPE_F32}


def dtype_code(dt):
    try:
        return _DT[dt]
    except KeyError:
        raise ValueError(f"unsupported dtype {dt}; fp16 / bf16 (and fp32 outputs) only") from None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pytorch_attention_b200 runs on sm_100 CUDA devices only (no CPU fallback); "
                               "got a CPU tensor")


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_ws_cache = {}


def workspace(nbytes, device):
    """Cached per-device scratch buffer (grow-only).  Safe because every consumer is stream-ordered on the
    current stream; callers using several streams must pass their own workspace."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def cast_f32(x, dtype):
    """fp32 CUDA tensor -> contiguous fp16 / bf16 copy, by this library's cast kernel (pa_cast_f32)."""
    require_cuda(x)
    assert x.dtype == torch.float32 and dtype in (torch.float16, torch.bfloat16)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().pa_cast_f32(_ptr(x), _ptr(out), x.numel(), dtype_code(dtype), stream_ptr(x.device)))
    return out


def gemm_tn(a, b, bias=None, out=None, out_dtype=None, bias_mode=None, block_n=0, residual=None, cluster=0):
    """out[z][m,n] = sum_k a[z][m,k] b[z][n,k] (+bias).  a: [M,K] or [Z,M,K]; b: [N,K] or [Z,N,K] (row pitch may
    exceed K); bias fp32 per column (default) or per row (bias_mode=2)."""
    require_cuda(a, b, bias, out)
    za = a.shape[0] if a.dim() == 3 else 0
    zb = b.shape[0] if b.dim() == 3 else 0
    Z = max(za, zb, 1)
    M, K = a.shape[-2], a.shape[-1]
    N = b.shape[-2]
    assert b.shape[-1] == K, "K mismatch"
    assert a.stride(-1) == 1 and b.stride(-1) == 1, "operands must be K-major (last dim contiguous)"
    out_dtype = out_dtype or (out.dtype if out is not None else a.dtype)
    if out is None:
        shape = (Z, M, N) if (za or zb) else (M, N)
        out = torch.empty(shape, dtype=out_dtype, device=a.device)
    assert out.stride(-1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    g = L.GemmArgs()
    g.a_dtype, g.b_dtype, g.out_dtype = dtype_code(a.dtype), dtype_code(b.dtype), dtype_code(out.dtype)
    g.M, g.N, g.K, g.Z = M, N, K, Z
    g.A, g.lda, g.a_batch = _ptr(a), a.stride(-2), (a.stride(0) if za else 0)
    g.B, g.ldb, g.b_batch = _ptr(b), b.stride(-2), (b.stride(0) if zb else 0)
    g.D, g.ldd, g.d_batch = _ptr(out), out.stride(-2), (out.stride(0) if out.dim() == 3 else 0)
    g.bias = _ptr(bias)
    g.bias_mode = 0 if bias is None else (bias_mode or 1)
    g.block_n = block_n
    g.cluster = cluster
    if residual is not None:
        require_cuda(residual)
        assert residual.stride(-1) == 1 and residual.shape[-2:] == (M, N)
        g.residual, g.ldr = _ptr(residual), residual.stride(-2)
        g.r_batch = residual.stride(0) if residual.dim() == 3 else 0
        g.res_dtype = dtype_code(residual.dtype)
    with torch.cuda.device(a.device):
        L.check(L.load().pa_gemm_tn(C.byref(g), stream_ptr(a.device)))
    return out


def attn_core(q, kv, n_heads, scale, q_col0, k_col0, v_col0, out=None, o_col0=0, head_dim=64):
    """q: [G, n_q, ldq] fp16, kv: [G, n_k, ldkv] fp16 (heads are 64-wide column slices starting at *_col0).
    Returns out [G, n_q, n_heads*64] fp16 (or writes into `out` at column o_col0)."""
    require_cuda(q, kv, out)
    assert q.dtype == torch.float16 and kv.dtype == torch.float16
    assert q.stride(-1) == 1 and kv.stride(-1) == 1
    G, n_q = q.shape[0], q.shape[1]
    n_k = kv.shape[1]
    if out is None:
        out = torch.empty(G, n_q, n_heads * head_dim, dtype=torch.float16, device=q.device)
    a = L.AttnArgs()
    a.G, a.H, a.n_q, a.n_k = G, n_heads, n_q, n_k
    a.q, a.ldq, a.q_group, a.q_col0 = _ptr(q), q.stride(1), q.stride(0), q_col0
    a.kv, a.ldkv, a.kv_group, a.k_col0, a.v_col0 = _ptr(kv), kv.stride(1), kv.stride(0), k_col0, v_col0
    a.o, a.ldo, a.o_group, a.o_col0 = _ptr(out), out.stride(1), out.stride(0), o_col0
    a.scale = float(scale)
    a.head_dim = head_dim
    with torch.cuda.device(q.device):
        L.check(L.load().pa_attn_core(C.byref(a), stream_ptr(q.device)))
    return out


def run_with_workspace(x, args, ws_fn, fwd_fn):
    """Common tail of the module forwards: query the workspace size (0 = invalid arguments, let the entry point
    report the error), fetch the cached workspace, launch on the current stream."""
    import ctypes
    lib = L.load()
    with torch.cuda.device(x.device):
        need = getattr(lib, ws_fn)(ctypes.byref(args))
        if need == 0:
            L.check(getattr(lib, fwd_fn)(ctypes.byref(args), None, 0, stream_ptr(x.device)))
        ws = workspace(need, x.device)
        L.check(getattr(lib, fwd_fn)(ctypes.byref(args), _ptr(ws), ws.numel(), stream_ptr(x.device)))
