"""Drop-in for ``vision_transformers/segformer.py:Attention`` (segformer.py:17-50): PVT's spatial-reduction attention with a
DENSE reduction conv and a fused kv Linear."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


class Attention(StagedModule):
    """Same constructor / ``forward(x[B,N,C], H, W)`` / ``state_dict`` keys as the reference (segformer.py:18-31): ``q``, ``kv``
    (fused, rows ordered (k|v, head, d) -- the [k|v] layout the PVT path uses anyway), ``proj`` Linear and, for sr_ratio > 1,
    ``sr = Conv2d(dim, dim, kernel_size=sr, stride=sr)`` -- dense, with bias, no norm behind it (segformer.py:27).
    Kernel == stride makes that conv a GEMM over non-overlapping patches: ``sr_patchify_kernel`` re-partitions x into the
    K-major patch matrix ``[B*M, sr*sr*C]`` (every element moves once) and the conv weight is staged as ``[C, (u, v, ci)]``.
    Launch sequence: [patchify -> GEMM(sr)] -> GEMM(q) -> GEMM(kv) -> attention core -> GEMM(proj)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0, proj_drop=0, sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, 2 * dim, bias=qkv_bias)
        self.sr_ratio = sr_ratio
        if self.sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.out_dtype = None
        self._init_stage()

    def _staged(self, dtype):
        srcs = [self.q.weight, self.q.bias, self.kv.weight, self.kv.bias, self.proj.weight, self.proj.bias]
        if self.sr_ratio > 1:
            srcs += [self.sr.weight, self.sr.bias]

        def build():
            kv_dtype = torch.float16 if self.sr_ratio > 1 else dtype
            d = dict(wq=w16(self.q.weight, dtype), bq=f32(self.q.bias), wkv=w16(self.kv.weight, kv_dtype), bkv=f32(self.kv.bias),
                     wp=w16(self.proj.weight, torch.float16), bp=f32(self.proj.bias), srw=None, srb=None)
            if self.sr_ratio > 1:
                C = self.sr.weight.shape[0]
                # [co, ci, u, v] -> [co, (u, v, ci)]: the K order of the patch matrix
                d["srw"] = self.sr.weight.detach().permute(0, 2, 3, 1).reshape(C, -1).to(dtype).contiguous()
                d["srb"] = f32(self.sr.bias)
            return d
        return self._stage.get(("w", dtype), srcs, build)

    def _check(self, x):
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))

    def _fill(self, a, x, y, H, W, s):
        B, N, C = x.shape
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.N, a.C, a.H = B, N, C, self.num_heads
        a.Himg, a.Wimg, a.sr = int(H), int(W), self.sr_ratio
        a.scale = float(self.scale)
        a.x, a.y = ops._ptr(x), ops._ptr(y)
        a.q_weight, a.q_bias = ops._ptr(s["wq"]), ops._ptr(s["bq"])
        a.kv_weight, a.kv_bias = ops._ptr(s["wkv"]), ops._ptr(s["bkv"])
        a.proj_weight, a.proj_bias = ops._ptr(s["wp"]), ops._ptr(s["bp"])
        a.sr_mode = 1
        a.sr_dense_weight, a.sr_dense_bias = ops._ptr(s["srw"]), ops._ptr(s["srb"])

    def forward(self, x, H, W):
        x, y_dtype = self._prepare_input(x)
        self._check(x)
        B, N, C = x.shape
        x = x.contiguous()
        s = self._staged(x.dtype)
        y = torch.empty(B, N, C, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.PvtArgs()
        self._fill(a, x, y, H, W, s)
        ops.run_with_workspace(x, a, "pa_pvt_workspace_bytes", "pa_pvt_fwd")
        return y
