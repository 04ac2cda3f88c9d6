"""Drop-in for ``vision_transformers/p2t.py:PoolingAttention`` (p2t.py:46-94)."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


class PoolingAttention(StagedModule):
    """Same constructor / ``state_dict`` keys (``q.0``, ``kv.0``, ``proj``, ``norm``) / ``forward(x[B,N,C], H, W, d_convs)`` as the
    reference (p2t.py:47-72, 74-94).  Queries come from x; keys / values from a pooling pyramid of the token map: per ratio
    ``adaptive_avg_pool2d`` to ``(round(H / r), round(W / r))``, ``pool + d_convs[i](pool)`` (the model's depthwise 3x3 convs,
    passed in exactly like the reference's ``d_convs``), concatenation, LayerNorm, ``kv`` Linear.  One C-ABI call
    (``pa_p2t_fwd``): pooling kernel -> depthwise conv + skip + LayerNorm kernel -> GEMM(q) -> GEMM(kv) -> attention core ->
    GEMM(proj)."""

    def __init__(self, dim, num_heads=2, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., pool_ratios=[1, 2, 3, 6]):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        self.num_elements = np.array([t * t for t in pool_ratios]).sum()
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.q = nn.Sequential(nn.Linear(dim, dim, bias=qkv_bias))
        self.kv = nn.Sequential(nn.Linear(dim, dim * 2, bias=qkv_bias))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.pool_ratios = pool_ratios
        self.pools = nn.ModuleList()
        self.norm = nn.LayerNorm(dim)
        self.out_dtype = None
        self._init_stage()

    def forward(self, x, H, W, d_convs=None):
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))
        if d_convs is None or len(d_convs) != len(self.pool_ratios):
            raise ValueError("d_convs must hold one depthwise 3x3 conv per pool ratio (p2t.py:79)")
        if len(self.pool_ratios) > 4:
            raise ValueError("at most four pyramid levels are supported")
        B, N, C = x.shape
        x = x.contiguous()
        q, kv, p, n = self.q[0], self.kv[0], self.proj, self.norm
        srcs = [q.weight, q.bias, kv.weight, kv.bias, p.weight, p.bias, n.weight, n.bias]
        for l in d_convs:
            if l.weight.device != x.device:
                raise RuntimeError(f"d_convs are on {l.weight.device} but the input is on {x.device}")
            srcs += [l.weight, l.bias]

        def build():
            return dict(wq=w16(q.weight, x.dtype), bq=f32(q.bias), wkv=w16(kv.weight, torch.float16), bkv=f32(kv.bias),
                        wp=w16(p.weight, torch.float16), bp=f32(p.bias), g=f32(n.weight), b=f32(n.bias),
                        dw=[l.weight.detach().float().reshape(C, 9).t().contiguous() for l in d_convs],
                        db=[f32(l.bias) for l in d_convs])
        s = self._stage.get(("w", x.dtype), srcs, build)
        y = torch.empty(B, N, C, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.P2tArgs()
        at = a.attn
        at.dtype, at.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        at.B, at.N, at.C, at.H = B, N, C, self.num_heads
        at.Himg, at.Wimg, at.sr = int(H), int(W), 1
        at.scale = float(self.scale)
        at.x, at.y = ops._ptr(x), ops._ptr(y)
        at.q_weight, at.q_bias = ops._ptr(s["wq"]), ops._ptr(s["bq"])
        at.kv_weight, at.kv_bias = ops._ptr(s["wkv"]), ops._ptr(s["bkv"])
        at.proj_weight, at.proj_bias = ops._ptr(s["wp"]), ops._ptr(s["bp"])
        a.n_levels = len(self.pool_ratios)
        for i, r in enumerate(self.pool_ratios):
            a.pool_h[i], a.pool_w[i] = round(H / r), round(W / r)          # Python's round, like the reference (p2t.py:80)
            a.dconv_weight_t[i] = s["dw"][i].data_ptr()
            a.dconv_bias[i] = s["db"][i].data_ptr() if s["db"][i] is not None else None
        a.norm_weight, a.norm_bias, a.norm_eps = ops._ptr(s["g"]), ops._ptr(s["b"]), float(n.eps)
        ops.run_with_workspace(x, a, "pa_p2t_workspace_bytes", "pa_p2t_fwd")
        return y
