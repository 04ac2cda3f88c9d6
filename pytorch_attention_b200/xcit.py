"""Drop-ins for ``vision_transformers/xcit.py``: ``XCA`` (xcit.py:233-265) and ``ClassAttention`` (xcit.py:159-188)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


class _XcitBase(StagedModule):
    def _common_init(self, dim, num_heads, qkv_bias, attn_drop, proj_drop):
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.out_dtype = None
        self._init_stage()

    def _args(self, x, y, extra_srcs=()):
        q, p = self.qkv, self.proj
        wq, bq, wp, bp = self._stage.get(
            ("w", x.dtype), (q.weight, q.bias, p.weight, p.bias),
            lambda: (w16(q.weight, x.dtype), f32(q.bias), w16(p.weight, torch.float16), f32(p.bias)))
        B, N, C = x.shape
        a = L.XcitArgs()
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.N, a.C, a.H = B, N, C, self.num_heads
        a.x, a.y = ops._ptr(x), ops._ptr(y)
        a.qkv_weight, a.qkv_bias, a.proj_weight, a.proj_bias = ops._ptr(wq), ops._ptr(bq), ops._ptr(wp), ops._ptr(bp)
        return a


class XCA(_XcitBase):
    """Cross-covariance attention: channel x channel softmax over L2-normalised q, k with a learned per-head
    temperature; ``qk_scale`` is accepted and ignored exactly like the reference (xcit.py:235-243)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self._common_init(dim, num_heads, qkv_bias, attn_drop, proj_drop)

    def forward(self, x):
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))
        x = x.contiguous()
        y = torch.empty(x.shape, dtype=self.out_dtype or y_dtype, device=x.device)
        a = self._args(x, y)
        temp = self._stage.get("t", (self.temperature,), lambda: self.temperature.detach().float().reshape(-1).contiguous())
        a.temperature = ops._ptr(temp)
        ops.run_with_workspace(x, a, "pa_xca_workspace_bytes", "pa_xca_fwd")
        return y


class ClassAttention(_XcitBase):
    """CLS-query-only attention; patch tokens pass through unchanged (xcit.py:174-188)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self._common_init(dim, num_heads, qkv_bias, attn_drop, proj_drop)

    def forward(self, x):
        x_in = x
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))
        x = x.contiguous()
        y = torch.empty_like(x)
        a = self._args(x, y)
        a.scale = float(self.scale)
        ops.run_with_workspace(x, a, "pa_class_attn_workspace_bytes", "pa_class_attn_fwd")
        if y_dtype != y.dtype:
            # fp32_input mode: the CLS row comes from the 16-bit path, the patch tokens pass through UNCHANGED (xcit.py:187)
            y = y.to(y_dtype)
            y[:, 1:] = x_in[:, 1:]
        return y
