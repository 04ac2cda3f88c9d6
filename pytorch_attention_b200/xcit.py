"""Drop-ins for ``vision_transformers/xcit.py``: ``XCA`` (xcit.py:233-265), ``ClassAttention`` (xcit.py:159-188) and the attention
half of ``XCABlock`` (xcit.py:291)."""
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


def xca_block_attention_half(attn: XCA, norm1: nn.LayerNorm, gamma1, x, out_dtype=None):
    """``x + gamma1 * attn(norm1(x))`` (xcit.py:291, first line of ``XCABlock.forward``) as ONE C-ABI call
    (``pa_xca_block_attn_fwd``): LayerNorm kernel -> qkv GEMM -> cross-covariance core -> proj GEMM with the residual in its
    epilogue.  LayerScale is folded into the staged projection, ``gamma1[:,None] * proj.weight`` and ``gamma1 * proj.bias`` --
    ``x + gamma1 * (O Wp^T + b) == x + O (gamma1 Wp)^T + gamma1 b``.  Usable on a reference ``XCABlock`` whose ``attn`` has been
    swapped for this module's ``XCA``: ``xca_block_attention_half(blk.attn, blk.norm1, blk.gamma1, x)``."""
    x, y_dtype = attn._prepare_input(x)
    check_forward_mode(attn, x, (attn.attn_drop.p, attn.proj_drop.p))
    for name, t in (("norm1.weight", norm1.weight), ("norm1.bias", norm1.bias), ("gamma1", gamma1)):
        if t.device != x.device:
            raise RuntimeError(f"'{name}' is on {t.device} but the input is on {x.device}")
    x = x.contiguous()
    B, N, C = x.shape
    q, p = attn.qkv, attn.proj
    wq, bq, wp, bp, g, b, temp = attn._stage.get(
        "blk", (q.weight, q.bias, p.weight, p.bias, norm1.weight, norm1.bias, gamma1, attn.temperature),
        lambda: (w16(q.weight, torch.float16), f32(q.bias),
                 (gamma1.detach().float()[:, None] * p.weight.detach().float()).to(torch.float16).contiguous(),
                 None if p.bias is None else (gamma1.detach().float() * p.bias.detach().float()).contiguous(),
                 f32(norm1.weight), f32(norm1.bias), attn.temperature.detach().float().reshape(-1).contiguous()))
    y = torch.empty(B, N, C, dtype=out_dtype or attn.out_dtype or y_dtype, device=x.device)
    a = L.XcaBlockArgs()
    a.attn.dtype, a.attn.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
    a.attn.B, a.attn.N, a.attn.C, a.attn.H = B, N, C, attn.num_heads
    a.attn.x, a.attn.y = ops._ptr(x), ops._ptr(y)
    a.attn.qkv_weight, a.attn.qkv_bias = ops._ptr(wq), ops._ptr(bq)
    a.attn.proj_weight, a.attn.proj_bias = ops._ptr(wp), ops._ptr(bp)
    a.attn.temperature = ops._ptr(temp)
    a.ln_weight, a.ln_bias, a.ln_eps = ops._ptr(g), ops._ptr(b), float(norm1.eps)
    ops.run_with_workspace(x, a, "pa_xca_block_attn_workspace_bytes", "pa_xca_block_attn_fwd")
    return y


class XCABlockAttentionHalf(StagedModule):
    """The parameters of the first line of ``XCABlock.forward`` (xcit.py:291) under the reference block's own ``state_dict`` keys
    (``norm1.*``, ``attn.*``, ``gamma1`` -- a subset of XCABlock's: ``load_state_dict(block.state_dict(), strict=False)``), and
    ``forward(x) = x + gamma1 * attn(norm1(x))`` as one C-ABI call.  The LPI and MLP lines of the block (xcit.py:292-293) are not
    on the attention path and stay with the reference's modules."""

    def __init__(self, dim, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., drop=0., norm_layer=nn.LayerNorm, eta=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = XCA(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.gamma1 = nn.Parameter(eta * torch.ones(dim), requires_grad=True)
        self.out_dtype = None
        self._init_stage()

    def attention_half(self, x):
        return xca_block_attention_half(self.attn, self.norm1, self.gamma1, x, out_dtype=self.out_dtype)

    def forward(self, x):
        return self.attention_half(x)
