"""Drop-ins for ``vision_transformers/cswin.py``: ``LePEAttention`` (cswin.py:51-127) and the attention half of
``CSWinBlock`` (cswin.py:130-194).  The MLP half of the block (cswin.py:195) is outside the hot path and stays
plain PyTorch modules so that the block remains a complete drop-in."""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


def _get_v_t(conv):
    C = conv.weight.shape[0]
    return conv.weight.detach().float().reshape(C, 9).t().contiguous()     # [9, C]


class LePEAttention(StagedModule):
    """Cross-shaped-window attention with locally-enhanced positional encoding.  ``forward(qkv[3,B,L,C])``: the
    three slices may be strided views (the block passes channel halves of one buffer, cswin.py:188-189).
    Launches: LePE depthwise-3x3 kernel, then the windowed tcgen05 attention core that gathers windows by TMA and
    accumulates onto the LePE term."""

    def __init__(self, dim, resolution, idx, split_size=7, dim_out=None, num_heads=8, attn_drop=0., proj_drop=0.,
                 qk_scale=None):
        super().__init__()
        self.dim = dim
        self.dim_out = dim_out or dim
        self.resolution = resolution
        self.split_size = split_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        if idx == -1:
            H_sp, W_sp = self.resolution, self.resolution
        elif idx == 0:
            H_sp, W_sp = self.resolution, self.split_size
        elif idx == 1:
            W_sp, H_sp = self.resolution, self.split_size
        else:
            # the reference prints "ERROR MODE" and calls exit(0) (cswin.py:68-70); a library raises instead
            raise ValueError(f"ERROR MODE {idx}")
        self.idx = idx
        self.H_sp = H_sp
        self.W_sp = W_sp
        self.get_v = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, groups=dim)
        self.attn_drop = nn.Dropout(attn_drop)
        self._init_stage()

    def staged_get_v(self):
        return self._stage.get("v", (self.get_v.weight, self.get_v.bias), lambda: (_get_v_t(self.get_v), f32(self.get_v.bias)))

    def forward(self, qkv):
        q, k, v = qkv[0], qkv[1], qkv[2]
        check_forward_mode(self, q, (self.attn_drop.p,))
        if q.dtype != torch.float16:
            raise ValueError("LePEAttention: q/k/v must be fp16 (CSWinBlock produces them in fp16 from fp16/bf16 x)")
        B, Lt, C = q.shape
        assert Lt == self.resolution * self.resolution, "flatten img_tokens has wrong size"
        for t in (q, k, v):
            if t.stride(2) != 1 or t.stride(0) != Lt * t.stride(1) or t.stride() != q.stride():
                q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
                break
        wt, bias = self.staged_get_v()
        out = torch.empty(B, Lt, C, dtype=torch.float16, device=q.device)
        a = L.LepeArgs()
        a.B, a.L, a.C, a.H = B, Lt, C, self.num_heads
        a.resolution, a.idx, a.split_size = self.resolution, self.idx, self.split_size
        a.scale = float(self.scale)
        a.q, a.k, a.v = ops._ptr(q), ops._ptr(k), ops._ptr(v)
        a.ld, a.batch_stride = q.stride(1), q.stride(0)
        a.get_v_weight_t, a.get_v_bias = ops._ptr(wt), ops._ptr(bias)
        a.out, a.ldo, a.out_batch_stride = ops._ptr(out), out.stride(1), out.stride(0)
        with torch.cuda.device(q.device):
            L.check(L.load().pa_cswin_lepe_fwd(ctypes.byref(a), ops.stream_ptr(q.device)))
        return out


class Mlp(nn.Module):
    """cswin.py:33-49 (outside the attention hot path; kept so CSWinBlock is a complete drop-in)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class CSWinBlock(StagedModule):
    """Same constructor and ``state_dict`` keys as the reference block (cswin.py:132-174).  ``attention_half(x)``
    is the B200 path (norm1 -> qkv -> two LePE branches -> proj -> residual, cswin.py:181-194) as ONE C-ABI call;
    ``forward`` adds the MLP half with ordinary PyTorch modules (cswin.py:195)."""

    def __init__(self, dim, reso, num_heads, split_size=7, mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop=0., attn_drop=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, last_stage=False):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.patches_resolution = reso
        self.split_size = split_size
        self.mlp_ratio = mlp_ratio
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.norm1 = norm_layer(dim)
        if self.patches_resolution == split_size:
            last_stage = True
        self.last_stage = last_stage
        self.branch_num = 1 if last_stage else 2
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(drop)
        if last_stage:
            self.attns = nn.ModuleList([
                LePEAttention(dim, resolution=self.patches_resolution, idx=-1, split_size=split_size,
                              num_heads=num_heads, dim_out=dim, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
                for i in range(self.branch_num)])
        else:
            self.attns = nn.ModuleList([
                LePEAttention(dim // 2, resolution=self.patches_resolution, idx=i, split_size=split_size,
                              num_heads=num_heads // 2, dim_out=dim // 2, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
                for i in range(self.branch_num)])
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, out_features=dim, act_layer=act_layer, drop=drop)
        self.norm2 = norm_layer(dim)
        self.out_dtype = None
        self._init_stage()

    def attention_half(self, x, residual=True):
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x, (self.proj_drop.p, self.attns[0].attn_drop.p))
        if not isinstance(self.norm1, nn.LayerNorm):
            raise NotImplementedError("only norm_layer=nn.LayerNorm is implemented on the B200 path")
        B, Lt, C = x.shape
        assert Lt == self.patches_resolution ** 2, "flatten img_tokens has wrong size"
        x = x.contiguous()
        n1, q, p = self.norm1, self.qkv, self.proj
        wq, bq, wp, bp, g, b = self._stage.get(
            "w", (q.weight, q.bias, p.weight, p.bias, n1.weight, n1.bias),
            lambda: (w16(q.weight, torch.float16), f32(q.bias), w16(p.weight, torch.float16), f32(p.bias),
                     f32(n1.weight), f32(n1.bias)))
        y = torch.empty(B, Lt, C, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.CswinBlockArgs()
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.L, a.C, a.H = B, Lt, C, self.num_heads
        a.reso, a.split_size, a.last_stage, a.residual = self.patches_resolution, self.split_size, int(self.last_stage), int(residual)
        a.scale, a.ln_eps = float(self.attns[0].scale), float(n1.eps)
        a.x, a.y = ops._ptr(x), ops._ptr(y)
        a.norm1_weight, a.norm1_bias = ops._ptr(g), ops._ptr(b)
        a.qkv_weight, a.qkv_bias, a.proj_weight, a.proj_bias = ops._ptr(wq), ops._ptr(bq), ops._ptr(wp), ops._ptr(bp)
        keep = []
        for i, att in enumerate(self.attns):
            wt, bias = att.staged_get_v()
            keep.append((wt, bias))
            a.get_v_weight_t[i], a.get_v_bias[i] = wt.data_ptr(), bias.data_ptr()
        ops.run_with_workspace(x, a, "pa_cswin_block_attn_workspace_bytes", "pa_cswin_block_attn_fwd")
        return y

    def forward(self, x):
        """cswin.py:176-197.  The attention half is the B200 path; the MLP half (cswin.py:195, outside the hot path) runs the
        block's own PyTorch modules in THEIR parameter dtype, so the block works in a 16-bit model as well as in an fp32 model
        whose attention half is fed 16-bit activations (or fp32 ones with ``fp32_input`` set); the result has x's dtype."""
        y = self.attention_half(x, residual=True)
        y = y.to(self.norm2.weight.dtype)
        y = y + self.mlp(self.norm2(y))
        return y.to(x.dtype)
