"""Drop-in for ``vision_transformers/pvt.py:Attention`` (spatial-reduction attention, pvt.py:52-91)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


def fold_bn(conv_bias, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm after a conv with bias == per-channel (scale, shift)."""
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    cb = conv_bias.detach().float() if conv_bias is not None else torch.zeros_like(scale)
    shift = (cb - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return scale.contiguous(), shift.contiguous()


def _cat_bias(*bs):
    if all(b is None for b in bs):
        return None
    return torch.cat([b.detach().float() for b in bs]).contiguous()


class Attention(StagedModule):
    """Same constructor / ``forward(x[B,N,C], H, W)`` / ``state_dict`` keys as the reference (pvt.py:53-71):
    ``q,k,v,proj`` Linear and, for sr_ratio > 1, ``sr = Sequential(Conv2d(depthwise, k=s=sr), BatchNorm2d)``.
    Launch sequence: [sr conv+BN kernel] -> q GEMM -> fused [k|v] GEMM -> attention core -> proj GEMM."""

    def __init__(self, dim, num_heads=8, sr_ratio=1, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.sr_ratio = sr_ratio
        if self.sr_ratio > 1:
            self.sr = nn.Sequential(
                nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio, groups=dim),
                nn.BatchNorm2d(dim))
        self.out_dtype = None
        self._init_stage()

    def _staged(self, dtype):
        srcs = [self.q.weight, self.q.bias, self.k.weight, self.k.bias, self.v.weight, self.v.bias,
                self.proj.weight, self.proj.bias]
        if self.sr_ratio > 1:
            bn = self.sr[1]
            srcs += [self.sr[0].weight, self.sr[0].bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]

        def build():
            kv_dtype = torch.float16 if self.sr_ratio > 1 else dtype
            d = dict(wq=w16(self.q.weight, dtype), bq=f32(self.q.bias),
                     wkv=torch.cat([self.k.weight.detach(), self.v.weight.detach()]).to(kv_dtype).contiguous(),
                     bkv=_cat_bias(self.k.bias, self.v.bias),
                     wp=w16(self.proj.weight, torch.float16), bp=f32(self.proj.bias), srw=None, srs=None, srb=None)
            if self.sr_ratio > 1:
                sr = self.sr_ratio
                d["srw"] = self.sr[0].weight.detach().float().reshape(-1, sr * sr).t().contiguous()   # [sr*sr, C]
                d["srs"], d["srb"] = fold_bn(self.sr[0].bias, self.sr[1])
            return d
        return self._stage.get(("w", dtype), srcs, build)

    def _fill(self, a, x, y, H, W, s):
        """Fill a PvtArgs for this module's parameters (shared by the stand-alone forward and the Block entry point)."""
        B, N, C = x.shape
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.N, a.C, a.H = B, N, C, self.num_heads
        a.Himg, a.Wimg, a.sr = int(H), int(W), self.sr_ratio
        a.scale = float(self.scale)
        a.x, a.y = ops._ptr(x), ops._ptr(y)
        a.q_weight, a.q_bias = ops._ptr(s["wq"]), ops._ptr(s["bq"])
        a.kv_weight, a.kv_bias = ops._ptr(s["wkv"]), ops._ptr(s["bkv"])
        a.proj_weight, a.proj_bias = ops._ptr(s["wp"]), ops._ptr(s["bp"])
        a.sr_weight_t, a.sr_scale, a.sr_shift = ops._ptr(s["srw"]), ops._ptr(s["srs"]), ops._ptr(s["srb"])

    def _check(self, x):
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))
        if self.sr_ratio > 1 and self.training:
            raise NotImplementedError("train-mode BatchNorm (batch statistics) is not implemented: call .eval()")

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


def block_attention_half(block, attn, norm1, x, H, W, rel_pos=None):
    """``x + attn(norm1(x), H, W[, relative_pos])`` (pvt.py:106, segformer.py:76, cmt.py:131) as ONE C-ABI call
    (``pa_pvt_block_attn_fwd``): LayerNorm kernel -> [reduction] -> q / [k|v] GEMMs -> attention core -> proj GEMM with the
    residual x added in its epilogue.  The projections read the fp16 LayerNorm output, so their weights are staged as fp16."""
    x, y_dtype = block._prepare_input(x)
    attn._check(x)
    check_forward_mode(block, x)
    B, N, C = x.shape
    x = x.contiguous()
    s = attn._staged(torch.float16)
    g, b = block._stage.get("ln", (norm1.weight, norm1.bias), lambda: (f32(norm1.weight), f32(norm1.bias)))
    y = torch.empty(B, N, C, dtype=block.out_dtype or y_dtype, device=x.device)
    a = L.PvtBlockArgs()
    attn._fill(a.attn, x, y, H, W, s)
    if rel_pos is not None:
        a.attn.rel_pos = ops._ptr(attn._rel_pos(rel_pos, N))
    a.ln_weight, a.ln_bias, a.ln_eps = ops._ptr(g), ops._ptr(b), float(norm1.eps)
    ops.run_with_workspace(x, a, "pa_pvt_block_attn_workspace_bytes", "pa_pvt_block_attn_fwd")
    return y


class Mlp(nn.Module):
    """pvt.py:13-31 (outside the attention hot path; kept so that Block is a complete drop-in)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, drop=0):
        super().__init__()
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.gelu = nn.GELU()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.gelu(self.fc1(x)))))


class Block(StagedModule):
    """Drop-in for ``pvt.Block`` (pvt.py:93-108): same constructor, sub-module names and ``state_dict`` keys (``norm1.*``,
    ``attn.*``, ``norm2.*``, ``mlp.fc1/fc2.*``).  ``attention_half(x, H, W)`` = ``x + attn(norm1(x), H, W)`` (pvt.py:106) as one
    C-ABI call; ``forward`` adds the MLP half (pvt.py:107) with the block's own PyTorch modules."""

    def __init__(self, dim, num_heads=8, mlp_ratio=4, sr_ratio=1, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, num_heads, sr_ratio, qkv_bias, attn_drop, proj_drop)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))
        self.out_dtype = None
        self._init_stage()

    def attention_half(self, x, H, W):
        return block_attention_half(self, self.attn, self.norm1, x, H, W)

    def forward(self, x, H, W):
        y = self.attention_half(x, H, W)
        y = y.to(self.norm2.weight.dtype)
        y = y + self.mlp(self.norm2(y))
        return y.to(x.dtype)
