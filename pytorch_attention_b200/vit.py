"""Drop-in for ``vision_transformers/ViT.py:Attention`` (reference ViT.py:67-89; the identical math in
setr.py:50-72 and moat.py:62-84 is covered by the same class)."""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


class Attention(StagedModule):
    """Same constructor, ``forward(x[B,N,C]) -> [B,N,C]`` contract and ``state_dict`` keys
    (``qkv.weight``, [``qkv.bias``], ``proj.weight``, ``proj.bias``) as the reference class (ViT.py:68-77).

    forward = ONE launch from libpa_b200.so for 64-wide heads and N <= 240 (the co-scheduled kernel: qkv / proj tcgen05 GEMMs
    running under the tcgen05/TMEM softmax chain), otherwise the sequenced single-launch kernel or three stream-ordered
    launches (other head dims: multiples of 16 up to 192, e.g. the reference's default num_heads=4 at dim 768)."""

    def __init__(self, dim, num_heads=4, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.out_dtype = None          # None: same dtype as x.  torch.float16/float32 selectable (bf16 y cannot meet 1e-3)
        self._init_stage()

    def _staged(self, dtype):
        q, p = self.qkv, self.proj
        return self._stage.get(
            ("w", dtype), (q.weight, q.bias, p.weight, p.bias),
            lambda: (w16(q.weight, dtype), f32(q.bias), w16(p.weight, torch.float16), f32(p.bias)))

    def forward(self, x):
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))
        B, N, C = x.shape
        x = x.contiguous()
        wq, bq, wp, bp = self._staged(x.dtype)
        y = torch.empty(B, N, C, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.VitArgs()
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.N, a.C, a.H = B, N, C, self.num_heads
        a.scale = float(self.scale)
        a.x, a.qkv_weight, a.qkv_bias = ops._ptr(x), ops._ptr(wq), ops._ptr(bq)
        a.proj_weight, a.proj_bias, a.y = ops._ptr(wp), ops._ptr(bp), ops._ptr(y)
        a.topk = int(getattr(self, "topk", 0) or 0)      # kvt.KNNAttention sets it (kvt.py:68); 0 = plain softmax
        lib = L.load()
        with torch.cuda.device(x.device):
            need = lib.pa_vit_workspace_bytes(ctypes.byref(a))
            if need == 0:   # invalid arguments: let the entry point report the proper error code
                L.check(lib.pa_vit_fwd(ctypes.byref(a), None, 0, ops.stream_ptr(x.device)))
            ws = ops.workspace(need, x.device)
            L.check(lib.pa_vit_fwd(ctypes.byref(a), ops._ptr(ws), ws.numel(), ops.stream_ptr(x.device)))
        return y



class Mlp(nn.Module):
    """ViT.py:47-65 (outside the attention hot path; kept so that TransformerEncoder is a complete drop-in).  The reference applies
    GELU after fc2 as well."""

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
        return self.drop2(self.gelu(self.fc2(self.drop1(self.gelu(self.fc1(x))))))


class TransformerEncoder(StagedModule):
    """Drop-in for ``ViT.TransformerEncoder`` (ViT.py:105-119): same constructor, sub-module names and ``state_dict`` keys
    (``attn.*``, ``layernorm1.*``, ``mlp.fc1/fc2.*``, ``layernorm2.*``).

    ``attention_half(x)`` = ``x + attn(layernorm1(x))`` (ViT.py:116) as ONE C-ABI call (``pa_vit_block_attn_fwd``): LayerNorm
    kernel -> qkv GEMM -> attention core -> proj GEMM with the residual added in its epilogue.  ``forward`` adds the MLP half
    (ViT.py:117) with the block's own PyTorch modules."""

    def __init__(self, dim, num_heads=4, mlp_ratio=4, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__()
        hidden_features = int(dim * mlp_ratio)
        self.attn = Attention(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        self.layernorm1 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, hidden_features)
        self.layernorm2 = nn.LayerNorm(dim)
        self.out_dtype = None
        self._init_stage()

    def attention_half(self, x):
        x, y_dtype = self._prepare_input(x)
        at = self.attn
        check_forward_mode(self, x, (at.attn_drop.p, at.proj_drop.p))
        B, N, C = x.shape
        x = x.contiguous()
        n1 = self.layernorm1
        wq, bq, wp, bp, g, b = self._stage.get(
            "w", (at.qkv.weight, at.qkv.bias, at.proj.weight, at.proj.bias, n1.weight, n1.bias),
            lambda: (w16(at.qkv.weight, torch.float16), f32(at.qkv.bias), w16(at.proj.weight, torch.float16), f32(at.proj.bias),
                     f32(n1.weight), f32(n1.bias)))
        y = torch.empty(B, N, C, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.VitBlockArgs()
        a.attn.dtype, a.attn.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.attn.B, a.attn.N, a.attn.C, a.attn.H = B, N, C, at.num_heads
        a.attn.scale = float(at.scale)
        a.attn.x, a.attn.qkv_weight, a.attn.qkv_bias = ops._ptr(x), ops._ptr(wq), ops._ptr(bq)
        a.attn.proj_weight, a.attn.proj_bias, a.attn.y = ops._ptr(wp), ops._ptr(bp), ops._ptr(y)
        a.ln_weight, a.ln_bias, a.ln_eps = ops._ptr(g), ops._ptr(b), float(n1.eps)
        ops.run_with_workspace(x, a, "pa_vit_block_attn_workspace_bytes", "pa_vit_block_attn_fwd")
        return y

    def forward(self, x):
        y = self.attention_half(x)
        y = y.to(self.layernorm2.weight.dtype)
        y = y + self.mlp(self.layernorm2(y))
        return y.to(x.dtype)
