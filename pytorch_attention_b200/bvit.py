"""Drop-in for ``vision_transformers/bvit.py:Broad_Attention`` (bvit.py:49-76)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16


class Broad_Attention(StagedModule):
    """Same constructor / ``state_dict`` keys as the reference (bvit.py:50-64): ``to_qkv`` Linear(dim, 3 * heads * dim_head,
    bias=False) and ``to_out = Sequential(Linear(inner, dim), Dropout)`` -- or ``nn.Identity`` when ``heads == 1 and
    dim_head == dim``.  ``forward(x)`` returns ``(out, q, k, v)`` like the reference (bvit.py:66-76): q, k, v are
    ``[B, heads, N, dim_head]`` VIEWS of the fp16 projection this call wrote (the reference's are fp32 views of its own
    projection); BViT's broad attention reads them from every layer (bvit.py:88-98)."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout)) if project_out else nn.Identity()
        self.out_dtype = None
        self._init_stage()

    def forward(self, x):
        x, y_dtype = self._prepare_input(x)
        proj = self.to_out[0] if isinstance(self.to_out, nn.Sequential) else None
        check_forward_mode(self, x, (self.to_out[1].p,) if proj is not None else ())
        B, N, C = x.shape
        x = x.contiguous()
        inner = self.heads * self.dim_head
        srcs = (self.to_qkv.weight,) + ((proj.weight, proj.bias) if proj is not None else ())
        wq, wo, bo = self._stage.get(
            ("w", x.dtype), srcs,
            lambda: (w16(self.to_qkv.weight, x.dtype),
                     w16(proj.weight, torch.float16) if proj is not None else None,
                     f32(proj.bias) if proj is not None else None))
        y = torch.empty(B, N, C, dtype=(self.out_dtype or y_dtype) if proj is not None else torch.float16, device=x.device)
        qkv = torch.empty(B, N, 3 * inner, dtype=torch.float16, device=x.device)
        a = L.BvitArgs()
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.N, a.C, a.H, a.dim_head = B, N, C, self.heads, self.dim_head
        a.scale = float(self.scale)
        a.x, a.qkv_weight, a.out_weight, a.out_bias = ops._ptr(x), ops._ptr(wq), ops._ptr(wo), ops._ptr(bo)
        a.qkv, a.y = ops._ptr(qkv), ops._ptr(y)
        ops.run_with_workspace(x, a, "pa_bvit_workspace_bytes", "pa_bvit_fwd")
        if proj is None and (self.out_dtype or y_dtype) != y.dtype:
            y = y.to(self.out_dtype or y_dtype)
        q, k, v = qkv.view(B, N, 3, self.heads, self.dim_head).permute(2, 0, 3, 1, 4).unbind(0)   # 'b n (h d) -> b h n d'
        return y, q, k, v
