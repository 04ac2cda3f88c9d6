"""Drop-in for ``attention_mechanisms/dual_attention.py:PAM`` (dual_attention.py:12-28)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32


class PAM(StagedModule):
    """Position attention module of DANet: same constructor / ``state_dict`` keys as the reference (``b``, ``c``, ``d`` 1x1 convs,
    ``alpha``) and ``forward(x[n,c,h,w]) -> alpha * attn(x) + x`` with ``attn = softmax(b(x)^T c(x)) d(x)`` -- ONE head as wide
    as the channel count, NO score scale (dual_attention.py:21-27).

    It runs through the CvT entry point (``pa_cvt_fwd``): the NCHW -> token-major kernel with a unit tap, the three 1x1 convs
    as one fused ``[b|c|d]`` GEMM, the attention core (online softmax over h*w keys), and the NCHW-writing batched GEMM with
    ``alpha * I`` as its weight and x as the residual of its epilogue -- the transposition back to NCHW, the scaling by alpha and
    the skip connection are that one GEMM.  Channel counts: multiples of 16 from 32 to 192."""

    def __init__(self, dim):
        super().__init__()
        self.b = nn.Conv2d(dim, dim, 1)
        self.c = nn.Conv2d(dim, dim, 1)
        self.d = nn.Conv2d(dim, dim, 1)
        self.alpha = nn.Parameter(torch.zeros(1))
        self.out_dtype = None
        self._init_stage()

    def _staged(self):
        srcs = [self.b.weight, self.b.bias, self.c.weight, self.c.bias, self.d.weight, self.d.bias, self.alpha]

        def build():
            C = self.b.weight.shape[0]
            dev = self.b.weight.device
            w = torch.cat([m.weight.detach().reshape(C, C) for m in (self.b, self.c, self.d)]).to(torch.float16).contiguous()
            bias = torch.cat([m.bias.detach().float() for m in (self.b, self.c, self.d)]).contiguous()
            eye = (self.alpha.detach().float() * torch.eye(C, device=dev)).to(torch.float16).contiguous()
            return dict(tap=torch.ones(C, 1, device=dev), one=torch.ones(C, device=dev), zero=torch.zeros(C, device=dev),
                        wqkv=w, bqkv=bias, wp=eye)
        return self._stage.get("w", srcs, build)

    def forward(self, x):
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x)
        n, c, h, w = x.shape
        x = x.contiguous()
        s = self._staged()
        y = torch.empty(n, c, h, w, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.CvtArgs()
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.C, a.H, a.Himg, a.Wimg, a.ks = n, c, 1, h, w, 1
        a.scale = 1.0                                   # (B @ C).softmax: no head_dim ** -0.5 (dual_attention.py:25)
        a.x, a.y, a.residual = ops._ptr(x), ops._ptr(y), ops._ptr(x)
        a.dw_weight, a.dw_scale, a.dw_shift = ops._ptr(s["tap"]), ops._ptr(s["one"]), ops._ptr(s["zero"])
        a.qkv_weight, a.qkv_bias = ops._ptr(s["wqkv"]), ops._ptr(s["bqkv"])
        a.proj_weight, a.proj_bias = ops._ptr(s["wp"]), None
        ops.run_with_workspace(x, a, "pa_cvt_workspace_bytes", "pa_cvt_fwd")
        return y
