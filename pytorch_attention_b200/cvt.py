"""Drop-in for ``vision_transformers/cvt.py:Attention`` (convolutional-projection attention, cvt.py:48-76)."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import ops
from ._common import StagedModule, check_forward_mode, f32, w16
from .pvt import fold_bn


class Attention(StagedModule):
    """NCHW in / NCHW out, same constructor and ``state_dict`` keys as the reference (cvt.py:49-62).
    Launch sequence: depthwise-conv+BN (NCHW -> token-major) -> qkv GEMM -> attention core -> proj GEMM that
    writes NCHW directly (y[b] = Wp . O[b]^T, bias per row)."""

    def __init__(self, dim, num_heads=8, ks=3, attn_drop=0, proj_drop=0):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.ks = ks
        self.conv_proj_qkv = nn.Sequential(
            nn.Conv2d(dim, dim, kernel_size=ks, stride=1, padding=(ks - 1) // 2, groups=dim),
            nn.BatchNorm2d(dim),
            nn.Conv2d(dim, 3 * dim, 1))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Conv2d(dim, dim, 1)
        self.proj_drop = nn.Dropout(proj_drop)
        self.out_dtype = None
        self._init_stage()

    def _staged(self):
        dw, bn, pw = self.conv_proj_qkv[0], self.conv_proj_qkv[1], self.conv_proj_qkv[2]
        srcs = [dw.weight, dw.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, pw.weight, pw.bias,
                self.proj.weight, self.proj.bias]

        def build():
            C = dw.weight.shape[0]
            scale, shift = fold_bn(dw.bias, bn)
            return dict(dww=dw.weight.detach().float().reshape(C, -1).contiguous(), dws=scale, dwb=shift,
                        wqkv=w16(pw.weight.reshape(3 * C, C), torch.float16), bqkv=f32(pw.bias),
                        wp=w16(self.proj.weight.reshape(C, C), torch.float16), bp=f32(self.proj.bias))
        return self._stage.get("w", srcs, build)

    def forward(self, x):
        x, y_dtype = self._prepare_input(x)
        check_forward_mode(self, x, (self.attn_drop.p, self.proj_drop.p))
        if self.training:
            raise NotImplementedError("train-mode BatchNorm (batch statistics) is not implemented: call .eval()")
        B, C, H, W = x.shape
        x = x.contiguous()
        s = self._staged()
        y = torch.empty(B, C, H, W, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.CvtArgs()
        a.dtype, a.out_dtype = ops.dtype_code(x.dtype), ops.dtype_code(y.dtype)
        a.B, a.C, a.H, a.Himg, a.Wimg, a.ks = B, C, self.num_heads, H, W, self.ks
        a.scale = float(self.scale)
        a.x, a.y = ops._ptr(x), ops._ptr(y)
        a.dw_weight, a.dw_scale, a.dw_shift = ops._ptr(s["dww"]), ops._ptr(s["dws"]), ops._ptr(s["dwb"])
        a.qkv_weight, a.qkv_bias = ops._ptr(s["wqkv"]), ops._ptr(s["bqkv"])
        a.proj_weight, a.proj_bias = ops._ptr(s["wp"]), ops._ptr(s["bp"])
        ops.run_with_workspace(x, a, "pa_cvt_workspace_bytes", "pa_cvt_fwd")
        return y
