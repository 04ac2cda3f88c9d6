"""Drop-in for ``vision_transformers/dilateformer.py:GlobalAttention`` (dilateformer.py:136-164).

ViT.Attention's math on a channels-last image: ``forward(x[B, H, W, C]) -> [B, H, W, C]`` attends over the H * W positions
(dilateformer.py:151-162).  The tokens of a contiguous ``[B, H, W, C]`` tensor ARE the ``[B, H*W, C]`` matrix, so the class
shares the ViT forward (co-scheduled single-launch kernel for 64-wide heads and <= 240 positions) through a view."""
from __future__ import annotations

from . import vit


class GlobalAttention(vit.Attention):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        if qk_scale:
            self.scale = qk_scale                  # dilateformer.py:144: qk_scale or head_dim ** -0.5

    def forward(self, x):
        B, H, W, C = x.shape
        return super().forward(x.reshape(B, H * W, C)).reshape(B, H, W, C)
