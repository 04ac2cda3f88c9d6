"""Drop-in for ``vision_transformers/moat.py:Attention`` (reference moat.py:62-84), the attention inside ``MOATBlock``
(moat.py:86-110, called on the flattened ``[B, H*W, C]`` token view).

Same math as ViT.Attention (ViT.py:79-89) with the constructor default ``num_heads=8``; shares the ViT forward."""
from __future__ import annotations

from . import vit


class Attention(vit.Attention):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
