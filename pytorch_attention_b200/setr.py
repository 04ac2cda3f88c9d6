"""Drop-in for ``vision_transformers/setr.py:Attention`` (reference setr.py:50-72).

The math is ViT.Attention's (ViT.py:79-89) line for line -- fused qkv Linear, ``reshape(B,N,3,H,hd)``, softmax(q k^T scale) v,
proj -- so the class shares the ViT forward (co-scheduled single-launch kernel for 64-wide heads); only the constructor default
differs: ``num_heads=8``.  ``state_dict`` keys: ``qkv.weight``, [``qkv.bias``], ``proj.weight``, ``proj.bias``."""
from __future__ import annotations

from . import vit


class Attention(vit.Attention):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0, proj_drop=0):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
