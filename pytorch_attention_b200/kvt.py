"""Drop-in for ``vision_transformers/kvt.py:KNNAttention`` (kvt.py:67-94).

ViT.Attention with one extra step: only the ``topk`` largest scores of every query row take part in the softmax, the others are
set to -inf (kvt.py:84-87).  On the B200 path the k-th largest score of every row is found by a selection kernel (exact radix
select, one warp per row) and handed to the attention kernel as a per-row threshold: scores below it are masked right where
they are read from TMEM.  64-wide heads, N <= 240 tokens, ``topk <= N`` (``torch.topk`` raises otherwise, and so does this)."""
from __future__ import annotations

from . import vit


class KNNAttention(vit.Attention):
    def __init__(self, dim, num_heads=4, qkv_bias=False, attn_drop=0, proj_drop=0, topk=100):
        super().__init__(dim, num_heads, qkv_bias, attn_drop, proj_drop)
        self.topk = topk

    def forward(self, x):
        if x.dim() == 3 and self.topk > x.shape[1]:
            raise RuntimeError(f"selected index k out of range: topk={self.topk} > N={x.shape[1]} (torch.topk, kvt.py:85)")
        return super().forward(x)
