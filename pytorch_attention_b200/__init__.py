"""pytorch_attention_b200 — B200-native (sm_100a) attention forward, drop-in for the attention modules of
changzy00/pytorch-attention's vision_transformers/ (ViT, PVT, CvT, CSWin, XCiT).

Python here is plumbing only (module state, device memory, streams); all arithmetic runs in the hand-written
CUDA library ``lib/libpa_b200.so`` behind the C ABI declared in ``include/pa_b200.h``.
"""
from . import _lib, ops  # noqa: F401
from . import vit  # noqa: F401
from .vit import Attention as ViTAttention  # noqa: F401

__all__ = ["ops", "vit", "ViTAttention"]
