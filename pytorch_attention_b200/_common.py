"""Shared host logic of the drop-in modules: parameter staging and mode checks."""
from __future__ import annotations

import torch
from torch import nn


class ParamStage:
    """Device-side staging of module parameters in the layout/dtype the kernels want (16-bit weights,
    fp32 bias vectors, fused/concatenated matrices).  Entries are rebuilt when the source tensors change
    (in-place update -> ``_version`` bump, or re-assignment -> new ``data_ptr``)."""

    def __init__(self):
        self._entries = {}

    def get(self, key, sources, builder):
        sig = tuple((None if s is None else (s.data_ptr(), s._version, s.dtype, s.device)) for s in sources)
        hit = self._entries.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = builder()
        self._entries[key] = (sig, val)
        return val

    def clear(self):
        self._entries.clear()


def f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def w16(t, dtype):
    return t.detach().to(dtype).contiguous()


def check_forward_mode(module: nn.Module, x: torch.Tensor, drops=()):
    """The reference modules are trainable PyTorch code; this path is forward-only.  Unsupported modes are
    explicit errors, never silent fallbacks."""
    if not x.is_cuda:
        raise RuntimeError("pytorch_attention_b200 runs on sm_100 CUDA devices only (no CPU fallback); got a CPU tensor")
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise ValueError(f"input dtype {x.dtype} unsupported: pass fp16 or bf16 tensors (fp32 accumulation inside)")
    if module.training and any(float(p) > 0.0 for p in drops):
        raise NotImplementedError("dropout p>0 in training mode is not implemented on the B200 forward path")
    if torch.is_grad_enabled() and x.requires_grad:
        raise NotImplementedError("backward is not implemented: call under torch.no_grad() / with inputs that do not require grad")
