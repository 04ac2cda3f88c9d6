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


class StagedModule(nn.Module):
    """Base of the drop-ins: owns the ParamStage and drops it whenever the parameters may have changed behind its back --
    ``load_state_dict`` and every ``_apply`` (``.to()``, ``.cuda()``, ``.half()``, ...).  In-place edits through ``.data``
    (``w.data.copy_()``, some EMA / init code) do not bump a tensor's version counter: call ``refresh()`` after those.

    ``fp32_input``: the reference's forward takes fp32 tensors (ViT.py:79).  By default fp32 input is an error (the B200 path
    computes on 16-bit operands and says so); set ``module.fp32_input = torch.float16`` (or ``torch.bfloat16``) to opt in:
    x is cast by ``pa_cast_f32`` (this library's kernel) and y comes back in fp32 unless ``out_dtype`` says otherwise."""

    fp32_input = None

    def _init_stage(self):
        self._stage = ParamStage()
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module._stage.clear())

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if "_stage" in self.__dict__:
            self._stage.clear()
        return out

    def refresh(self):
        """Forget the staged (16-bit / fused / BN-folded) copies of the parameters; the next forward rebuilds them."""
        for m in self.modules():                 # self and every descendant (children may sit inside ModuleLists)
            if isinstance(m, StagedModule) and "_stage" in m.__dict__:
                m._stage.clear()

    def _prepare_input(self, x):
        """Returns (x16, default output dtype).  fp32 activations are cast when the module opted in (see the class docstring)."""
        if x.dtype == torch.float32 and self.fp32_input is not None and x.is_cuda:
            from . import ops
            return ops.cast_f32(x, self.fp32_input), torch.float32
        return x, x.dtype


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
        raise ValueError(f"input dtype {x.dtype} unsupported: pass fp16 or bf16 tensors (fp32 accumulation inside), or opt in to "
                         "the cast with module.fp32_input = torch.float16")
    # raw device pointers of the parameters go into TMA descriptors: a module left on the CPU or on another GPU must be a
    # clean error here (the reference raises a device-mismatch error), not an illegal address inside a kernel
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        if t.device != x.device:
            raise RuntimeError(f"parameter/buffer '{name}' is on {t.device} but the input is on {x.device}: "
                               "move the module with .to(x.device) first")
    if module.training and any(float(p) > 0.0 for p in drops):
        raise NotImplementedError("dropout p>0 in training mode is not implemented on the B200 forward path")
    if torch.is_grad_enabled() and x.requires_grad:
        raise NotImplementedError("backward is not implemented: call under torch.no_grad() / with inputs that do not require grad")
