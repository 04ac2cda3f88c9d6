"""Drop-in for ``vision_transformers/cmt.py:Attention`` (cmt.py:72-111): PVT's spatial-reduction attention plus an additive
``relative_pos`` on the scaled scores before the softmax."""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from . import pvt


class Attention(pvt.Attention):
    """Same constructor / ``state_dict`` keys as the reference (cmt.py:73-91, identical to pvt.Attention's) and
    ``forward(x[B,N,C], H, W, relative_pos)``: ``softmax(q k^T * scale + relative_pos) v`` (cmt.py:100).  ``relative_pos`` is the
    model's ``[num_heads, N, M]`` parameter (cmt.py:169-180; anything broadcastable to that is expanded once and cached); the
    single-slot attention kernel adds ``relative_pos / scale`` to the raw scores it reads from TMEM in both softmax passes.
    Needs 64-wide heads (cmt_s / cmt_b) and at most 240 keys after the reduction (the zoo's CMT stages have 49)."""

    def _rel_pos(self, relative_pos, N):
        M = N // (self.sr_ratio * self.sr_ratio) if self.sr_ratio > 1 else N

        def build():
            r = relative_pos.detach().float()
            if r.dim() == 4:
                if r.shape[0] != 1:
                    raise ValueError("relative_pos with a batch dimension > 1 is not supported (the reference's CMT passes [heads, N, M])")
                r = r[0]
            return r.expand(self.num_heads, N, M).contiguous()
        return self._stage.get(("rel", N), (relative_pos,), build)

    def forward(self, x, H, W, relative_pos):
        x, y_dtype = self._prepare_input(x)
        self._check(x)
        if relative_pos.device != x.device:
            raise RuntimeError(f"relative_pos is on {relative_pos.device} but the input is on {x.device}")
        B, N, C = x.shape
        x = x.contiguous()
        s = self._staged(x.dtype)
        y = torch.empty(B, N, C, dtype=self.out_dtype or y_dtype, device=x.device)
        a = L.PvtArgs()
        self._fill(a, x, y, H, W, s)
        a.rel_pos = ops._ptr(self._rel_pos(relative_pos, N))
        ops.run_with_workspace(x, a, "pa_pvt_workspace_bytes", "pa_pvt_fwd")
        return y
