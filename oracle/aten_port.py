"""CPU *baseline* port of ViT.Attention.forward (ViT.py:79-89) using the same ATen operator sequence the
reference executes (linear -> reshape/permute -> matmul * scale -> softmax -> matmul -> linear), so that the CPU
timing beside the GPU numbers reflects what the reference itself costs on the host cores.  TEST/BENCH
INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference legs; tests check it against the einsum oracle).
The einsum restatement in oracle/attention.py stays the parity oracle; this port only exists because einsum picks
slow contraction paths and would under-state the reference's CPU speed."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def vit_attention_aten(x, qkv_weight, qkv_bias, proj_weight, proj_bias, num_heads):
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, qkv_weight, qkv_bias).view(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5)
    a = a.softmax(dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, proj_weight, proj_bias)
