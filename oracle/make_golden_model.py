#!/usr/bin/env python
"""Generate tests/golden/models/vit_tiny.npz from the LIVE reference model (build container only).

A whole reference `ViT.VisionTransformer` (2 blocks, dim 128, 2 heads of 64, 64x64 images in 16x16 patches -> 17 tokens,
10 classes) with every parameter randomised and rounded once to fp16-representable values is run in fp32 on CPU.  The
script checks that the restated glue (oracle/model_glue.py) with the oracle attention reproduces the reference logits,
then stores parameters, input and logits.  tests/test_model_dropin.py re-runs the glue with the B200 drop-in attention
loaded from the same `blocks.{i}.attn.*` entries.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PA_REFERENCE", "/root/reference/vision_transformers")
CFG = dict(image_size=64, patch_size=16, in_channels=3, depths=2, num_heads=2, embedding_dim=128, num_classes=10)


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not mounted at {REF}")
    sys.path.insert(0, REF)
    import ViT  # the reference module
    from oracle import vit_attention
    from oracle.model_glue import vit_model_forward, attention_state

    torch.manual_seed(1234)
    model = ViT.VisionTransformer(**CFG).eval()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("layernorm1.weight") or name.endswith("layernorm2.weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif p.dim() >= 2 and "embedding" not in name and "cls_token" not in name:
                fan_in = p[0].numel()
                p.copy_(torch.randn_like(p) * (1.5 / fan_in ** 0.5))
            elif p.dim() >= 2:
                p.copy_(0.5 * torch.randn_like(p))
            else:
                p.copy_(0.1 * torch.randn_like(p))
            p.copy_(p.half().float())
        x = torch.randn(3, CFG["in_channels"], CFG["image_size"], CFG["image_size"]).half().float()
        logits = model(x)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        H = CFG["num_heads"]

        def oracle_attn(i, t):
            a = attention_state(sd, i)
            return vit_attention(t, a["qkv.weight"], a.get("qkv.bias"), a["proj.weight"], a["proj.bias"], H)

        glue = vit_model_forward(sd, x, oracle_attn, CFG["patch_size"], CFG["depths"])
    err = (glue - logits).abs().max().item()
    print(f"logits{tuple(logits.shape)} max|y|={logits.abs().max():.4f}  glue+oracle vs reference model max-abs = {err:.3e}")
    assert err <= 2e-5 * max(1.0, logits.abs().max().item())
    blob = {"p." + k: v.numpy().astype(np.float16) for k, v in sd.items()}
    blob["in.x"] = x.numpy().astype(np.float16)
    blob["y_ref"] = logits.numpy().astype(np.float32)
    out = os.path.join(ROOT, "tests", "golden", "models")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "vit_tiny.npz"), **blob)
    print("wrote", os.path.join(out, "vit_tiny.npz"))


DEFAULT_SEED = 77


def main_default():
    """tests/golden/models/vit_default.npz: the reference's ``VisionTransformer()`` with EVERY constructor default (224x224,
    patch 16, 12 blocks, dim 768, 4 heads of 192, 1000 classes -- README.md:331-334) on two images.  Parameters come from
    oracle.model_glue.synth_state_dict (seeded), so only the logits are stored."""
    sys.path.insert(0, REF)
    import ViT
    from oracle import vit_attention
    from oracle.model_glue import vit_model_forward, attention_state, vit_state_shapes, synth_state_dict
    model = ViT.VisionTransformer().eval()
    shapes = vit_state_shapes()
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == shapes, "restated shapes differ from the live model"
    sd = synth_state_dict(shapes, DEFAULT_SEED)
    model.load_state_dict(sd, strict=True)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(DEFAULT_SEED)).half().float()
    with torch.no_grad():
        logits = model(x)

        def oracle_attn(i, t):
            a = attention_state(sd, i)
            return vit_attention(t, a["qkv.weight"], a.get("qkv.bias"), a["proj.weight"], a["proj.bias"], 4)
        glue = vit_model_forward(sd, x, oracle_attn, 16, 12)
    err = (glue - logits).abs().max().item()
    print(f"default VisionTransformer(): logits{tuple(logits.shape)} max|y|={logits.abs().max():.4f}  glue+oracle vs reference max-abs = {err:.3e}")
    assert err <= 5e-5 * max(1.0, logits.abs().max().item())
    out = os.path.join(ROOT, "tests", "golden", "models", "vit_default.npz")
    np.savez_compressed(out, y_ref=logits.numpy().astype(np.float32), seed=np.int64(DEFAULT_SEED))
    print("wrote", out)


if __name__ == "__main__":
    main()
    main_default()
