"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the glue of the reference's ViT model AROUND its attention modules,
restated with plain torch functional ops so that a parity test can run a whole model with a pluggable attention on a
machine where /root/reference is not mounted (the GPU box).  Not part of the product path.

Follows /root/reference/vision_transformers/ViT.py:
  PatchEmbedding.forward   :86-90    conv(k = stride = patch) -> flatten(2).transpose(1, 2)
  TransformerEncoder.forward :113-116  x + attn(LN1(x));  x + mlp(LN2(x))
  Mlp.forward              :57-64    fc1 -> GELU -> fc2 -> GELU   (the reference applies GELU after fc2 as well)
  VisionTransformer.forward :181-193  tokens = cat([patches, cls_token]) (CLS is appended LAST), + position embedding
                                      (no interpolation when the image has the trained size, :150-153), blocks,
                                      global_pool == "token" takes x[:, 0], head.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def vit_model_forward(sd, x, attn_fn, patch_size: int, depths: int, eps: float = 1e-5):
    """sd: the reference VisionTransformer state_dict (tensors on x's device, fp32); attn_fn(block_index, tokens) -> tokens
    stands for `blocks[i].attn`; returns the logits."""
    t = F.conv2d(x, sd["patch_embedding.proj.weight"], sd["patch_embedding.proj.bias"], stride=patch_size)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([t, sd["cls_token"].expand(t.shape[0], -1, -1)], dim=1)
    assert t.shape[1] == sd["position_embedding"].shape[1], "glue restates only the non-interpolating case (ViT.py:150-153)"
    t = t + sd["position_embedding"]
    C = t.shape[-1]
    for i in range(depths):
        p = f"blocks.{i}."
        h = F.layer_norm(t, (C,), sd[p + "layernorm1.weight"], sd[p + "layernorm1.bias"], eps)
        t = t + attn_fn(i, h)
        h = F.layer_norm(t, (C,), sd[p + "layernorm2.weight"], sd[p + "layernorm2.bias"], eps)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        h = F.gelu(F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]))
        t = t + h
    return F.linear(t[:, 0], sd["head.weight"], sd["head.bias"])


def attention_state(sd, i):
    """The sub-state_dict of blocks[i].attn with the keys the Attention module itself uses."""
    p = f"blocks.{i}.attn."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def vit_state_shapes(image_size=224, patch_size=16, in_channels=3, depths=12, num_heads=4, mlp_ratio=4, embedding_dim=768,
                     qkv_bias=False, num_classes=1000, **_unused):
    """Keys and shapes of the reference ``ViT.VisionTransformer(...)`` state_dict for a configuration (defaults = the
    reference's own defaults, ViT.py:121-134).  oracle/make_golden_model.py asserts it against the live model."""
    D, P = embedding_dim, patch_size
    n = (image_size // patch_size) ** 2
    hid = int(D * mlp_ratio)
    sh = {"cls_token": (1, 1, D), "position_embedding": (1, n + 1, D),
          "patch_embedding.proj.weight": (D, in_channels, P, P), "patch_embedding.proj.bias": (D,)}
    for i in range(depths):
        p = f"blocks.{i}."
        sh[p + "attn.qkv.weight"] = (3 * D, D)
        if qkv_bias:
            sh[p + "attn.qkv.bias"] = (3 * D,)
        sh[p + "attn.proj.weight"] = (D, D)
        sh[p + "attn.proj.bias"] = (D,)
        for ln in ("layernorm1", "layernorm2"):
            sh[p + ln + ".weight"] = (D,)
            sh[p + ln + ".bias"] = (D,)
        sh[p + "mlp.fc1.weight"] = (hid, D)
        sh[p + "mlp.fc1.bias"] = (hid,)
        sh[p + "mlp.fc2.weight"] = (D, hid)
        sh[p + "mlp.fc2.bias"] = (D,)
    sh["head.weight"] = (num_classes, D)
    sh["head.bias"] = (num_classes,)
    return sh


def synth_state_dict(shapes, seed):
    """Deterministic parameters from a seed (one CPU generator per tensor, in sorted key order), rounded once to
    fp16-representable values: lets a whole-model golden be stored as logits only (an 86 M-parameter ViT-B would not fit a
    fixture) -- generator and test rebuild the identical state_dict."""
    sd = {}
    for idx, key in enumerate(sorted(shapes)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        shape = shapes[key]
        t = torch.randn(shape, generator=g)
        if key.endswith("layernorm1.weight") or key.endswith("layernorm2.weight"):
            t = 1.0 + 0.1 * t
        elif key in ("cls_token", "position_embedding"):
            t = 0.1 * t
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.0 / fan_in ** 0.5)
        else:
            t = 0.05 * t
        sd[key] = t.half().float()
    return sd
