"""The drop-in attention inside a WHOLE reference model (SURVEY.md §8b "what calls it").

tests/golden/models/vit_tiny.npz holds the parameters, an input batch and the logits of the real reference
`ViT.VisionTransformer` (oracle/make_golden_model.py).  The model glue around the attention modules is restated in
oracle/model_glue.py, so the same model can be evaluated where /root/reference is absent:
  * CPU: glue + oracle attention reproduces the reference logits (pins the glue);
  * CPU, reference mounted: the drop-in class substituted into the reference's own model constructs and loads the
    reference state_dict strictly (constructor signature + keys), and refuses to run on CPU;
  * GPU: glue + B200 drop-in attention (fp16 I/O) against the reference logits.
"""
import os
import sys

import numpy as np
import pytest
import torch

from _util import GOLDEN_DIR, rel_fro

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit_attention  # noqa: E402
from oracle.model_glue import attention_state, vit_model_forward  # noqa: E402

CFG = dict(image_size=64, patch_size=16, in_channels=3, depths=2, num_heads=2, embedding_dim=128, num_classes=10)
REF = "/root/reference/vision_transformers"


def _load():
    z = np.load(os.path.join(GOLDEN_DIR, "models", "vit_tiny.npz"))
    sd = {k[2:]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith("p.")}
    return sd, torch.from_numpy(z["in.x"].astype(np.float32)), torch.from_numpy(z["y_ref"])


def test_glue_with_oracle_attention_reproduces_reference_model_logits():
    sd, x, y_ref = _load()

    def attn(i, t):
        a = attention_state(sd, i)
        return vit_attention(t, a["qkv.weight"], a.get("qkv.bias"), a["proj.weight"], a["proj.bias"], CFG["num_heads"])

    y = vit_model_forward(sd, x, attn, CFG["patch_size"], CFG["depths"])
    assert (y - y_ref).abs().max().item() <= 2e-5 * y_ref.abs().max().item()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
@pytest.mark.parametrize("level", ["attention", "block"])
def test_class_swap_inside_the_reference_model_constructs_and_loads(monkeypatch, level):
    import pytorch_attention_b200 as pa
    monkeypatch.syspath_prepend(REF)
    import ViT
    if level == "attention":
        monkeypatch.setattr(ViT, "Attention", pa.vit.Attention)    # the substitution INTEGRATION.md describes
    else:
        monkeypatch.setattr(ViT, "TransformerEncoder", pa.vit.TransformerEncoder)   # whole block: LN + attention + residual fused
    model = ViT.VisionTransformer(**CFG).eval()                    # ViT.py:98 constructs the drop-in with the reference's arguments
    assert all(isinstance(b.attn, pa.vit.Attention) for b in model.blocks)
    if level == "block":
        assert all(isinstance(b, pa.vit.TransformerEncoder) for b in model.blocks)
    sd, x, _ = _load()
    model.load_state_dict(sd, strict=True)                         # identical keys and shapes
    with pytest.raises(RuntimeError, match="CPU"):                 # product path: no CPU fallback
        with torch.no_grad():
            model(x)


@pytest.mark.gpu
def test_dropin_attention_inside_the_whole_model_matches_reference_logits():
    import pytorch_attention_b200 as pa
    sd, x, y_ref = _load()
    dev = torch.device("cuda", 0)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    mods = []
    for i in range(CFG["depths"]):
        m = pa.vit.Attention(CFG["embedding_dim"], CFG["num_heads"]).eval()
        m.load_state_dict(attention_state(sd, i), strict=True)
        mods.append(m.to(dev))

    def attn(i, t):
        return mods[i](t.half()).float()                           # LayerNorm output rounded once to the kernel's fp16 input

    with torch.no_grad():
        y = vit_model_forward(sdg, x.to(dev), attn, CFG["patch_size"], CFG["depths"]).cpu()
    # the path's own tolerance is 1e-3 per attention; two blocks, fp16-rounded attention inputs and the MLP/head amplify it
    assert rel_fro(y, y_ref) < 3e-3, rel_fro(y, y_ref)


def _swap_targets():
    import pytorch_attention_b200 as pa
    return {
        "pvt": ("pvt_t", {"Attention": pa.pvt.Attention}),
        "cvt": ("cvt_13", {"Attention": pa.cvt.Attention}),
        "cswin": ("CSWin_64_12211_tiny_224", {"LePEAttention": pa.cswin.LePEAttention, "CSWinBlock": pa.cswin.CSWinBlock}),
        "xcit": ("xcit_nano_12_p16", {"XCA": pa.xcit.XCA, "ClassAttention": pa.xcit.ClassAttention}),
        "moat": ("moat_0", {"Attention": pa.moat.Attention}),      # 48- and 96-wide heads (dims 384 / 768, 8 heads)
    }


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
@pytest.mark.parametrize("modname", ["pvt", "cvt", "cswin", "xcit", "moat"])
def test_class_swap_in_the_reference_zoo_models(modname, monkeypatch):
    """The reference's own model constructors build the drop-ins with the reference's arguments, and a stock model's
    state_dict loads strictly into the swapped model (same keys, same shapes) -- for every model family on the path."""
    import importlib
    monkeypatch.syspath_prepend(REF)
    mod = importlib.import_module(modname)
    factory, swaps = _swap_targets()[modname]
    torch.manual_seed(0)
    stock = getattr(mod, factory)()
    sd = stock.state_dict()
    for name, cls in swaps.items():
        monkeypatch.setattr(mod, name, cls)
    swapped = getattr(mod, factory)()
    n_dropins = sum(isinstance(m, tuple(swaps.values())) for m in swapped.modules())
    n_stock = sum(type(m).__name__ in swaps for m in stock.modules())
    assert n_dropins == n_stock and n_dropins > 0
    swapped.load_state_dict(sd, strict=True)


# ---------------------------------------------------------------- the reference's DEFAULT model: VisionTransformer()
def _load_default():
    from oracle.model_glue import vit_state_shapes, synth_state_dict
    z = np.load(os.path.join(GOLDEN_DIR, "models", "vit_default.npz"))
    seed = int(z["seed"])
    sd = synth_state_dict(vit_state_shapes(), seed)              # 86 M parameters rebuilt from the seed (only logits are stored)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(seed)).half().float()
    return sd, x, torch.from_numpy(z["y_ref"])


def test_default_model_glue_with_oracle_attention_reproduces_reference_logits():
    """`VisionTransformer()` with every constructor default (12 blocks, dim 768, 4 heads of 192: README.md:331-334)."""
    sd, x, y_ref = _load_default()

    def attn(i, t):
        a = attention_state(sd, i)
        return vit_attention(t, a["qkv.weight"], a.get("qkv.bias"), a["proj.weight"], a["proj.bias"], 4)

    with torch.no_grad():
        y = vit_model_forward(sd, x, attn, 16, 12)
    assert (y - y_ref).abs().max().item() <= 5e-5 * y_ref.abs().max().item()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_default_reference_model_accepts_the_dropin_and_the_synthetic_state(monkeypatch):
    import pytorch_attention_b200 as pa
    from oracle.model_glue import vit_state_shapes, synth_state_dict
    monkeypatch.syspath_prepend(REF)
    import ViT
    monkeypatch.setattr(ViT, "Attention", pa.vit.Attention)
    model = ViT.VisionTransformer()                                # every default: 4 heads of 192
    assert all(b.attn.num_heads == 4 for b in model.blocks)
    model.load_state_dict(synth_state_dict(vit_state_shapes(), 77), strict=True)


@pytest.mark.gpu
def test_default_model_with_dropin_attention_matches_reference_logits():
    """The INTEGRATION.md §1 example at full size: the reference's default ViT (192-wide heads -> the panelled attention
    core) with the B200 drop-in in all 12 blocks, against the logits of the real reference model."""
    import pytorch_attention_b200 as pa
    sd, x, y_ref = _load_default()
    dev = torch.device("cuda", 0)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    mods = []
    for i in range(12):
        m = pa.vit.Attention(768).eval()                            # constructor defaults, as ViT.py:111 builds it
        m.load_state_dict(attention_state(sd, i), strict=True)
        m = m.to(dev)
        m.fp32_input = torch.float16                                # fp32 model: opt-in cast, fp32 result
        mods.append(m)
    with torch.no_grad():
        y = vit_model_forward(sdg, x.to(dev), lambda i, t: mods[i](t), 16, 12).cpu()
    # 12 blocks of a 1e-3 path with fp16-rounded attention inputs, amplified by the MLPs and the head
    assert rel_fro(y, y_ref) < 1e-2, rel_fro(y, y_ref)
    assert (y.argmax(-1) == y_ref.argmax(-1)).all()
