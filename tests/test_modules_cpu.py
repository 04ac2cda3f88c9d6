"""Host-side contract of the drop-in modules (no GPU): constructor signatures, state_dict keys/shapes,
error behaviour, and no silent CPU fallback."""
import inspect
import os
import sys

import pytest
import torch

import pytorch_attention_b200 as pa

REF = os.environ.get("PA_REFERENCE", "/root/reference/vision_transformers")
HAVE_REF = os.path.isdir(REF)


def _ref(mod):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return __import__(mod)


def test_vit_state_dict_keys_and_shapes():
    m = pa.ViTAttention(768, 12)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "qkv.weight": (2304, 768), "proj.weight": (768, 768), "proj.bias": (768,)}
    m = pa.ViTAttention(128, 2, qkv_bias=True)
    assert "qkv.bias" in m.state_dict()
    assert m.scale == 64 ** -0.5


def test_vit_constructor_asserts_like_reference():
    with pytest.raises(AssertionError):       # ViT.py:70
        pa.ViTAttention(100, 3)


def test_cpu_tensor_is_an_error_not_a_fallback():
    m = pa.ViTAttention(128, 2).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 128).half())


def test_unsupported_modes_raise():
    from pytorch_attention_b200._common import check_forward_mode

    class Fake:  # a CUDA-looking tensor stand-in is not constructible here; exercise the dtype / mode checks
        is_cuda, dtype, requires_grad = True, torch.float32, False
    m = pa.ViTAttention(128, 2, attn_drop=0.1)
    with pytest.raises(ValueError):
        check_forward_mode(m.eval(), Fake(), (0.1,))
    Fake.dtype = torch.float16
    with pytest.raises(NotImplementedError):
        check_forward_mode(m.train(), Fake(), (0.1,))
    check_forward_mode(m.eval(), Fake(), (0.1,))


@pytest.mark.skipif(not HAVE_REF, reason="reference not mounted")
def test_vit_matches_live_reference_contract():
    ref = _ref("ViT").Attention
    assert str(inspect.signature(ref.__init__)) == str(inspect.signature(pa.ViTAttention.__init__))
    r = ref(192, 3, qkv_bias=True)
    m = pa.ViTAttention(192, 3, qkv_bias=True)
    assert {k: v.shape for k, v in r.state_dict().items()} == {k: v.shape for k, v in m.state_dict().items()}
    m.load_state_dict(r.state_dict())          # reference weights load unchanged
