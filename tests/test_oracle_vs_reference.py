"""Whenever the live reference is mounted (build container), re-check the oracle and the
index paths against it directly, on fresh seeds and on arange tensors."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_CASES, build_reference_case, run_oracle_case, load_reference
from oracle import cswin_window_table

REF = os.environ.get("PA_REFERENCE", "/root/reference/vision_transformers")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_vs_live_reference(name, seed):
    spec = GOLDEN_CASES[name]
    case = build_reference_case(spec, REF, seed=seed)
    y = run_oracle_case(spec, case["inputs"], case["params"])
    tol = 5e-6 * max(1.0, case["y_ref"].abs().max().item())
    assert (y - case["y_ref"]).abs().max().item() <= tol


@pytest.mark.parametrize("reso,idx,split", [(14, 0, 7), (14, 1, 7), (7, -1, 7), (28, 0, 2)])
def test_window_table_vs_reference_img2windows(reso, idx, split):
    """Push an arange image through the reference's img2windows: bit-exact index table."""
    cswin = load_reference(REF)["cswin"]
    H_sp, W_sp = {(-1): (reso, reso), 0: (reso, split), 1: (split, reso)}[idx]
    img = torch.arange(reso * reso, dtype=torch.float32).reshape(1, 1, reso, reso)
    win = cswin.img2windows(img, H_sp, W_sp)[..., 0].long().numpy()
    assert np.array_equal(win, cswin_window_table(reso, idx, split))
    back = cswin.windows2img(torch.from_numpy(win).float().reshape(-1, H_sp, W_sp, 1), H_sp, W_sp, reso, reso)
    assert np.array_equal(back.reshape(-1).long().numpy(), np.arange(reso * reso))


EXTRA_SPECS = {
    # configurations of the GPU oracle cases that have no committed golden file: the restatement is checked against
    # the live reference here, so those GPU cases are pinned through it
    "xca_nano_hd32": dict(variant="xca", ctor=dict(dim=128, num_heads=4), x=(3, 196, 128)),
    "classattn_nano_hd32": dict(variant="class_attn", ctor=dict(dim=128, num_heads=4), x=(3, 197, 128)),
    "vit_l_like": dict(variant="vit", ctor=dict(dim=256, num_heads=4), x=(2, 197, 256)),
}


@pytest.mark.parametrize("name", sorted(EXTRA_SPECS))
def test_oracle_vs_live_reference_extra_configs(name):
    spec = EXTRA_SPECS[name]
    case = build_reference_case(spec, REF, seed=7)
    y = run_oracle_case(spec, case["inputs"], case["params"])
    tol = 5e-6 * max(1.0, case["y_ref"].abs().max().item())
    assert (y - case["y_ref"]).abs().max().item() <= tol


@pytest.mark.parametrize("ctor", [dict(dim=192, heads=3, dim_head=64), dict(dim=64, heads=1, dim_head=64)])
def test_bvit_q_k_v_and_identity_projection_vs_live_reference(ctor):
    """Broad_Attention returns (out, q, k, v) (bvit.py:76); with heads == 1 and dim_head == dim there is no output projection
    (bvit.py:52, 61-64).  All four outputs of the restatement against the live module."""
    from oracle import attention as A
    from oracle.cases import round_fp16_
    ref_cls = load_reference(REF)["bvit"].Broad_Attention
    torch.manual_seed(3)
    mod = ref_cls(**ctor).eval()
    with torch.no_grad():
        for p in mod.parameters():
            round_fp16_(p)
        x = round_fp16_(torch.randn(2, 37, ctor["dim"]))
        ref = mod(x)
    sd = mod.state_dict()
    got = A.bvit_broad_attention(x, sd["to_qkv.weight"], sd.get("to_out.0.weight"), sd.get("to_out.0.bias"), ctor["heads"], ctor["dim_head"])
    for r, g in zip(ref, got):
        assert r.shape == g.shape
        assert (r - g).abs().max().item() <= 5e-6 * max(1.0, r.abs().max().item())
