"""Helpers shared by the tests: golden loading and error metrics."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    inputs, params = {}, {}
    for k in z.files:
        a = z[k]
        if k.startswith("in."):
            inputs[k[3:]] = torch.from_numpy(a.astype(np.float32))
        elif k.startswith("p."):
            params[k[2:]] = torch.from_numpy(a if a.dtype.kind in "iu" else a.astype(np.float32))
    return inputs, params, torch.from_numpy(z["y_ref"])


def rel_fro(y, ref):
    y, ref = y.double(), ref.double()
    return ((y - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def rel_max(y, ref):
    y, ref = y.double(), ref.double()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def dropin_class(variant):
    import pytorch_attention_b200 as pa
    return {"vit": pa.vit.Attention, "vit_block": pa.vit.TransformerEncoder, "setr": pa.setr.Attention, "moat": pa.moat.Attention,
            "bvit": pa.bvit.Broad_Attention, "dilateformer": pa.dilateformer.GlobalAttention, "p2t": pa.p2t.PoolingAttention, "kvt": pa.kvt.KNNAttention, "pvt": pa.pvt.Attention, "pvt_block": pa.pvt.Block, "segformer": pa.segformer.Attention, "cmt": pa.cmt.Attention,
            "cvt": pa.cvt.Attention, "lepe": pa.cswin.LePEAttention, "cswin_block": pa.cswin.CSWinBlock, "xca": pa.xcit.XCA,
            "xca_block": pa.xcit.XCABlockAttentionHalf, "pam": pa.dual_attention.PAM, "class_attn": pa.xcit.ClassAttention}[variant]


def build_dropin(spec, params, out_dtype=None):
    """Instantiate the B200 drop-in for a golden/oracle case spec and load the reference state_dict into it."""
    m = dropin_class(spec["variant"])(**spec["ctor"]).eval()
    if spec.get("keep"):
        # block cases store only the attention half's parameters: everything stored must load, nothing stored may be unknown,
        # and what is missing must lie outside the stored prefixes (the block's MLP half)
        res = m.load_state_dict(params, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        assert not [k for k in res.missing_keys if k.startswith(tuple(spec["keep"]))], res.missing_keys
    else:
        m.load_state_dict(params)          # strict: keys and shapes must match the reference's
    if hasattr(m, "out_dtype"):
        m.out_dtype = out_dtype
    return m


def run_dropin(spec, m, x, inputs=None):
    import torch
    with torch.no_grad():
        if spec["variant"] in ("pvt", "segformer"):
            return m(x, *spec["hw"])
        if spec["variant"] == "cmt":
            return m(x, *spec["hw"], inputs["relative_pos"].float().to(x.device))
        if spec["variant"] == "p2t":
            from oracle.cases import p2t_d_convs
            return m(x, *spec["hw"], d_convs=p2t_d_convs(inputs, x.device))
        if spec["variant"] == "pvt_block":
            return m.attention_half(x, *spec["hw"])
        if spec["variant"] == "bvit":
            return m(x)[0]
        if spec["variant"] in ("cswin_block", "vit_block", "xca_block"):
            return m.attention_half(x)
        return m(x)
