"""fp32 CPU restatement of the reference attention forwards (TEST ORACLE).

Every function is functional (weights passed explicitly, names = the
reference ``state_dict`` keys with dots turned into underscores) and written in
einsum / explicit-index form, independent of the reference's
reshape/permute chains, so that an index-path mistake in either side shows up
as a mismatch.  All math is fp32 (or fp64 when ``dtype=torch.float64``) on CPU.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

_F = torch.float32


def _lin(x, w, b=None):
    """y[..., o] = sum_c x[..., c] w[o, c] (+ b[o])  — nn.Linear semantics."""
    y = torch.einsum("...c,oc->...o", x, w)
    return y if b is None else y + b


def _softmax_last(s):
    m = s.amax(dim=-1, keepdim=True)
    e = torch.exp(s - m)
    return e / e.sum(dim=-1, keepdim=True)


def _bn_eval(t, w, b, mean, var, eps):
    """Eval-mode BatchNorm2d on [B,C,H,W]: per-channel affine of running stats."""
    inv = torch.rsqrt(var + eps)
    return (t - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


# --------------------------------------------------------------------------
# ViT  (reference: vision_transformers/ViT.py:67-89; setr.py:50-72 and
# moat.py:62-84 are the same math)
# --------------------------------------------------------------------------
def vit_attention(x, qkv_weight, qkv_bias, proj_weight, proj_bias, num_heads, scale=None):
    """ViT.Attention.forward (ViT.py:79-89).

    x [B,N,C].  Row o of qkv_weight maps to (s,h,d) = (o//C, (o%C)//hd, o%hd)
    (the reshape(B,N,3,H,hd) at ViT.py:81).
    """
    B, N, C = x.shape
    H = num_heads
    hd = C // H
    scale = hd ** -0.5 if scale is None else scale
    w = qkv_weight.reshape(3, H, hd, C)
    qkv = torch.einsum("bnc,shdc->sbhnd", x, w)
    if qkv_bias is not None:
        qkv = qkv + qkv_bias.reshape(3, 1, H, 1, hd)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = torch.einsum("bhnd,bhmd->bhnm", q, k) * scale          # ViT.py:83
    p = _softmax_last(s)                                        # ViT.py:84
    o = torch.einsum("bhnm,bhmd->bnhd", p, v).reshape(B, N, C)  # ViT.py:86
    return _lin(o, proj_weight, proj_bias)                      # ViT.py:87


def kvt_knn_attention(x, qkv_weight, qkv_bias, proj_weight, proj_bias, num_heads, topk, store_dtype=None):
    """kvt.KNNAttention.forward (kvt.py:79-94): ViT's attention in which only the topk largest scaled scores of every row take
    part in the softmax -- the rest are -inf (mask built by torch.topk + scatter, kvt.py:84-87).

    The top-k SELECTION is discontinuous: rounding q / k to 16 bits moves scores by ~5e-4 relative, which swaps the k-th and
    (k+1)-th entry of a few per cent of the rows and changes those rows by ~1/k of their mass (measured on the golden cases: 1.5 %
    of the rows, rel-Frobenius 6e-3 at k = 100 of 197; 4e-2 at k = 7 of 50) -- no 16-bit implementation can meet 1e-3 against the
    fp32 forward.  ``store_dtype=torch.float16`` rounds the qkv projection the way the B200 path stores it, so that both sides
    select on the same scores; tests compare against that, and against the fp32 reference row by row."""
    B, N, C = x.shape
    H = num_heads
    hd = C // H
    qkv = _lin(x, qkv_weight, qkv_bias)
    if store_dtype is not None:
        qkv = qkv.to(store_dtype).to(x.dtype)
    qkv = qkv.reshape(B, N, 3, H, hd)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    s = torch.einsum("bnhd,bmhd->bhnm", q, k) * hd ** -0.5
    kth = torch.sort(s, dim=-1, descending=True).values[..., topk - 1:topk]      # the k-th largest score of every row
    s = torch.where(s >= kth, s, torch.full_like(s, float("-inf")))
    p = _softmax_last(s)
    o = torch.einsum("bhnm,bmhd->bnhd", p, v).reshape(B, N, C)
    return _lin(o, proj_weight, proj_bias)


def bvit_broad_attention(x, to_qkv_weight, to_out_weight, to_out_bias, heads, dim_head):
    """bvit.Broad_Attention.forward (bvit.py:66-76): returns (out, q, k, v) with q, k, v as [B, heads, N, dim_head].
    Inner width heads * dim_head; to_qkv has no bias; without an output projection (to_out = Identity, bvit.py:52) out is the
    concatenated heads."""
    B, N, _ = x.shape
    qkv = _lin(x, to_qkv_weight).reshape(B, N, 3, heads, dim_head)          # chunk(3) + 'b n (h d) -> b h n d'
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = torch.einsum("bhnd,bhmd->bhnm", q, k) * dim_head ** -0.5
    p = _softmax_last(s)
    o = torch.einsum("bhnm,bhmd->bnhd", p, v).reshape(B, N, heads * dim_head)
    out = o if to_out_weight is None else _lin(o, to_out_weight, to_out_bias)
    return out, q, k, v


def vit_block_attention_half(x, ln_weight, ln_bias, qkv_weight, qkv_bias, proj_weight, proj_bias, num_heads, eps=1e-5):
    """First half of ViT.TransformerEncoder.forward (ViT.py:116): x + attn(layernorm1(x)).  LayerNorm over the channel axis,
    biased variance, eps inside the square root (nn.LayerNorm)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    h = (x - mu) / torch.sqrt(var + eps) * ln_weight + ln_bias
    return x + vit_attention(h, qkv_weight, qkv_bias, proj_weight, proj_bias, num_heads)


# --------------------------------------------------------------------------
# PVT spatial-reduction attention (reference: vision_transformers/pvt.py:52-91)
# --------------------------------------------------------------------------
def pvt_attention(x, H_img, W_img, q_weight, q_bias, k_weight, k_bias, v_weight, v_bias,
                  proj_weight, proj_bias, num_heads, sr_ratio=1,
                  sr_0_weight=None, sr_0_bias=None, sr_1_weight=None, sr_1_bias=None,
                  sr_1_running_mean=None, sr_1_running_var=None, bn_eps=1e-5, relative_pos=None):
    """pvt.Attention.forward(x, H, W) (pvt.py:73-91), eval-mode BatchNorm.
    cmt.Attention.forward(x, H, W, relative_pos) (cmt.py:93-111) is the same module with ``relative_pos`` ([heads, N, M],
    broadcast over the batch) added to the scaled scores before the softmax (cmt.py:100).

    K/V tokens: x_[b, i*(W/sr)+j, c] = BN_c( sum_{u,v<sr} w[c,0,u,v] *
    x[b, (sr*i+u)*W + (sr*j+v), c] + bias[c] )   (pvt.py:77-78).
    """
    B, N, C = x.shape
    nh = num_heads
    hd = C // nh
    scale = hd ** -0.5
    q = _lin(x, q_weight, q_bias).reshape(B, N, nh, hd)
    if sr_ratio > 1:
        sr = sr_ratio
        Hs, Ws = H_img // sr, W_img // sr
        # image view of the tokens: xi[b, i, u, j, v, c]
        xi = x.reshape(B, H_img, W_img, C)[:, : Hs * sr, : Ws * sr].reshape(B, Hs, sr, Ws, sr, C)
        t = torch.einsum("biujvc,cuv->bcij", xi, sr_0_weight[:, 0])
        t = t + sr_0_bias[None, :, None, None]
        t = _bn_eval(t, sr_1_weight, sr_1_bias, sr_1_running_mean, sr_1_running_var, bn_eps)
        kv_in = t.reshape(B, C, Hs * Ws).permute(0, 2, 1)   # [B, M, C]
    else:
        kv_in = x
    M = kv_in.shape[1]
    k = _lin(kv_in, k_weight, k_bias).reshape(B, M, nh, hd)
    v = _lin(kv_in, v_weight, v_bias).reshape(B, M, nh, hd)
    s = torch.einsum("bnhd,bmhd->bhnm", q, k) * scale
    if relative_pos is not None:
        s = s + relative_pos                                      # cmt.py:100
    p = _softmax_last(s)
    o = torch.einsum("bhnm,bmhd->bnhd", p, v).reshape(B, N, C)
    return _lin(o, proj_weight, proj_bias)


def _layer_norm_rows(x, w, b, eps=1e-5):
    """nn.LayerNorm over the channel axis: biased variance, eps inside the square root."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def pvt_block_attention_half(x, H_img, W_img, norm1_weight, norm1_bias, eps=1e-5, **attn_kw):
    """First half of pvt.Block.forward (pvt.py:106): x + attn(norm1(x), H, W); attn_kw = pvt_attention's parameters."""
    return x + pvt_attention(_layer_norm_rows(x, norm1_weight, norm1_bias, eps), H_img, W_img, **attn_kw)


# --------------------------------------------------------------------------
# SegFormer efficient self-attention (reference: segformer.py:17-50)
# --------------------------------------------------------------------------
def segformer_attention(x, H_img, W_img, q_weight, q_bias, kv_weight, kv_bias, proj_weight, proj_bias, num_heads,
                        sr_ratio=1, sr_weight=None, sr_bias=None):
    """segformer.Attention.forward(x, H, W) (segformer.py:33-50).

    Reduction (segformer.py:38-39): a dense conv with kernel = stride = sr, no norm behind it:
    x_[b, i*(W/sr)+j, co] = sum_{ci,u,v} w[co,ci,u,v] * x[b, (sr*i+u)*W + (sr*j+v), ci] + bias[co].
    kv = Linear(C, 2C): output row o -> (k|v, head, d) = (o // C, (o % C) // hd, o % hd)   (segformer.py:40 / 43).
    """
    B, N, C = x.shape
    nh = num_heads
    hd = C // nh
    scale = hd ** -0.5
    q = _lin(x, q_weight, q_bias).reshape(B, N, nh, hd)
    if sr_ratio > 1:
        sr = sr_ratio
        Hs, Ws = H_img // sr, W_img // sr
        xi = x.reshape(B, H_img, W_img, C)[:, : Hs * sr, : Ws * sr].reshape(B, Hs, sr, Ws, sr, C)
        t = torch.einsum("biujvc,ocuv->bijo", xi, sr_weight)
        if sr_bias is not None:
            t = t + sr_bias
        kv_in = t.reshape(B, Hs * Ws, C)
    else:
        kv_in = x
    M = kv_in.shape[1]
    kv = _lin(kv_in, kv_weight, kv_bias).reshape(B, M, 2, nh, hd)
    k, v = kv[:, :, 0], kv[:, :, 1]
    s = torch.einsum("bnhd,bmhd->bhnm", q, k) * scale
    p = _softmax_last(s)
    o = torch.einsum("bhnm,bmhd->bnhd", p, v).reshape(B, N, C)
    return _lin(o, proj_weight, proj_bias)


# --------------------------------------------------------------------------
# CvT convolutional-projection attention (reference: cvt.py:48-76)
# --------------------------------------------------------------------------
def _dwconv_same(x, w, b, ks):
    """Depthwise ks x ks conv, stride 1, zero pad (ks-1)//2, on [B,C,H,W]."""
    B, C, H, W = x.shape
    pad = (ks - 1) // 2
    xp = torch.zeros(B, C, H + 2 * pad, W + 2 * pad, dtype=x.dtype)
    xp[:, :, pad:pad + H, pad:pad + W] = x
    out = torch.zeros(B, C, H + 2 * pad - ks + 1, W + 2 * pad - ks + 1, dtype=x.dtype)
    Ho, Wo = out.shape[2], out.shape[3]
    for u in range(ks):
        for v in range(ks):
            out = out + xp[:, :, u:u + Ho, v:v + Wo] * w[None, :, 0, u, v, None, None]
    return out + b[None, :, None, None]


def cvt_attention(x, conv_proj_qkv_0_weight, conv_proj_qkv_0_bias,
                  conv_proj_qkv_1_weight, conv_proj_qkv_1_bias,
                  conv_proj_qkv_1_running_mean, conv_proj_qkv_1_running_var,
                  conv_proj_qkv_2_weight, conv_proj_qkv_2_bias,
                  proj_weight, proj_bias, num_heads, ks=3, bn_eps=1e-5):
    """cvt.Attention.forward (cvt.py:64-76).  x and result are NCHW.

    out-channel o of the 1x1 qkv conv maps to (s,h,d)=(o//C,(o%C)//hd,o%hd)
    (reshape(B,3,H,hd,Hh,Ww) at cvt.py:66); tokens n = r*W + c.
    """
    B, C, Hh, Ww = x.shape
    nh = num_heads
    hd = C // nh
    scale = hd ** -0.5
    t = _dwconv_same(x, conv_proj_qkv_0_weight, conv_proj_qkv_0_bias, ks)
    t = _bn_eval(t, conv_proj_qkv_1_weight, conv_proj_qkv_1_bias,
                 conv_proj_qkv_1_running_mean, conv_proj_qkv_1_running_var, bn_eps)
    t = t.reshape(B, C, Hh * Ww)
    w = conv_proj_qkv_2_weight.reshape(3, nh, hd, C)
    qkv = torch.einsum("bcn,shdc->sbhnd", t, w) + conv_proj_qkv_2_bias.reshape(3, 1, nh, 1, hd)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = torch.einsum("bhnd,bhmd->bhnm", q, k) * scale
    p = _softmax_last(s)
    o = torch.einsum("bhnm,bhmd->bhdn", p, v).reshape(B, C, Hh * Ww)   # channel = h*hd+d
    y = torch.einsum("bcn,oc->bon", o, proj_weight.reshape(C, C)) + proj_bias[None, :, None]
    return y.reshape(B, C, Hh, Ww)


# --------------------------------------------------------------------------
# CSWin (reference: cswin.py:51-127, 130-197, 199-216)
# --------------------------------------------------------------------------
def _cswin_window_shape(resolution, idx, split_size):
    if idx == -1:
        return resolution, resolution
    if idx == 0:
        return resolution, split_size
    if idx == 1:
        return split_size, resolution
    raise ValueError(f"ERROR MODE {idx}")   # reference prints and exit(0)s, cswin.py:68-70


def cswin_window_table(resolution, idx, split_size):
    """Integer index path of img2windows/windows2img (cswin.py:199-216).

    Returns int64 array T[nwin, H_sp*W_sp]: image-token index (r*W+c) of token t
    of window w, with window id i*(W/W_sp)+j and in-window index r*W_sp+c.
    """
    H = W = resolution
    H_sp, W_sp = _cswin_window_shape(resolution, idx, split_size)
    nI, nJ = H // H_sp, W // W_sp
    T = np.empty((nI * nJ, H_sp * W_sp), dtype=np.int64)
    for i in range(nI):
        for j in range(nJ):
            for r in range(H_sp):
                for c in range(W_sp):
                    T[i * nJ + j, r * W_sp + c] = (i * H_sp + r) * W + (j * W_sp + c)
    return T


def cswin_lepe_attention(qkv, get_v_weight, get_v_bias, resolution, idx, split_size,
                         num_heads, scale=None):
    """LePEAttention.forward (cswin.py:101-127).  qkv [3,B,L,C'] -> [B,L,C']."""
    q, k, v = qkv[0], qkv[1], qkv[2]
    B, L, C = q.shape
    H = W = resolution
    assert L == H * W, "flatten img_tokens has wrong size"
    H_sp, W_sp = _cswin_window_shape(resolution, idx, split_size)
    nh = num_heads
    hd = C // nh
    scale = hd ** -0.5 if scale is None else scale
    T = torch.from_numpy(cswin_window_table(resolution, idx, split_size))   # [nw, Nw]
    nw, Nw = T.shape
    qw = q[:, T].reshape(B, nw, Nw, nh, hd)
    kw = k[:, T].reshape(B, nw, Nw, nh, hd)
    vw = v[:, T].reshape(B, nw, Nw, nh, hd)
    s = torch.einsum("bwnhd,bwmhd->bwhnm", qw * scale, kw)          # cswin.py:116-117
    p = _softmax_last(s)
    o = torch.einsum("bwhnm,bwmhd->bwnhd", p, vw)
    # LePE: depthwise 3x3 over each window, zero padded at window borders (cswin.py:93-96)
    vimg = vw.reshape(B * nw, H_sp, W_sp, C).permute(0, 3, 1, 2)
    lepe = _dwconv_same(vimg, get_v_weight, get_v_bias, 3)          # [B*nw, C, H_sp, W_sp]
    lepe = lepe.permute(0, 2, 3, 1).reshape(B, nw, Nw, nh, hd)
    o = (o + lepe).reshape(B, nw * Nw, C)
    out = torch.empty(B, L, C, dtype=o.dtype)
    out[:, T.reshape(-1)] = o                                        # windows2img, cswin.py:208-216
    return out


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def cswin_block_attention(x, norm1_weight, norm1_bias, qkv_weight, qkv_bias,
                          proj_weight, proj_bias, get_v_weights, get_v_biases,
                          reso, num_heads, split_size=7, last_stage=False,
                          qk_scale=None, residual=True):
    """Attention half of CSWinBlock.forward (cswin.py:176-194).

    get_v_weights/biases: list over branches (attns.{i}.get_v.*).  Returns
    x + proj(cat(branches)) if ``residual`` else proj(cat(branches)).
    """
    B, L, C = x.shape
    assert L == reso * reso, "flatten img_tokens has wrong size"
    if reso == split_size:
        last_stage = True
    img = _layer_norm(x, norm1_weight, norm1_bias)
    qkv = _lin(img, qkv_weight, qkv_bias).reshape(B, L, 3, C).permute(2, 0, 1, 3)   # o -> (s, c)
    if last_stage:
        att = cswin_lepe_attention(qkv, get_v_weights[0], get_v_biases[0], reso, -1, split_size,
                                   num_heads, qk_scale)
    else:
        half = C // 2
        a0 = cswin_lepe_attention(qkv[..., :half], get_v_weights[0], get_v_biases[0], reso, 0,
                                  split_size, num_heads // 2, qk_scale)
        a1 = cswin_lepe_attention(qkv[..., half:], get_v_weights[1], get_v_biases[1], reso, 1,
                                  split_size, num_heads // 2, qk_scale)
        att = torch.cat([a0, a1], dim=2)
    y = _lin(att, proj_weight, proj_bias)
    return x + y if residual else y


# --------------------------------------------------------------------------
# XCiT (reference: xcit.py:159-188, 233-265)
# --------------------------------------------------------------------------
def xca_attention(x, qkv_weight, qkv_bias, proj_weight, proj_bias, temperature, num_heads):
    """XCA.forward (xcit.py:245-265): channel x channel attention."""
    B, N, C = x.shape
    H = num_heads
    hd = C // H
    w = qkv_weight.reshape(3, H, hd, C)
    qkv = torch.einsum("bnc,shdc->sbhdn", x, w)       # already [.., hd, N] (xcit.py:251-253)
    if qkv_bias is not None:
        qkv = qkv + qkv_bias.reshape(3, 1, H, hd, 1)
    q, k, v = qkv[0], qkv[1], qkv[2]
    qn = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)      # F.normalize, xcit.py:255
    kn = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    a = torch.einsum("bhdn,bhen->bhde", qn, kn) * temperature.reshape(1, H, 1, 1)
    a = _softmax_last(a)
    o = torch.einsum("bhde,bhen->bnhd", a, v).reshape(B, N, C)  # xcit.py:262
    return _lin(o, proj_weight, proj_bias)


def class_attention(x, qkv_weight, qkv_bias, proj_weight, proj_bias, num_heads, scale=None):
    """ClassAttention.forward (xcit.py:174-188): CLS query only; patch tokens pass through."""
    B, N, C = x.shape
    H = num_heads
    hd = C // H
    scale = hd ** -0.5 if scale is None else scale
    w = qkv_weight.reshape(3, H, hd, C)
    qkv = torch.einsum("bnc,shdc->sbhnd", x, w)
    if qkv_bias is not None:
        qkv = qkv + qkv_bias.reshape(3, 1, H, 1, hd)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = torch.einsum("bhd,bhnd->bhn", q[:, :, 0], k) * scale     # xcit.py:180-181
    a = _softmax_last(a)
    cls = torch.einsum("bhn,bhnd->bhd", a, v).reshape(B, 1, C)   # xcit.py:185
    cls = _lin(cls, proj_weight, proj_bias)
    return torch.cat([cls, x[:, 1:]], dim=1)                     # xcit.py:187


def xca_block_attention_half(x, norm1_weight, norm1_bias, gamma1, qkv_weight, qkv_bias, proj_weight, proj_bias, temperature,
                             num_heads, eps=1e-6):
    """First line of XCABlock.forward (xcit.py:291): x + gamma1 * attn(norm1(x)).  eps: XCiT builds its blocks with
    norm_layer = partial(nn.LayerNorm, eps=1e-6); a bare XCABlock uses nn.LayerNorm's 1e-5 -- the caller passes the module's."""
    h = _layer_norm_rows(x, norm1_weight, norm1_bias, eps)
    return x + gamma1 * xca_attention(h, qkv_weight, qkv_bias, proj_weight, proj_bias, temperature, num_heads)


# --------------------------------------------------------------------------
# DANet position attention module (reference: attention_mechanisms/dual_attention.py:12-28)
# --------------------------------------------------------------------------
def pam_attention(x, b_weight, b_bias, c_weight, c_bias, d_weight, d_bias, alpha):
    """PAM.forward (dual_attention.py:21-28): NCHW in and out, one head as wide as the channel count, no score scale:
    attn[i, j] = softmax_j( b(x)[:, i] . c(x)[:, j] ),  y[:, i] = sum_j attn[i, j] d(x)[:, j],  out = alpha * y + x."""
    n, c, h, w = x.shape
    t = x.reshape(n, c, h * w)
    def conv(wt, bs):
        return torch.einsum("oc,bcn->bon", wt.reshape(c, c), t) + bs[None, :, None]
    B_, C_, D_ = conv(b_weight, b_bias), conv(c_weight, c_bias), conv(d_weight, d_bias)
    p = _softmax_last(torch.einsum("bci,bcj->bij", B_, C_))
    y = torch.einsum("bij,bcj->bci", p, D_).reshape(n, c, h, w)
    return alpha * y + x


# --------------------------------------------------------------------------
# P2T pooling attention (reference: vision_transformers/p2t.py:46-94)
# --------------------------------------------------------------------------
def _adaptive_avg_pool2d(t, oh, ow):
    """F.adaptive_avg_pool2d on [B,C,H,W]: output (i, j) averages rows [floor(i*H/oh), ceil((i+1)*H/oh)) x the same in W."""
    B, C, H, W = t.shape
    out = t.new_zeros(B, C, oh, ow)
    for i in range(oh):
        h0, h1 = (i * H) // oh, -((-(i + 1) * H) // oh)
        for j in range(ow):
            w0, w1 = (j * W) // ow, -((-(j + 1) * W) // ow)
            out[:, :, i, j] = t[:, :, h0:h1, w0:w1].mean(dim=(2, 3))
    return out


def p2t_pooling_attention(x, H_img, W_img, q_weight, q_bias, kv_weight, kv_bias, proj_weight, proj_bias, norm_weight, norm_bias,
                          num_heads, pool_ratios, dconv_weights, dconv_biases, scale=None, eps=1e-5):
    """PoolingAttention.forward(x, H, W, d_convs) (p2t.py:74-94).  dconv_weights[i]: [C,1,3,3] depthwise, zero pad 1."""
    B, N, C = x.shape
    nh = num_heads
    hd = C // nh
    scale = scale or hd ** -0.5
    q = _lin(x, q_weight, q_bias).reshape(B, N, nh, hd)                     # p2t.py:76
    x_ = x.permute(0, 2, 1).reshape(B, C, H_img, W_img)
    pools = []
    for r, w, b in zip(pool_ratios, dconv_weights, dconv_biases):
        pool = _adaptive_avg_pool2d(x_, round(H_img / r), round(W_img / r))  # p2t.py:80
        pool = pool + _dwconv_same(pool, w, b, 3)                             # p2t.py:81
        pools.append(pool.reshape(B, C, -1))
    pools = torch.cat(pools, dim=2).permute(0, 2, 1)                         # [B, M, C]  p2t.py:84-85
    mu = pools.mean(dim=-1, keepdim=True)
    var = ((pools - mu) ** 2).mean(dim=-1, keepdim=True)
    pools = (pools - mu) / torch.sqrt(var + eps) * norm_weight + norm_bias
    M = pools.shape[1]
    kv = _lin(pools, kv_weight, kv_bias).reshape(B, M, 2, nh, hd)            # p2t.py:87
    k, v = kv[:, :, 0], kv[:, :, 1]
    s = torch.einsum("bnhd,bmhd->bhnm", q, k) * scale
    p = _softmax_last(s)
    o = torch.einsum("bhnm,bmhd->bnhd", p, v).reshape(B, N, C)
    return _lin(o, proj_weight, proj_bias)
