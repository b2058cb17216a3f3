"""CPU-only: the algebraic identities the B200 kernels rest on, checked in fp64 against the oracle (which
tests/test_oracle_golden.py pins to the real reference).  A kernel that computes the restated form computes what the
reference computes - the GPU suite then only has to show that the kernel computes the restated form."""
import math

import pytest
import torch

from cases import SUBMODULE_CFG
from conformer_b200.synth import synth_encoder_state_dict
from oracle import conformer_oracle as O

torch.set_grad_enabled(False)
PRE = "layers.0.attention.attention."


def _case(dtype=torch.float64):
    c = SUBMODULE_CFG
    cfg = c["cfg"]
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in synth_encoder_state_dict(cfg, c["seed"]).items()}
    g = torch.Generator().manual_seed(c["x_seed"])
    x = torch.randn(c["B"], c["T"], cfg.d_model, generator=g).to(dtype)
    return cfg, sd, x, torch.tensor(c["lengths"])


def factorised_attention(sd, x, div_term, key_is_pad, n_heads):
    """What rel_attention_fact_kernel computes (attention_fact.cu header), written out with dense tensors:
    pos[i,j] = sum_k A_ik cos(j w_k) + B_ik sin(j w_k) with (a, b) = the sin / cos input columns of Wp_h^T (q_i + v),
    A = a sin(i w) + b cos(i w), B = b sin(i w) - a cos(i w): no relative shift, no [T, 2T-1] tensor, no projected
    table, and the pos_proj bias dropped (a per-query constant under the softmax)."""
    B, T, d = x.shape
    dh = d // n_heads
    lin = lambda name, t: t @ sd[PRE + name + ".weight"].T + sd[PRE + name + ".bias"]     # noqa: E731
    q = lin("query_proj", x).reshape(B, T, n_heads, dh).transpose(1, 2)
    k = lin("key_proj", x).reshape(B, T, n_heads, dh).transpose(1, 2)
    v = lin("value_proj", x).reshape(B, T, n_heads, dh).transpose(1, 2)
    u = sd[PRE + "content_bias"][None, :, None, :]
    vb = sd[PRE + "position_bias"][None, :, None, :]
    wp = sd[PRE + "pos_proj.weight"].reshape(n_heads, dh, d)                                # [H, dh, d]
    g = torch.einsum("bhtc,hcm->bhtm", q + vb, wp)                                           # Wp_h^T (q + v): [B,H,T,d]
    a, b = g[..., 0::2], g[..., 1::2]                                                        # sin / cos input columns
    w = div_term.reshape(-1).to(x.dtype)
    ang = torch.arange(T, dtype=x.dtype)[:, None] * w[None, :]                               # [T, d/2]: i w_k (= j w_k)
    s, c = torch.sin(ang), torch.cos(ang)
    A = a * s + b * c
    Bq = b * s - a * c
    positional = A @ c.T + Bq @ s.T                                                          # rotated query x ABSOLUTE key table
    content = (q + u) @ k.transpose(-1, -2)
    score = (content + positional) / math.sqrt(dh)
    if key_is_pad is not None:
        score = score.masked_fill(key_is_pad[:, None, None, :], float("-inf"))
    ctx = (torch.softmax(score, dim=-1) @ v).transpose(1, 2).reshape(B, T, d)
    return lin("out_proj", ctx)


@pytest.mark.parametrize("masked", [False, True])
def test_sinusoid_factorisation_equals_the_relative_shift(masked):
    cfg, sd, x, lengths = _case()
    lengths = torch.tensor([x.shape[1], 30, 7])                     # every sample keeps usable keys
    pad = ~O.padding_mask(lengths, x.shape[1]) if masked else None
    pos = O.rel_pos_table(x.shape[1], cfg.d_model, sd["rel_pe.div_term"], dtype=torch.float64)
    want = O.rel_attention(sd, PRE, x, pos, pad, cfg.n_heads)
    got = factorised_attention(sd, x, sd["rel_pe.div_term"], pad, cfg.n_heads)
    assert float((got - want).abs().max()) < 1e-10 * (1.0 + float(want.abs().max()))


def test_pos_proj_bias_and_key_proj_bias_cancel_under_the_softmax():
    """both biases add a per-query constant to every score of a row (the training tests rely on it too:
    their gradients are exactly zero)."""
    cfg, sd, x, _ = _case()
    pos = O.rel_pos_table(x.shape[1], cfg.d_model, sd["rel_pe.div_term"], dtype=torch.float64)
    want = O.rel_attention(sd, PRE, x, pos, None, cfg.n_heads)
    sd2 = dict(sd)
    sd2[PRE + "pos_proj.bias"] = sd[PRE + "pos_proj.bias"] + 3.0 * torch.randn_like(sd[PRE + "pos_proj.bias"])
    sd2[PRE + "key_proj.bias"] = sd[PRE + "key_proj.bias"] - 2.0 * torch.randn_like(sd[PRE + "key_proj.bias"])
    got = O.rel_attention(sd2, PRE, x, pos, None, cfg.n_heads)
    assert float((got - want).abs().max()) < 1e-10 * (1.0 + float(want.abs().max()))


def test_split_key_online_softmax_with_lazy_rescaling_is_exact():
    """The attention kernels keep one running maximum / sum / accumulator per HALF of each 64-key tile, move the
    reference maximum only when a tile exceeds it by more than 2^8, and merge the halves once at the end: any
    reference below the true maximum by a bounded amount gives the same normalised result."""
    g = torch.Generator().manual_seed(3)
    T, dh = 200, 16
    s = (4.0 * torch.randn(5, T, generator=g, dtype=torch.float64)) * math.log2(math.e)      # scores in log2 units
    v = torch.randn(T, dh, generator=g, dtype=torch.float64)
    want = torch.softmax(s * math.log(2.0), dim=-1) @ v
    m = [torch.full((5,), -math.inf, dtype=torch.float64) for _ in range(2)]
    acc = [torch.zeros(5, dh, dtype=torch.float64) for _ in range(2)]
    lsum = [torch.zeros(5, dtype=torch.float64) for _ in range(2)]
    for j0 in range(0, T, 64):
        for half in range(2):
            lo, hi = j0 + 32 * half, min(j0 + 32 * half + 32, T)
            if lo >= hi:
                continue
            tmax = s[:, lo:hi].max(dim=-1).values
            m_to = torch.where(tmax > m[half] + 8.0, tmax, m[half])                      # lazy: move only past 2^8
            f = torch.where(torch.isinf(m[half]), torch.zeros_like(m_to), torch.exp2(m[half] - m_to))
            acc[half] = acc[half] * f[:, None]
            lsum[half] = lsum[half] * f
            m[half] = m_to
            p = torch.exp2(s[:, lo:hi] - m[half][:, None])
            assert float(p.max()) <= 2.0 ** 8 + 1e-9                                        # P stays inside bf16's comfortable range
            acc[half] = acc[half] + p @ v[lo:hi]
            lsum[half] = lsum[half] + p.sum(dim=-1)
    m_all = torch.maximum(m[0], m[1])
    f0, f1 = torch.exp2(m[0] - m_all), torch.exp2(m[1] - m_all)
    got = (f0[:, None] * acc[0] + f1[:, None] * acc[1]) / (f0 * lsum[0] + f1 * lsum[1])[:, None]
    assert float((got - want).abs().max()) < 1e-12


def test_swish_through_one_tanh_and_folded_batchnorm():
    """ffn_tc.cu / conv_module_tc.cu: x sigmoid(x) = h + h tanh(h) with h = x / 2 (one MUFU op per element, biases
    pre-halved); BatchNorm(eval) folds into the depthwise taps and bias (convolution.py:27-29)."""
    x = torch.linspace(-12, 12, 4001, dtype=torch.float64)
    h = 0.5 * x
    assert float((h + h * torch.tanh(h) - x * torch.sigmoid(x)).abs().max()) < 1e-14
    g = torch.Generator().manual_seed(9)
    C, K, T = 6, 31, 50
    z = torch.randn(2, C, T, generator=g, dtype=torch.float64)
    w = torch.randn(C, 1, K, generator=g, dtype=torch.float64)
    b = torch.randn(C, generator=g, dtype=torch.float64)
    bw, bb = torch.randn(C, generator=g, dtype=torch.float64), torch.randn(C, generator=g, dtype=torch.float64)
    rm, rv = torch.randn(C, generator=g, dtype=torch.float64), torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    y = torch.nn.functional.conv1d(z, w, b, padding=K // 2, groups=C)
    want = torch.nn.functional.batch_norm(y, rm, rv, bw, bb, training=False, eps=1e-5)
    scale = bw / torch.sqrt(rv + 1e-5)
    got = torch.nn.functional.conv1d(z, w * scale[:, None, None], (b - rm) * scale + bb, padding=K // 2, groups=C)
    assert float((got - want).abs().max()) < 1e-12
