"""CPU oracle for the Conformer encoder forward path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functional torch code over a ``state_dict``, the arithmetic
that Alan-404/Conformer performs on its encoder hot path.  It exists so that the CUDA
kernels in ``conformer_b200/csrc`` can be checked on any box (the GPU box has no copy of
the reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` leg may import it; nothing under ``conformer_b200/``
does, and the product path has no CPU fallback.

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden.py``
imports the real modules from /root/reference, runs them on seeded weights/inputs and
commits the outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function below against those files (the reference ships no tests or golden vectors
of its own - SURVEY.md section 4).

Every function names the reference lines it follows.  The code is a restatement, not a
copy: it is functional (no nn.Module), takes weights by their reference ``state_dict``
key, uses closed forms where the reference uses reshape tricks (relative shift,
positional table), and runs in whatever dtype/device the caller's tensors have
(float64 on CPU gives the "truth" for tolerance calibration).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

LN_EPS = 1e-5   # nn.LayerNorm default, used at ffn.py:8, attention.py:10, convolution.py:12, block.py:15
BN_EPS = 1e-5   # nn.BatchNorm1d default, convolution.py:16


def _p(sd: StateDict, key: str, like: torch.Tensor) -> torch.Tensor:
    return sd[key].to(device=like.device, dtype=like.dtype)


# --------------------------------------------------------------------------- activations
def swish(x: torch.Tensor) -> torch.Tensor:
    """model/utils/activation.py:4-8 -- x * sigmoid(x)."""
    return x * torch.sigmoid(x)


def glu(x: torch.Tensor, dim: int) -> torch.Tensor:
    """model/utils/activation.py:10-17 -- first half times sigmoid(second half)."""
    n = x.shape[dim] // 2
    a = x.narrow(dim, 0, n)
    g = x.narrow(dim, n, n)
    return a * torch.sigmoid(g)


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


# --------------------------------------------------------------------------- masking / lengths
def subsampled_lengths(lengths: torch.Tensor) -> torch.Tensor:
    """model/utils/convolution.py:54-55 -- ((L-1)//2-1)//2."""
    return torch.div(torch.div(lengths - 1, 2, rounding_mode="floor") - 1, 2, rounding_mode="floor")


def padding_mask(lengths: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    """model/utils/masking.py:4-13 -- bool [B, max_len], True = valid frame."""
    if max_length is None:
        max_length = int(lengths.max())
    pos = torch.arange(max_length, device=lengths.device, dtype=lengths.dtype)
    return pos[None, :] < lengths[:, None]


# --------------------------------------------------------------------------- positional table
def rel_pos_table(T: int, d_model: int, div_term: Optional[torch.Tensor] = None,
                  dtype=torch.float32, device="cpu") -> torch.Tensor:
    """model/utils/position.py:11-28 -- table [2T-1, d]; row p encodes relative position
    r = T-1-p; even columns sin(r*w_k), odd columns cos(r*w_k).  (The reference builds a
    flipped positive half and a negative half and concatenates; this is the closed form.
    It also repeats the table over the batch, position.py:26, which is not restated here.)
    """
    if div_term is None:
        div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32)
                             * -(math.log(10000.0) / d_model))
    w = div_term.reshape(-1).to(device=device, dtype=dtype)
    r = torch.arange(T - 1, -T, -1, device=device, dtype=dtype)          # T-1 ... -(T-1)
    ang = r[:, None] * w[None, :]
    pe = torch.empty((2 * T - 1, d_model), device=device, dtype=dtype)
    pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe


def relative_shift(pos_score: torch.Tensor) -> torch.Tensor:
    """model/utils/attention.py:94-102 -- net effect out[..., i, j] = pos[..., i, T-1+j-i]
    for j in [0, T), with pos of width 2T-1 (closed form of the pad/view/slice trick)."""
    T = pos_score.shape[-2]
    i = torch.arange(T, device=pos_score.device)
    idx = (T - 1) + i[None, :] - i[:, None]                              # [T, T]
    return torch.gather(pos_score, -1, idx.expand(*pos_score.shape[:-2], T, T))


# --------------------------------------------------------------------------- sub-modules
def feed_forward(sd: StateDict, pre: str, x: torch.Tensor) -> torch.Tensor:
    """model/utils/ffn.py:15-23 (dropouts are identity in eval)."""
    y = layer_norm(x, _p(sd, pre + "layer_norm.weight", x), _p(sd, pre + "layer_norm.bias", x))
    y = y @ _p(sd, pre + "hidden_linear.weight", x).T + _p(sd, pre + "hidden_linear.bias", x)
    y = swish(y)
    return y @ _p(sd, pre + "out_linear.weight", x).T + _p(sd, pre + "out_linear.bias", x)


def rel_attention(sd: StateDict, pre: str, x: torch.Tensor, pos: torch.Tensor,
                  key_is_pad: Optional[torch.Tensor], n_heads: int) -> torch.Tensor:
    """RelativeMultiHeadAttention.forward + scaled_dot_product_relative_attention,
    model/utils/attention.py:47-92, called with q=k=v=x (attention.py:16).

    x [B,T,d]; pos [2T-1,d] (or [B,2T-1,d]); key_is_pad bool [B,T], True = masked key
    (the reference passes it as [B,1,1,T], encoder.py:30)."""
    B, T, d = x.shape
    dh = d // n_heads
    if pos.dim() == 2:
        pos = pos[None].expand(B, -1, -1)
    lin = lambda name, t: t @ _p(sd, pre + name + ".weight", x).T + _p(sd, pre + name + ".bias", x)
    q = lin("query_proj", x).reshape(B, T, n_heads, dh).transpose(1, 2)      # [B,H,T,dh]
    k = lin("key_proj", x).reshape(B, T, n_heads, dh).transpose(1, 2)
    v = lin("value_proj", x).reshape(B, T, n_heads, dh).transpose(1, 2)
    p = lin("pos_proj", pos.to(x.dtype)).reshape(B, -1, n_heads, dh).transpose(1, 2)   # [B,H,2T-1,dh]
    u = _p(sd, pre + "content_bias", x)[None, :, None, :]                    # [1,H,1,dh]
    w = _p(sd, pre + "position_bias", x)[None, :, None, :]
    content = (q + u) @ k.transpose(-1, -2)                                  # attention.py:49
    positional = relative_shift((q + w) @ p.transpose(-1, -2))               # attention.py:52-53
    score = (content + positional) / math.sqrt(dh)                           # attention.py:56
    if key_is_pad is not None:                                               # attention.py:59-62
        score = score.masked_fill(key_is_pad[:, None, None, :], torch.finfo(score.dtype).min)
    attn = torch.softmax(score, dim=-1)                                      # attention.py:65
    ctx = (attn @ v).transpose(1, 2).reshape(B, T, d)                        # attention.py:70,87-88
    return lin("out_proj", ctx)                                              # attention.py:90


def mhsa_module(sd: StateDict, pre: str, x: torch.Tensor, pos: torch.Tensor,
                key_is_pad: Optional[torch.Tensor], n_heads: int) -> torch.Tensor:
    """MultiHeadSelfAttentionModule.forward, model/utils/attention.py:14-18."""
    y = layer_norm(x, _p(sd, pre + "layer_norm.weight", x), _p(sd, pre + "layer_norm.bias", x))
    return rel_attention(sd, pre + "attention.", y, pos, key_is_pad, n_heads)


def conv_module(sd: StateDict, pre: str, x: torch.Tensor) -> torch.Tensor:
    """ConvolutionModule.forward in eval mode, model/utils/convolution.py:21-32.
    Written channels-last (the reference transposes to [B,C,T] and back, :23,:31).
    The depthwise conv zero-pads at the tensor edges only; it never sees lengths."""
    B, T, d = x.shape
    y = layer_norm(x, _p(sd, pre + "layer_norm.weight", x), _p(sd, pre + "layer_norm.bias", x))
    w1 = _p(sd, pre + "pointwise_conv_1.weight", x).reshape(2 * d, d)
    y = glu(y @ w1.T + _p(sd, pre + "pointwise_conv_1.bias", x), dim=-1)     # :24-25
    wd = _p(sd, pre + "deepwise_conv.weight", x).reshape(d, -1)              # [d, k]
    ksz = wd.shape[1]
    half = (ksz - 1) // 2
    ypad = F.pad(y, (0, 0, half, half))                                      # zero pad along T
    acc = torch.zeros_like(y)
    for j in range(ksz):                                                     # :26 cross-correlation
        acc = acc + ypad[:, j:j + T, :] * wd[:, j]
    acc = acc + _p(sd, pre + "deepwise_conv.bias", x)
    rm = _p(sd, pre + "batch_norm.running_mean", x)
    rv = _p(sd, pre + "batch_norm.running_var", x)
    acc = (acc - rm) * torch.rsqrt(rv + BN_EPS) * _p(sd, pre + "batch_norm.weight", x) \
        + _p(sd, pre + "batch_norm.bias", x)                                 # :27 (eval)
    acc = swish(acc)                                                         # :28
    w2 = _p(sd, pre + "pointwise_conv_2.weight", x).reshape(d, d)
    return acc @ w2.T + _p(sd, pre + "pointwise_conv_2.bias", x)             # :29


def conv_subsampling(sd: StateDict, pre: str, x: torch.Tensor,
                     lengths: Optional[torch.Tensor]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """ConvolutionSubsampling.forward, model/utils/convolution.py:42-57.
    x [B, n_mel, Tin] -> [B, T', d*F2] with feature index c*F2 + f (channel-major)."""
    y = x[:, None]                                                           # [B,1,mel,Tin]
    y = F.relu(F.conv2d(y, _p(sd, pre + "conv_1.weight", x), _p(sd, pre + "conv_1.bias", x), stride=2))
    y = F.relu(F.conv2d(y, _p(sd, pre + "conv_2.weight", x), _p(sd, pre + "conv_2.bias", x), stride=2))
    B, C, F2, T2 = y.shape
    y = y.permute(0, 3, 1, 2).reshape(B, T2, C * F2)
    if lengths is not None:
        lengths = subsampled_lengths(lengths)
    return y, lengths


def conformer_block(sd: StateDict, pre: str, x: torch.Tensor, pos: torch.Tensor,
                    key_is_pad: Optional[torch.Tensor], n_heads: int) -> torch.Tensor:
    """ConformerBlock.forward, model/utils/block.py:17-29."""
    x = x + 0.5 * feed_forward(sd, pre + "ffn_1.", x)
    x = x + mhsa_module(sd, pre + "attention.", x, pos, key_is_pad, n_heads)
    x = x + conv_module(sd, pre + "conv.", x)
    x = x + 0.5 * feed_forward(sd, pre + "ffn_2.", x)
    return layer_norm(x, _p(sd, pre + "layer_norm.weight", x), _p(sd, pre + "layer_norm.bias", x))


def count_blocks(sd: StateDict, pre: str = "") -> int:
    n = 0
    while f"{pre}layers.{n}.layer_norm.weight" in sd:
        n += 1
    return n


def encoder(sd: StateDict, x: torch.Tensor, lengths: Optional[torch.Tensor], n_heads: int,
            pre: str = "", return_intermediates: bool = False):
    """Encoder.forward, model/modules/encoder.py:18-37."""
    inter: List[torch.Tensor] = []
    y, out_len = conv_subsampling(sd, pre + "downsampling_conv.", x, lengths)
    y = y @ _p(sd, pre + "linear.weight", x).T + _p(sd, pre + "linear.bias", x)   # encoder.py:23
    B, T, d = y.shape
    key_is_pad = None
    if out_len is not None:
        key_is_pad = ~padding_mask(out_len, T)                               # encoder.py:28-30
    pos = rel_pos_table(T, d, sd.get(pre + "rel_pe.div_term"), dtype=y.dtype, device=y.device)
    inter.append(y)
    for i in range(count_blocks(sd, pre)):                                   # encoder.py:34-35
        y = conformer_block(sd, f"{pre}layers.{i}.", y, pos, key_is_pad, n_heads)
        inter.append(y)
    if return_intermediates:
        return y, out_len, inter
    return y, out_len


# --------------------------------------------------------------------------- CTC head (harness only)
def decoder_head(sd: StateDict, x: torch.Tensor, lengths: Optional[torch.Tensor],
                 pre: str = "decoder.") -> torch.Tensor:
    """Decoder.forward in eval mode, model/modules/decoder.py:15-27.  Out of the hot path;
    restated only so that greedy CTC tokens can be compared.  Single-layer LSTM written as
    an explicit time loop (gate order i,f,g,o as in nn.LSTM); frames at or beyond a
    sample's length produce zeros like pad_packed_sequence does (decoder.py:16-21)."""
    B, T, d = x.shape
    if lengths is not None:
        T = int(lengths.max())
    seq = x[:, :T]
    layer = 0
    while pre + f"lstm.weight_ih_l{layer}" in sd:                            # nn.LSTM(num_layers=n), decoder.py:10
        w_ih = _p(sd, pre + f"lstm.weight_ih_l{layer}", x)
        w_hh = _p(sd, pre + f"lstm.weight_hh_l{layer}", x)
        b = _p(sd, pre + f"lstm.bias_ih_l{layer}", x) + _p(sd, pre + f"lstm.bias_hh_l{layer}", x)
        Hd = w_hh.shape[1]
        h = x.new_zeros(B, Hd)
        c = x.new_zeros(B, Hd)
        xin = seq @ w_ih.T + b
        outs = []
        for t in range(T):
            gates = xin[:, t] + h @ w_hh.T
            i, f, g, o = gates.chunk(4, dim=-1)
            c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h_new = torch.sigmoid(o) * torch.tanh(c_new)
            if lengths is not None:
                live = (lengths.to(x.device) > t)[:, None]
                c = torch.where(live, c_new, c)
                h = torch.where(live, h_new, h)
                outs.append(torch.where(live, h_new, torch.zeros_like(h_new)))
            else:
                c, h = c_new, h_new
                outs.append(h_new)
        seq = torch.stack(outs, dim=1)
        layer += 1
    outs = None
    y = swish(seq)                                                           # decoder.py:22
    rm = _p(sd, pre + "norm.running_mean", x)
    rv = _p(sd, pre + "norm.running_var", x)
    y = (y - rm) * torch.rsqrt(rv + BN_EPS) * _p(sd, pre + "norm.weight", x) + _p(sd, pre + "norm.bias", x)
    return y @ _p(sd, pre + "linear.weight", x).T + _p(sd, pre + "linear.bias", x)


def greedy_ctc_ids(logits_or_ids: torch.Tensor, pad_id: int, unk_id: int) -> List[int]:
    """processing/processor.py:301-319 -- collapsed token ids for ONE utterance.
    argmax, skip pad(blank) and unk WITHOUT resetting prev, emit when id != prev.
    (So "a blank a" collapses to a single "a" - non-standard CTC, kept on purpose.)"""
    ids = logits_or_ids.argmax(dim=-1) if logits_or_ids.dim() == 2 else logits_or_ids
    out: List[int] = []
    prev = None
    for t in ids.tolist():
        if t == pad_id or t == unk_id:
            continue
        if prev is None or t != prev:
            prev = t
            out.append(t)
    return out


def greedy_ctc_batch(logits: torch.Tensor, pad_id: int, unk_id: int) -> List[List[int]]:
    """processing/processor.py:324-328 -- every row of the batch, full padded length."""
    return [greedy_ctc_ids(row, pad_id, unk_id) for row in logits]


# --------------------------------------------------------------------------- work counting
def encoder_flops(B: int, Tin: int, d: int, n_blocks: int, n_mel: int = 80, ksz: int = 31) -> float:
    """Algorithmic FLOPs of one encoder forward (SURVEY.md section 8d / BASELINE.md section 3)."""
    T1 = (Tin - 1) // 2
    T2 = (T1 - 1) // 2
    F1 = (n_mel - 1) // 2
    F2 = (F1 - 1) // 2
    front = 2 * 9 * d * F1 * T1 * B + 2 * 9 * d * d * F2 * T2 * B + 2 * F2 * d * d * T2 * B
    per_tok = 32 * d * d + 8 * d * d + (6 * d * d + 2 * ksz * d) + 6 * T2 * d
    per_layer = per_tok * B * T2 + 2 * d * d * (2 * T2 - 1)
    return float(front + n_blocks * per_layer)
