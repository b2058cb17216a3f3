"""Seeded synthetic weights and inputs in the reference's ``state_dict`` contract.

There is no network for checkpoints or datasets, so tests, ``bench.py`` and the golden
fixture generator all draw weights from here: the same (config, seed) gives the same
tensors on every box (CPU ``torch.Generator``), which is what lets the fixtures under
``tests/golden`` hold only *outputs* of the real reference.

Key names and shapes follow the reference modules' ``state_dict``:
``model/modules/encoder.py:12-16``, ``model/utils/block.py:11-15``, ``ffn.py:8-13``,
``attention.py:10-11,31-41``, ``convolution.py:12-18,37-40``, ``position.py:9`` and, for the
CTC head used only by the token-parity harness, ``model/modules/decoder.py:10-13``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch


@dataclass(frozen=True)
class EncoderConfig:
    n_mel_channels: int = 80
    n_blocks: int = 16
    d_model: int = 256
    n_heads: int = 4
    kernel_size: int = 31

    @property
    def freq_bins(self) -> int:            # encoder.py:13
        return ((self.n_mel_channels - 1) // 2 - 1) // 2


CONFORMER_S = EncoderConfig(80, 16, 144, 4, 31)
CONFORMER_M = EncoderConfig(80, 16, 256, 4, 31)
CONFORMER_L = EncoderConfig(80, 17, 512, 8, 31)


def subsampled_frames(t_in: int) -> int:
    """Frames after the two stride-2 3x3 convs (convolution.py:37,39)."""
    return ((t_in - 1) // 2 - 1) // 2


def _uniform(g: torch.Generator, shape, bound: float) -> torch.Tensor:
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _ln(g, d):
    return (1.0 + 0.1 * torch.randn(d, generator=g), 0.05 * torch.randn(d, generator=g))


def _linear(g, sd, name, n_out, n_in, w_shape=None):
    bound = 1.0 / math.sqrt(n_in)
    sd[name + ".weight"] = _uniform(g, w_shape or (n_out, n_in), bound)
    sd[name + ".bias"] = _uniform(g, (n_out,), bound)


def synth_encoder_state_dict(cfg: EncoderConfig, seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    d, H, k = cfg.d_model, cfg.n_heads, cfg.kernel_size
    dh = d // H
    sd: Dict[str, torch.Tensor] = {}
    p = prefix
    _linear(g, sd, p + "downsampling_conv.conv_1", d, 9, (d, 1, 3, 3))
    _linear(g, sd, p + "downsampling_conv.conv_2", d, 9 * d, (d, d, 3, 3))
    _linear(g, sd, p + "linear", d, d * cfg.freq_bins)
    sd[p + "rel_pe.div_term"] = torch.exp(
        torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d)).unsqueeze(0)
    for i in range(cfg.n_blocks):
        b = f"{p}layers.{i}."
        for ffn in ("ffn_1.", "ffn_2."):
            sd[b + ffn + "layer_norm.weight"], sd[b + ffn + "layer_norm.bias"] = _ln(g, d)
            _linear(g, sd, b + ffn + "hidden_linear", 4 * d, d)
            _linear(g, sd, b + ffn + "out_linear", d, 4 * d)
        a = b + "attention."
        sd[a + "layer_norm.weight"], sd[a + "layer_norm.bias"] = _ln(g, d)
        xav = math.sqrt(6.0 / (H + dh))
        sd[a + "attention.content_bias"] = _uniform(g, (H, dh), xav)
        sd[a + "attention.position_bias"] = _uniform(g, (H, dh), xav)
        for proj in ("query_proj", "key_proj", "value_proj", "pos_proj", "out_proj"):
            _linear(g, sd, a + "attention." + proj, d, d)
        c = b + "conv."
        sd[c + "layer_norm.weight"], sd[c + "layer_norm.bias"] = _ln(g, d)
        _linear(g, sd, c + "pointwise_conv_1", 2 * d, d, (2 * d, d, 1))
        _linear(g, sd, c + "deepwise_conv", d, k, (d, 1, k))
        sd[c + "batch_norm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        sd[c + "batch_norm.bias"] = 0.05 * torch.randn(d, generator=g)
        sd[c + "batch_norm.running_mean"] = 0.1 * torch.randn(d, generator=g)
        sd[c + "batch_norm.running_var"] = 0.5 + torch.rand(d, generator=g)
        sd[c + "batch_norm.num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
        _linear(g, sd, c + "pointwise_conv_2", d, d, (d, d, 1))
        sd[b + "layer_norm.weight"], sd[b + "layer_norm.bias"] = _ln(g, d)
    return sd


def synth_decoder_state_dict(d_model: int, vocab_size: int, hidden: int = 640, seed: int = 1,
                             prefix: str = "decoder.", peaky: float = 1.0) -> Dict[str, torch.Tensor]:
    """CTC head weights (decoder.py:10-13).  ``peaky`` scales the output projection so that
    argmax margins are far from bf16 noise (no trained checkpoint ships with the reference)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    bound = 1.0 / math.sqrt(hidden)
    sd[prefix + "lstm.weight_ih_l0"] = _uniform(g, (4 * hidden, d_model), bound * 4.0)   # input-driven,
    sd[prefix + "lstm.weight_hh_l0"] = _uniform(g, (4 * hidden, hidden), bound * 0.25)    # weakly recurrent
    sd[prefix + "lstm.bias_ih_l0"] = _uniform(g, (4 * hidden,), bound)
    sd[prefix + "lstm.bias_hh_l0"] = _uniform(g, (4 * hidden,), bound)
    sd[prefix + "norm.weight"] = 1.0 + 0.1 * torch.randn(hidden, generator=g)
    sd[prefix + "norm.bias"] = 0.05 * torch.randn(hidden, generator=g)
    sd[prefix + "norm.running_mean"] = 0.05 * torch.randn(hidden, generator=g)
    sd[prefix + "norm.running_var"] = 0.05 + 0.1 * torch.rand(hidden, generator=g)
    sd[prefix + "norm.num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
    sd[prefix + "linear.weight"] = _uniform(g, (vocab_size, hidden), bound) * peaky
    sd[prefix + "linear.bias"] = _uniform(g, (vocab_size,), bound)
    return sd


def synth_logmel(batch: int, n_mel: int, t_in: int, seed: int = 1234) -> torch.Tensor:
    """Log-mel-like input: clamp(randn*2-4, min=log(1e-5)); the floor is
    the one applied at processing/processor.py:157."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((batch, n_mel, t_in), generator=g, dtype=torch.float32) * 2.0 - 4.0
    return x.clamp_(min=math.log(1e-5))


def synth_lengths(batch: int, t_in: int, min_frac: float = 0.5) -> torch.Tensor:
    """Sorted-descending frame counts with lengths[0] == t_in (collate pads to the batch max,
    processing/processor.py:376-392; dataset.py:97,106 sorts by length)."""
    if batch == 1:
        return torch.tensor([t_in], dtype=torch.long)
    v = torch.linspace(float(t_in), float(t_in) * min_frac, batch).floor().long()
    v[0] = t_in
    return v


def encoder_flops(batch: int, t_in: int, d: int, n_blocks: int, n_mel: int = 80, ksz: int = 31) -> float:
    """Algorithmic FLOPs of one encoder forward (SURVEY.md section 8d / BASELINE.md section 3): per token per block
    32d^2 (two FFNs) + 8d^2 (q,k,v,out) + 6d^2 + 2*ksz*d (conv module) + 6*T'*d (attention), pos_proj once per layer,
    plus the two subsampling convs and Encoder.linear.  What bench.py's roofline figures are computed from."""
    t1 = (t_in - 1) // 2
    t2 = (t1 - 1) // 2
    f1 = (n_mel - 1) // 2
    f2 = (f1 - 1) // 2
    front = 2 * 9 * d * f1 * t1 * batch + 2 * 9 * d * d * f2 * t2 * batch + 2 * f2 * d * d * t2 * batch
    per_tok = 32 * d * d + 8 * d * d + (6 * d * d + 2 * ksz * d) + 6 * t2 * d
    per_layer = per_tok * batch * t2 + 2 * d * d * (2 * t2 - 1)
    return float(front + n_blocks * per_layer)
