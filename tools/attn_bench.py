#!/usr/bin/env python
"""Micro-benchmark of the relative-position attention module core (bf16, head size 64) through the C ABI."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

B, T, d, H = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 374, 256, 4)))
sd = {k[len("layers.0.attention.attention."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, H, 31), 1).items()
      if k.startswith("layers.0.attention.attention.")}
m = cb.RelativeMultiHeadAttention(d, H)
m.load_state_dict(sd)
m = m.cuda().eval().set_compute("bf16")
x = torch.randn(B, T, d, device="cuda")
pos = torch.randn(1, 2 * T - 1, d, device="cuda")
if os.environ.get("ATTN_POS", "none") == "none":
    pos = None          # module's own sinusoids -> factorised kernel
for _ in range(3):
    m(x, x, x, pos, None)
m.profile(True)
for _ in range(int(sys.argv[5]) if len(sys.argv) > 5 else 10):
    m(x, x, x, pos, None)
torch.cuda.synchronize()
for k, v in m.collect_profile().items():
    if v["launches"]:
        print(k, f"{v['ms'] / v['launches'] * 1e3:.1f} us/launch", f"{v['flops'] / v['ms'] / 1e9:.1f} TFLOP/s" if v["flops"] else "")
