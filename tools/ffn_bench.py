#!/usr/bin/env python
"""Micro-benchmark of the fused FeedForwardModule kernel (bf16, d=256) through the C ABI."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

B, T, d = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 374, 256)))
sd = {k[len("layers.0.ffn_1."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, 4, 31), 1).items()
      if k.startswith("layers.0.ffn_1.")}
m = cb.FeedForwardModule(d)
m.load_state_dict(sd)
m = m.cuda().eval().set_compute("bf16")
x = torch.randn(B, T, d, device="cuda")
for _ in range(3):
    m(x)
m.profile(True)
for _ in range(10):
    m(x)
torch.cuda.synchronize()
for k, v in m.collect_profile().items():
    if v["launches"]:
        print(k, f"{v['ms'] / v['launches'] * 1e3:.1f} us/launch", f"{v['flops'] / v['ms'] / 1e9:.1f} TFLOP/s" if v["flops"] else "")
