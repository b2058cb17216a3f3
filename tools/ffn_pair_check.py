#!/usr/bin/env python
"""Development check of the CTA-pair fused FFN kernel: bf16 FeedForwardModule against a torch fp32 reference of the same
op on several row counts (tile edges of the 256-row pair tiles), then timing at the C2 shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

d = 256
sd = {k[len("layers.0.ffn_1."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, 4, 31), 1).items()
      if k.startswith("layers.0.ffn_1.")}
m = cb.FeedForwardModule(d)
m.load_state_dict(sd)
m = m.cuda().eval().set_compute("bf16")


def ref(x):
    w = {k: v.cuda().float() for k, v in sd.items()}
    a = torch.nn.functional.layer_norm(x, (d,), w["layer_norm.weight"], w["layer_norm.bias"])
    a = a.bfloat16().float()
    h = a @ w["hidden_linear.weight"].bfloat16().float().T + w["hidden_linear.bias"]
    h = (h * torch.sigmoid(h)).bfloat16().float()
    return h @ w["out_linear.weight"].bfloat16().float().T + w["out_linear.bias"]


worst = 0.0
for B, T in [(1, 1), (1, 127), (1, 128), (1, 129), (1, 255), (1, 256), (1, 257), (2, 300), (3, 374), (7, 511), (64, 374), (128, 296),
             (37, 513)]:
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    x = torch.randn(B, T, d, device="cuda", generator=g)
    y = m(x)
    torch.cuda.synchronize()
    r = ref(x)
    err = (y - r).abs().max().item() / r.abs().max().item()
    y2 = m(x)
    rep = (y2 - y).abs().max().item()
    worst = max(worst, err)
    print(f"B={B:4d} T={T:4d} rows={B * T:6d}  max err / max|ref| = {err:.2e}   run-to-run diff {rep:.1e}", flush=True)
print("worst", worst)
assert worst < 2e-2
x = torch.randn(64, 374, d, device="cuda")
for _ in range(3):
    m(x)
m.profile(True)
for _ in range(20):
    m(x)
torch.cuda.synchronize()
for k, v in m.collect_profile().items():
    if v["launches"]:
        print(k, f"{v['ms'] / v['launches'] * 1e3:.1f} us/launch", f"{v['flops'] / v['ms'] / 1e9:.1f} TFLOP/s" if v["flops"] else "")
