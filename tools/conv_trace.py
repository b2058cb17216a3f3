#!/usr/bin/env python
"""Timeline (ns) of one CTA of the fused conv-module kernel (debug): [0] start [1] LN done [2] pw1 half0 ready
[3] half1 ready [4] GLU done [5] depthwise start [6] depthwise done [7] pw2 ready [8] staged [9] stored"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200._lib import lib  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

B, T, d = 64, 374, 256
sd = {k[len("layers.0.conv."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, 4, 31), 1).items()
      if k.startswith("layers.0.conv.")}
m = cb.ConvolutionModule(d, 31)
m.load_state_dict(sd)
m = m.cuda().eval().set_compute("bf16")
x = torch.randn(B, T, d, device="cuda")
for _ in range(3):
    m(x)
trace = torch.zeros(128, dtype=torch.int64, device="cuda")
fn = lib().cfm_debug_set_gemm_trace
fn.argtypes = [C.c_void_p]
torch.cuda.synchronize()
fn(trace.data_ptr())
m(x)
torch.cuda.synchronize()
fn(None)
t = trace.cpu()[:10]
print(" ".join(f"{int(v - t[0]):6d}" for v in t))
