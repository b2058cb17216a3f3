#!/usr/bin/env python
"""Timeline (ns) of CTA 0 of the fused FFN kernel (debug)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200._lib import lib  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

d, M = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 23936
sd = {k[len("layers.0.ffn_1."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, 4, 31), 1).items()
      if k.startswith("layers.0.ffn_1.")}
m = cb.FeedForwardModule(d)
m.load_state_dict(sd)
m = m.cuda().eval().set_compute("bf16")
x = torch.randn(1, M, d, device="cuda")
for _ in range(3):
    m(x)
trace = torch.zeros(64, dtype=torch.int64, device="cuda")
fn = lib().cfm_debug_set_gemm_trace
fn.argtypes = [C.c_void_p]
torch.cuda.synchronize()
fn(trace.data_ptr())
m(x)
torch.cuda.synchronize()
fn(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
rel = lambda i: int(t[i]) - t0 if t[i] > 0 else -1  # noqa: E731
print(f"mma start {rel(0)}  xs ready {rel(1)}  acc2 ready {rel(2)}  staged {rel(3)}  stored {rel(4)}")
for c in range(8):
    b = 8 + 6 * c
    print(f"chunk {c}: g1 start {rel(b):6d} issued {rel(b + 1):6d} | g2 start {rel(b + 2):6d} issued {rel(b + 3):6d} | "
          f"epi acc1 ready {rel(b + 4):6d} hs written {rel(b + 5):6d}")
