#!/usr/bin/env python
"""Timeline (ns) of step 100 of CTA 0 of the tensor-core LSTM kernel (debug)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200._lib import lib  # noqa: E402
from conformer_b200.synth import synth_decoder_state_dict  # noqa: E402

B, T, d, H, V = 64, 374, 256, 640, 370
sd = {k[len("decoder."):]: v for k, v in synth_decoder_state_dict(d, V, hidden=H, seed=3).items()}
dec = cb.Decoder(V, d, H, 1)
dec.load_state_dict(sd, strict=True)
dec = dec.cuda().eval().set_compute("bf16")
x = torch.randn(B, T, d, device="cuda")
lengths = torch.full((B,), T, dtype=torch.long, device="cuda")
for _ in range(3):
    dec(x, lengths)
trace = torch.zeros(128, dtype=torch.int64, device="cuda")
fn = lib().cfm_debug_set_gemm_trace
fn.argtypes = [C.c_void_p]
torch.cuda.synchronize()
fn(trace.data_ptr())
dec(x, lengths)
torch.cuda.synchronize()
fn(None)
t = trace.cpu()
t0 = int(t[96])
names = ["barrier passed", "copies issued", "first half of h landed", "last k-block reached", "xin loads issued", "accumulator ready",
         "cells done", "arrived"]
print("step 100 of CTA 0: " + ", ".join(f"{n} {int(t[96 + i]) - t0}" for i, n in enumerate(names)))
dec.profile(True)
for _ in range(5):
    dec(x, lengths)
torch.cuda.synchronize()
for k, v in dec.collect_profile().items():
    if v["launches"]:
        print(k, f"{v['ms'] / v['launches'] * 1e3:.1f} us/launch x {v['launches'] // 5} per forward")
