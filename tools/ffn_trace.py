#!/usr/bin/env python
"""Wait-time breakdown (ns) of the MMA-issuing warp of CTA 0 of the fused FFN kernel (debug)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200._lib import lib  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

B, T, d = 64, 374, 256
sd = {k[len("layers.0.ffn_1."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, 4, 31), 1).items()
      if k.startswith("layers.0.ffn_1.")}
m = cb.FeedForwardModule(d)
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
t = trace.cpu()
print("MMA warp of pair 0's leader: total %d ns; waiting for LN(x) %d, weights (full_bar) %d, hidden (hs_full) %d, acc2 drained %d"
      % tuple(int(v) for v in t[:5]))
t0 = int(t[8])
print("G2 issued      ", [int(v) for v in t[16:40]])
print("epilogue done  ", [int(v) - t0 if v else 0 for v in t[48:72]])
print("final start", [int(v) - t0 if v else 0 for v in t[80:84]], "end", [int(v) - t0 if v else 0 for v in t[84:88]])
