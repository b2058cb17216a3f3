#!/usr/bin/env python
"""Timeline (ns) of one CTA of the tcgen05 attention kernels (debug; the stamps perturb the traced CTA).
per tile j: MMA: [0] got TMEM buffer [1] K/V/window landed [2] P_j ready -> issue PV | softmax (warp 4):
[3] top [4] O_{j-1} ready [5] S,PS_j ready [6] max exchanged [7] P_j published"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200._lib import lib  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict  # noqa: E402

B, T, d, H = 64, 374, 256, 4          # the traced CTA is batch index 30 (third wave)
sd = {k[len("layers.0.attention.attention."):]: v for k, v in synth_encoder_state_dict(EncoderConfig(80, 1, d, H, 31), 1).items()
      if k.startswith("layers.0.attention.attention.")}
m = cb.RelativeMultiHeadAttention(d, H)
m.load_state_dict(sd)
m = m.cuda().eval().set_compute("bf16")
x = torch.randn(B, T, d, device="cuda")
pos = torch.randn(1, 2 * T - 1, d, device="cuda")
POS = None if os.environ.get('ATTN_POS', 'none') == 'none' else pos
for _ in range(3):
    m(x, x, x, POS, None)
trace = torch.zeros(128, dtype=torch.int64, device="cuda")
fn = lib().cfm_debug_set_gemm_trace
fn.argtypes = [C.c_void_p]
torch.cuda.synchronize()
fn(trace.data_ptr())
m(x, x, x, POS, None)
torch.cuda.synchronize()
fn(None)
t = trace.cpu()[64:]
if os.environ.get("ATTN_POS", "none") == "none":
    t0 = int(t[0])
    print("factorised kernel. prologue: start 0, Q written %d, ab ready %d, Q' written %d, end %d" % tuple(int(t[i]) - t0 for i in (1, 2, 3, 4)))
    print("SM clock during the traced CTA: %.0f MHz" % ((int(t[7]) - int(t[6])) * 1e3 / max(1, int(t[4]) - int(t[0]))))
    for j in range(8):
        r = t[8 + 4 * j: 12 + 4 * j]
        if int(r.max()) == 0:
            break
        print(f"tile {j}: S ready {int(r[1]) - t0:6d}  exp done {int(r[2]) - t0:6d}  P written {int(r[3]) - t0:6d}")
    for j in range(7):
        r = t[36 + 4 * j: 40 + 4 * j]
        if int(r.max()) == 0:
            break
        landed = f"K|E_{j + 3} landed {int(r[3]) - t0:6d}" if int(r[3]) > 0 else "(no further S)"
        print(f"MMA warp, tile {j}: loop top {int(r[0]) - t0:6d}  P_{j} ready {int(r[2]) - t0:6d}  P_{j}.V issued {int(r[1]) - t0:6d}  "
              f"{landed}  (S_{j + 3} is then issued until the next loop top)")
else:
    t0 = int(t[60])
    print("shift kernel: softmax warps start 0, Q staged %d, first tile top %d, end %d" % (int(t[61]) - t0, int(t[0]) - t0, int(t[62]) - t0))
    for j in range(7):
        r = t[8 * j: 8 * j + 5]
        if int(r.max()) == 0:
            break
        print(f"shift kernel tile {j}: top {int(r[0]) - t0:6d}  S/PS ready {int(r[1]) - t0:6d}  skewed+max {int(r[2]) - t0:6d}  "
              f"rescale done {int(r[3]) - t0:6d}  P written {int(r[4]) - t0:6d}")
