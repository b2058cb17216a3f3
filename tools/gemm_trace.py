#!/usr/bin/env python
"""Timeline of CTA 0 of one tcgen05 GEMM (debug): per tile, ns relative to the first stamp.
slots: 0 producer first load issued, 1 producer last load issued, 2 MMA got accumulator, 3 MMA first k-block
landed, 4 MMA last k-block landed, 5 epilogue staging free, 6 epilogue accumulator ready, 7 epilogue store issued"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conformer_b200._lib import check, lib  # noqa: E402

m, n, k, epi = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (23936, 1024, 256, 1))]
A = torch.randn(m, k, device="cuda").bfloat16()
W = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
bias = torch.randn(n, device="cuda")
out = torch.zeros(m, n, device="cuda", dtype=torch.float32 if epi >= 3 else torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
run = lambda: check(lib().cfm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), m, n, k, epi,  # noqa: E731
                                        out.data_ptr(), 0.5, st))
for _ in range(3):
    run()
trace = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
fn = lib().cfm_debug_set_gemm_trace
fn.argtypes = [C.c_void_p]
fn(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
fn(None)
t = trace.cpu().view(64, 8)
t0 = int(t[t > 0].min())
print(f"M={m} N={n} K={k} epi={epi}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us")
for i in range(64):
    if int(t[i].max()) == 0:
        break
    print(f"tile {i:2d}: " + " ".join(f"{(int(v) - t0) if v > 0 else -1:7d}" for v in t[i]))
