#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM through the C ABI on the shapes the Conformer-M/L blocks use.
   python tools/gemm_bench.py [--shapes ffn1,ffn2,...] [--iters 20]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conformer_b200._lib import check, lib  # noqa: E402

M = 23936
SHAPES = {  # name: (M, N, K, epi)
    "ffn1": (M, 1024, 256, 1), "ffn2": (M, 256, 1024, 3), "qkv": (M, 768, 256, 0), "out": (M, 256, 256, 3),
    "pw1": (M, 512, 256, 0), "lin": (M, 256, 4864, 4), "L_ffn1": (15968, 2048, 512, 1), "L_ffn2": (15968, 512, 2048, 3),
    "big": (8192, 8192, 8192, 0),
}

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="ffn1,ffn2,qkv,out,pw1,lin,L_ffn1,L_ffn2,big")
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in args.shapes.split(","):
    m, n, k, epi = SHAPES[name]
    A = torch.randn(m, k, device="cuda").bfloat16()
    W = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    bias = torch.randn(n, device="cuda")
    out = torch.zeros(m, n, device="cuda", dtype=torch.float32 if epi >= 3 else torch.bfloat16)
    run = lambda: check(lib().cfm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), m, n, k, epi,  # noqa: E731
                                            out.data_ptr(), 0.5, st))
    for _ in range(3):
        run()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    # torch reference (cuBLAS) for context
    tt = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(A, W.T); e1.record()
        torch.cuda.synchronize()
        tt.append(e0.elapsed_time(e1))
    tt.sort()
    fl = 2.0 * m * n * k
    print(f"{name:8s} M={m} N={n} K={k} epi={epi}: {med * 1e3:8.1f} us  {fl / med / 1e9:7.1f} TFLOP/s   "
          f"(cuBLAS plain matmul {tt[len(tt) // 2] * 1e3:8.1f} us {fl / tt[len(tt) // 2] / 1e9:7.1f} TFLOP/s)", flush=True)
