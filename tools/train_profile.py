#!/usr/bin/env python
"""Kernel-time breakdown of one C4 training step (torch.profiler / CUPTI): which kernels the step spends its time in."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import CONFORMER_M, synth_decoder_state_dict, synth_encoder_state_dict  # noqa: E402

cfg, B, Tin = CONFORMER_M, int(os.environ.get("B", "32")), 1501
dev = torch.device("cuda", 0)
sd = synth_encoder_state_dict(cfg, 0, prefix="encoder.")
sd.update(synth_decoder_state_dict(cfg.d_model, bench.VOCAB, seed=1))
model = cb.Conformer(bench.VOCAB, cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, dropout_rate=0.1)
model.load_state_dict(sd)
model = model.to(dev).train().set_compute("bf16")
opt = torch.optim.Adam(model.parameters(), lr=2e-5)
x, xl, y, yl = bench.train_batch(cfg, B, Tin, 1234, dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, ol = model(x, xl)
    loss = bench.ctc_loss_fn(out, y, ol, yl)
    opt.zero_grad()
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"total device time {total / 1e3:.2f} ms")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f}%  x{e.count:4d}  {e.key[:110]}")
