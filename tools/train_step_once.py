#!/usr/bin/env python
"""A few C4 training steps (for ncu captures of the training kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import CONFORMER_M, EncoderConfig, synth_decoder_state_dict, synth_encoder_state_dict  # noqa: E402

cfg = EncoderConfig(80, int(os.environ.get("LAYERS", "2")), 256, 4, 31)
B, Tin = int(os.environ.get("B", "32")), 1501
dev = torch.device("cuda", 0)
sd = synth_encoder_state_dict(cfg, 0, prefix="encoder.")
sd.update(synth_decoder_state_dict(cfg.d_model, bench.VOCAB, seed=1))
model = cb.Conformer(bench.VOCAB, cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, dropout_rate=0.1)
model.load_state_dict(sd)
model = model.to(dev).train().set_compute("bf16")
x, xl, y, yl = bench.train_batch(cfg, B, Tin, 1234, dev)
for _ in range(int(os.environ.get("STEPS", "2"))):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, ol = model(x, xl)
    loss = bench.ctc_loss_fn(out, y, ol, yl)
    model.zero_grad()
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
