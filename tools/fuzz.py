#!/usr/bin/env python
"""Shape fuzz (debug): random batch sizes, utterance lengths and length vectors through a 3-block Conformer-M encoder;
the bf16 tensor-core path must stay within 2e-2 (norm-wise) of the fp32 CUDA-core path - which the GPU suite pins to the
reference - and nothing may hang (watchdog) or produce non-finite values.  Exercises every shape-dependent schedule:
FFN tile sharing for any tile count, attention ring / S-buffer rotation for 1..40 key tiles, conv-module tile edges."""
import faulthandler
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict, synth_logmel  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
faulthandler.enable()
rng = random.Random(7)
cfg = EncoderConfig(80, 3, 256, 4, 31)
sd = synth_encoder_state_dict(cfg, 3)
enc = cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size)
enc.load_state_dict(sd)
enc = enc.cuda().eval()
worst = 0.0
with torch.no_grad():
    for trial in range(N):
        faulthandler.dump_traceback_later(60, exit=True)
        kind = trial % 4
        if kind == 0:        # many short utterances
            B, Tin = rng.randint(1, 200), rng.randint(7, 300)
        elif kind == 1:      # few long ones
            B, Tin = rng.randint(1, 6), rng.randint(1000, 10000)
        elif kind == 2:      # around the tile edges of the kernels (T' near multiples of 64 / 96 / 128)
            B, Tin = rng.randint(1, 40), 4 * rng.choice([63, 64, 65, 95, 96, 97, 127, 128, 129, 191, 192, 193, 383, 384, 385]) + rng.randint(3, 6)
        else:
            B, Tin = rng.randint(1, 70), rng.randint(7, 2500)
        while B * Tin > 400000:
            B = max(1, B // 2)
        x = synth_logmel(B, 80, Tin, seed=100 + trial).cuda()
        lengths = None
        if trial % 3 != 0:
            ls = sorted((rng.randint(max(7, Tin // 4), Tin) for _ in range(B)), reverse=True)
            ls[0] = Tin
            lengths = torch.tensor(ls, device="cuda")
        y32, l32 = enc.set_compute("fp32")(x, lengths)
        y16, l16 = enc.set_compute("bf16")(x, lengths)
        torch.cuda.synchronize()
        assert torch.isfinite(y16).all() and torch.isfinite(y32).all(), (trial, B, Tin)
        if lengths is not None:
            assert torch.equal(l32, l16)
        rel = ((y16 - y32).norm() / y32.norm()).item()
        worst = max(worst, rel)
        print(f"trial {trial}: B={B} Tin={Tin} T'={y32.shape[1]} lengths={'yes' if lengths is not None else 'no'} rel {rel:.2e}",
              file=sys.stderr, flush=True)
        assert rel < 2e-2, (trial, B, Tin, rel)
        faulthandler.cancel_dump_traceback_later()
print(f"fuzz done: {N} shapes, worst bf16-vs-fp32 relative error {worst:.2e}", file=sys.stderr)
