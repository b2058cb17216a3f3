#!/usr/bin/env python
"""Stream launches vs CUDA-graph replay of the Conformer-M encoder forward (64 x 15 s, bf16): one stream and two streams."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_b200 as cb  # noqa: E402
from conformer_b200.synth import EncoderConfig, synth_encoder_state_dict, synth_logmel  # noqa: E402

cfg = EncoderConfig(80, 16, 256, 4, 31)
enc = cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size)
enc.load_state_dict(synth_encoder_state_dict(cfg, 0))
enc = enc.cuda().eval().set_compute("bf16")
x = synth_logmel(64, 80, 1501, seed=1).cuda()
K = 40


def timed(step, streams):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for st in streams:
        st.wait_event(e0)
    for i in range(K):
        with torch.cuda.stream(streams[i % len(streams)]):
            step(i % len(streams))
    for st in streams:
        cur.wait_stream(st)
    e1.record(cur)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


with torch.inference_mode():
    streams = [torch.cuda.Stream() for _ in range(2)]
    for _ in range(5):
        enc(x)
    y_ref = enc(x)[0].clone()
    caps = [enc.capture(x) for _ in range(2)]
    assert torch.equal(caps[0].replay()[0], caps[0].y)
    torch.cuda.synchronize()
    print("max |graph - stream| =", (caps[0].y - y_ref).abs().max().item())
    for name, fn in (("stream launches", lambda i: enc(x)), ("graph replay", lambda i: caps[i].replay())):
        for _ in range(4):
            fn(0); fn(1)
        print(f"{name:16s}: 1 stream {timed(fn, streams[:1]):.3f} ms/step, 2 streams {timed(fn, streams):.3f} ms/step")
