"""Case definitions shared by make_golden.py (which runs the real reference) and the tests
(which regenerate the same weights/inputs from seeds and compare with the stored outputs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from conformer_b200.synth import CONFORMER_L, CONFORMER_M, CONFORMER_S, EncoderConfig  # noqa: E402

TINY = EncoderConfig(n_mel_channels=80, n_blocks=2, d_model=32, n_heads=2, kernel_size=7)

# name, cfg, B, Tin, seed, with_lengths
ENCODER_CASES = [
    dict(name="tiny", cfg=TINY, B=3, Tin=83, seed=11, with_lengths=True),
    dict(name="tiny_nolen", cfg=TINY, B=2, Tin=64, seed=12, with_lengths=False),
    dict(name="S", cfg=CONFORMER_S, B=2, Tin=201, seed=21, with_lengths=True),       # T'=49
    dict(name="M", cfg=CONFORMER_M, B=3, Tin=301, seed=22, with_lengths=True),       # T'=74
    dict(name="M_T150", cfg=CONFORMER_M, B=2, Tin=605, seed=23, with_lengths=True),  # T'=150 (>1 key tile)
    dict(name="L", cfg=CONFORMER_L, B=2, Tin=201, seed=24, with_lengths=True),       # T'=49
    # BASELINE configs at their own shape: C1 in full (Conformer-S, 4 x 10 s, lengths all 1001) and one
    # long-form utterance of C5 (Conformer-M, 60 s -> T'=1499)
    dict(name="C1", cfg=CONFORMER_S, B=4, Tin=1001, seed=25, with_lengths="full"),   # T'=249
    dict(name="C5_B1", cfg=CONFORMER_M, B=1, Tin=6001, seed=26, with_lengths="full"),  # T'=1499
]


def case_lengths(case):
    """lengths tensor of an encoder case: None, synth_lengths (sorted, first = Tin) or all = Tin ("full")."""
    import torch
    from conformer_b200.synth import synth_lengths
    if not case["with_lengths"]:
        return None
    if case["with_lengths"] == "full":
        return torch.full((case["B"],), case["Tin"], dtype=torch.long)
    return synth_lengths(case["B"], case["Tin"])

SUBMODULE_CFG = dict(cfg=EncoderConfig(80, 1, 64, 4, 31), B=3, T=41, seed=5, x_seed=6,
                     lengths=[41, 30, 1], Tin=171, mel_lengths=[171, 100, 9])

TOKEN_CASE = dict(cfg=EncoderConfig(80, 4, 144, 4, 31), B=4, Tin=401, seed=31, vocab=370, peaky=8.0)
# d_model 256 so that the bf16 tensor-core path takes it: the bf16 greedy-token gate
TOKEN_CASE_BF16 = dict(cfg=EncoderConfig(80, 4, 256, 4, 31), B=4, Tin=401, seed=33, vocab=370, peaky=8.0, name="tokens_M")

# tiny Conformer whose checkpoint file is written by the reference's own CheckpointManager (checkpoint interop, 8f-4)
TINY_CKPT = dict(cfg=TINY, seed=41, vocab=37, hidden=48)

# waveforms for the log-mel front-end (8f-3): [B, L] samples at 16 kHz
LOGMEL_CASES = [
    dict(name="short", B=2, L=4000, seed=51),                       # 26 frames
    dict(name="odd", B=3, L=16001 + 77, seed=52, pad_from=9000),    # L not a multiple of hop; one zero-padded utterance
    dict(name="sec3", B=2, L=48000, seed=53),                       # 301 frames
]

# the reference Decoder alone: two LSTM layers, unsorted lengths with max == T (decoder.py:15-27)
DECODER_CASE = dict(d=64, vocab=37, hidden=96, layers=2, B=5, T=33, lengths=[20, 33, 1, 33, 7], seed=71)

# training-step gradients from the reference's own autograd (config 4): Encoder in train mode, dropout 0,
# loss = sum(y * R) with a seeded cotangent R
TRAIN_CASES = [
    dict(name="train_tiny", cfg=TINY, B=3, Tin=83, seed=81, with_lengths=True),
    dict(name="train_k31", cfg=EncoderConfig(80, 1, 64, 4, 31), B=2, Tin=131, seed=82, with_lengths=False),
]
