#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  ``PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py``

For every case it builds the reference module from /root/reference, loads the seeded
weights of ``conformer_b200.synth`` with a strict ``load_state_dict`` (so the key/shape
contract is checked against the real thing), runs it on CPU in fp32 (and fp64 for a
noise floor), and stores the outputs as .npz.  Weights and inputs are NOT stored: they are
regenerated from (config, seed) by ``conformer_b200.synth``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from conformer_b200.synth import (EncoderConfig, synth_decoder_state_dict,   # noqa: E402
                                  synth_encoder_state_dict, synth_lengths, synth_logmel)
from cases import ENCODER_CASES, SUBMODULE_CFG, TINY_CKPT, TOKEN_CASE, TOKEN_CASE_BF16, case_lengths  # noqa: E402

from model.conformer import Conformer                                        # noqa: E402
from model.modules.encoder import Encoder                                    # noqa: E402
from model.utils.attention import MultiHeadSelfAttentionModule              # noqa: E402
from model.utils.block import ConformerBlock                                 # noqa: E402
from model.utils.convolution import ConvolutionModule, ConvolutionSubsampling  # noqa: E402
from model.utils.ffn import FeedForwardModule                                # noqa: E402
from model.utils.masking import generate_padding_mask                        # noqa: E402
from model.utils.position import RelativePositionalEncoding                  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(8)


def sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def reset_mask_value(m):
    for mod in m.modules():
        if hasattr(mod, "mask_value"):
            mod.mask_value = None


def run_encoder_case(name, cfg: EncoderConfig, B, Tin, seed, with_lengths):
    sd = synth_encoder_state_dict(cfg, seed)
    x = synth_logmel(B, cfg.n_mel_channels, Tin, seed=1000 + seed)
    lengths = case_lengths(dict(B=B, Tin=Tin, with_lengths=with_lengths))
    enc = Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size).eval()
    enc.load_state_dict(sd, strict=True)
    y32, l32 = enc(x, lengths)
    out = {"y": y32.numpy()}
    if with_lengths:
        out["out_lengths"] = l32.numpy()
    if cfg.d_model <= 256:                       # fp64 truth for the noise floor (skip the 460 MB L model)
        enc64 = enc.double()
        reset_mask_value(enc64)
        y64, _ = enc64(x.double(), lengths)
        out["y64"] = y64.numpy()
    np.savez_compressed(os.path.join(HERE, f"encoder_{name}.npz"), **out)
    print(f"encoder_{name}: y {tuple(y32.shape)} mean|y|={y32.abs().mean():.4f}")


def run_submodules():
    cfg, B, T, seed = SUBMODULE_CFG["cfg"], SUBMODULE_CFG["B"], SUBMODULE_CFG["T"], SUBMODULE_CFG["seed"]
    d, H, k = cfg.d_model, cfg.n_heads, cfg.kernel_size
    sd = synth_encoder_state_dict(cfg, seed)
    g = torch.Generator().manual_seed(SUBMODULE_CFG["x_seed"])
    x = torch.randn(B, T, d, generator=g)
    lengths = torch.tensor(SUBMODULE_CFG["lengths"], dtype=torch.long)
    out = {}

    ffn = FeedForwardModule(d).eval()
    ffn.load_state_dict(sub(sd, "layers.0.ffn_1."), strict=True)
    out["ffn"] = ffn(x).numpy()

    pe = RelativePositionalEncoding(d)
    pe.load_state_dict(sub(sd, "rel_pe."), strict=True)
    pos = pe(x)                                         # [B, 2T-1, d]
    out["rel_pe"] = pos[0].numpy()
    assert torch.equal(pos[0], pos[-1])

    valid = generate_padding_mask(lengths)              # [B, T] True = valid
    out["padding_mask"] = valid.numpy()
    mask = (~valid)[:, None, None, :]

    att = MultiHeadSelfAttentionModule(d, H).eval()
    att.load_state_dict(sub(sd, "layers.0.attention."), strict=True)
    out["mhsa_nomask"] = att(x, pos, None).numpy()
    reset_mask_value(att)
    out["mhsa_mask"] = att(x, pos, mask).numpy()
    # raw relative shift on a random score tensor
    ps = torch.randn(2, H, T, 2 * T - 1, generator=g)
    out["shift_in"] = ps.numpy()
    out["shift_out"] = att.attention._relative_shift(ps).numpy()

    conv = ConvolutionModule(d, k).eval()
    conv.load_state_dict(sub(sd, "layers.0.conv."), strict=True)
    out["conv"] = conv(x).numpy()

    blk = ConformerBlock(d, H, k).eval()
    blk.load_state_dict(sub(sd, "layers.0."), strict=True)
    out["block_nomask"] = blk(x, pos, None).numpy()
    reset_mask_value(blk)
    out["block_mask"] = blk(x, pos, mask).numpy()

    subs = ConvolutionSubsampling(d).eval()
    subs.load_state_dict(sub(sd, "downsampling_conv."), strict=True)
    mel = synth_logmel(B, cfg.n_mel_channels, SUBMODULE_CFG["Tin"], seed=77)
    mel_len = torch.tensor(SUBMODULE_CFG["mel_lengths"], dtype=torch.long)
    ys, ls = subs(mel, mel_len)
    out["subsampling"] = ys.numpy()
    out["subsampling_lengths"] = ls.numpy()
    np.savez_compressed(os.path.join(HERE, "submodules.npz"), **out)
    print("submodules:", {k_: v.shape for k_, v in out.items()})


def run_tokens(case=TOKEN_CASE):
    cfg, B, Tin, seed, vocab = (case[k] for k in ("cfg", "B", "Tin", "seed", "vocab"))
    sd = synth_encoder_state_dict(cfg, seed, prefix="encoder.")
    sd.update(synth_decoder_state_dict(cfg.d_model, vocab, seed=seed + 1, peaky=case["peaky"]))
    model = Conformer(vocab, cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size).eval()
    model.load_state_dict(sd, strict=True)
    x = synth_logmel(B, cfg.n_mel_channels, Tin, seed=2000 + seed)
    lengths = synth_lengths(B, Tin)
    logits, out_len = model(x, lengths)
    ids = logits.argmax(-1)
    # greedy decode, processing/processor.py:301-319 (processor.py itself needs librosa, absent here):
    pad_id, unk_id = 0, vocab - 1
    toks = []
    for row in ids:
        items, prev = [], None
        for t in row.tolist():
            if t == pad_id or t == unk_id:
                continue
            if prev is None:
                prev = t
                items.append(t)
            elif prev != t:
                prev = t
                items.append(t)
        toks.append(items)
    top2 = logits.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    flat = np.full((B, max(len(t) for t in toks)), -1, dtype=np.int64)
    for i, t in enumerate(toks):
        flat[i, :len(t)] = t
    np.savez_compressed(os.path.join(HERE, case.get("name", "tokens") + ".npz"), logits=logits.numpy(), ids=ids.numpy(),
                        tokens=flat, out_lengths=out_len.numpy(), margin=margin.numpy())
    print(f"tokens: logits {tuple(logits.shape)} median margin {margin.median():.4f} "
          f"min margin {margin.min():.5f} tokens/utt {[len(t) for t in toks]}")


if __name__ == "__main__":
    only = set(sys.argv[1:])                       # optional: names of the fixtures to (re)generate
    if not only or "submodules" in only:
        run_submodules()
    for c in ENCODER_CASES:
        if not only or c["name"] in only:
            run_encoder_case(**c)
    if not only or "tokens" in only:
        run_tokens()
    if not only or "tokens_M" in only:
        run_tokens(TOKEN_CASE_BF16)


def dump_state_dict_keys():
    """Key/shape table of the real reference modules, for the CPU-only contract test."""
    import json
    enc = Encoder(80, 2, 32, 2, 7)
    conf = Conformer(370, 80, 2, 32, 2, 7, 64, 1)
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as fh:
        json.dump({"encoder": {k: list(v.shape) for k, v in enc.state_dict().items()},
                   "conformer": {k: list(v.shape) for k, v in conf.state_dict().items()}}, fh, indent=0)


if __name__ == "__main__":
    dump_state_dict_keys()


def dump_reference_checkpoint():
    """A checkpoint file written by the REFERENCE's own CheckpointManager (manager.py:33-41) for a DDP-style wrapped tiny
    Conformer: the fixture of the checkpoint-interop test (keys carry the 'module.' prefix, as train.py:185-186 produces)."""
    import tempfile
    from manager import CheckpointManager
    cfg = TINY_CKPT["cfg"]
    sd = synth_encoder_state_dict(cfg, TINY_CKPT["seed"], prefix="encoder.")
    sd.update(synth_decoder_state_dict(cfg.d_model, TINY_CKPT["vocab"], hidden=TINY_CKPT["hidden"], seed=TINY_CKPT["seed"] + 1))
    model = Conformer(TINY_CKPT["vocab"], cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size,
                      TINY_CKPT["hidden"], 1)
    model.load_state_dict(sd, strict=True)

    class Wrapped(torch.nn.Module):                    # what DistributedDataParallel does to the key names
        def __init__(self, m):
            super().__init__()
            self.module = m
    wrapped = Wrapped(model)
    opt = torch.optim.Adam(wrapped.parameters(), lr=2e-5)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    with tempfile.TemporaryDirectory() as td:
        CheckpointManager(td).save_checkpoint(wrapped, opt, sched, 7, 1)
        with open(os.path.join(td, "7.pt"), "rb") as src, open(os.path.join(HERE, "ref_checkpoint_tiny_ddp.pt"), "wb") as dst:
            dst.write(src.read())
    x = synth_logmel(2, cfg.n_mel_channels, 83, seed=TINY_CKPT["seed"] + 2)
    lengths = synth_lengths(2, 83)
    logits, out_len = model.eval()(x, lengths)
    np.savez_compressed(os.path.join(HERE, "ref_checkpoint_tiny_out.npz"), logits=logits.numpy(), out_lengths=out_len.numpy())
    print("ref_checkpoint_tiny_ddp.pt:", os.path.getsize(os.path.join(HERE, "ref_checkpoint_tiny_ddp.pt")), "bytes")


if __name__ == "__main__":
    if not set(sys.argv[1:]) or "checkpoint" in sys.argv[1:]:
        dump_reference_checkpoint()


def run_logmel():
    """torchaudio.transforms.MelSpectrogram with the reference's arguments (processing/processor.py:53-63) followed by the
    reference's log-clamp (:155-158), in fp32 (what the reference runs) and fp64 (noise floor)."""
    from torchaudio.transforms import MelSpectrogram
    sys.path.insert(0, ROOT)
    from oracle.logmel_oracle import synth_waveform
    from cases import LOGMEL_CASES
    out = {}
    for c in LOGMEL_CASES:
        x = synth_waveform(c["B"], c["L"], c["seed"])
        if c.get("pad_from"):                               # a zero-padded shorter utterance, as the collate makes them
            x[1, c["pad_from"]:] = 0.0
        ms = MelSpectrogram(sample_rate=16000, n_fft=400, win_length=400, hop_length=160, f_min=0.0, f_max=8000.0, n_mels=80,
                            norm="slaney", mel_scale="slaney")
        out[c["name"]] = torch.log(torch.clamp(ms(x), min=1e-5)).numpy()
        out[c["name"] + "_f64"] = torch.log(torch.clamp(ms.double()(x.double()), min=1e-5)).numpy()
        print("logmel", c["name"], out[c["name"]].shape, "mean", out[c["name"]].mean())
    np.savez_compressed(os.path.join(HERE, "logmel.npz"), **out)


if __name__ == "__main__":
    if not set(sys.argv[1:]) or "logmel" in sys.argv[1:]:
        run_logmel()


def run_decoder():
    """The reference Decoder (model/modules/decoder.py) with TWO LSTM layers and unsorted lengths, eval mode."""
    from model.modules.decoder import Decoder
    from cases import DECODER_CASE as c
    sd = sub(synth_decoder_state_dict(c["d"], c["vocab"], hidden=c["hidden"], seed=c["seed"]), "decoder.")
    g = torch.Generator().manual_seed(c["seed"] + 1)
    for l in range(1, c["layers"]):                           # synth_decoder_state_dict makes layer 0 only
        bound = 1.0 / c["hidden"] ** 0.5
        for name, shape in (("weight_ih", (4 * c["hidden"], c["hidden"])), ("weight_hh", (4 * c["hidden"], c["hidden"])),
                            ("bias_ih", (4 * c["hidden"],)), ("bias_hh", (4 * c["hidden"],))):
            sd[f"lstm.{name}_l{l}"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    dec = Decoder(c["vocab"], c["d"], c["hidden"], c["layers"]).eval()
    dec.load_state_dict(sd, strict=True)
    x = torch.randn(c["B"], c["T"], c["d"], generator=g)
    lengths = torch.tensor(c["lengths"])
    out = {"x": x.numpy(), "logits": dec(x, lengths).numpy(), "logits_nolen": dec(x, None).numpy()}
    out.update({"sd." + k: v.numpy() for k, v in sd.items()})
    np.savez_compressed(os.path.join(HERE, "decoder_2layer.npz"), **out)
    print("decoder_2layer:", out["logits"].shape)


if __name__ == "__main__":
    if not set(sys.argv[1:]) or "decoder" in sys.argv[1:]:
        run_decoder()


def run_train_grads():
    """Reference Encoder in TRAINING mode (batch-statistics BatchNorm, dropout 0) under autograd, CPU fp32 and fp64."""
    from cases import TRAIN_CASES
    torch.set_grad_enabled(True)
    for c in TRAIN_CASES:
        cfg = c["cfg"]
        out = {}
        for dt, tag in ((torch.float32, ""), (torch.float64, "64")):
            sd = synth_encoder_state_dict(cfg, c["seed"])
            enc = Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, dropout_rate=0.0)
            enc.load_state_dict(sd, strict=True)
            enc = enc.to(dt).train()
            x = synth_logmel(c["B"], cfg.n_mel_channels, c["Tin"], seed=1000 + c["seed"]).to(dt)
            lengths = case_lengths(c)
            y, _ = enc(x, lengths)
            R = torch.randn(y.shape, generator=torch.Generator().manual_seed(c["seed"] + 7)).to(dt)
            (y * R).sum().backward()
            out["y" + tag] = y.detach().numpy()
            for n, p in enc.named_parameters():
                if p.requires_grad:
                    out[f"grad{tag}." + n] = p.grad.numpy()
            if not tag:
                for n, b in enc.named_buffers():
                    if "running" in n:
                        out["buf." + n] = b.numpy()
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), **out)
        print(c["name"], "y", out["y"].shape, "grads", sum(1 for k in out if k.startswith("grad.")))
    torch.set_grad_enabled(False)


if __name__ == "__main__":
    if not set(sys.argv[1:]) or "train" in sys.argv[1:]:
        run_train_grads()
