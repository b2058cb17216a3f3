"""CPU-only checks of the host side: state_dict contract, synthetic-data determinism, shape rules, the
greedy-CTC restatement's edge cases, and that the product path refuses to run without CUDA."""
import copy
import json
import os

import pytest
import torch

import conformer_b200 as cb
from conformer_b200.synth import (CONFORMER_M, EncoderConfig, subsampled_frames, synth_decoder_state_dict,
                                  synth_encoder_state_dict, synth_lengths, synth_logmel)
from helpers import GOLDEN_DIR
from oracle import conformer_oracle as O


def test_state_dict_contract_matches_reference_key_dump():
    """tests/golden/state_dict_keys.json was dumped from the real reference modules by make_golden.py."""
    want = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))
    enc = cb.Encoder(80, 2, 32, 2, 7)
    got = {k: list(v.shape) for k, v in enc.state_dict().items()}
    assert list(got.keys()) == list(want["encoder"].keys())
    assert got == want["encoder"]
    conf = cb.Conformer(370, 80, 2, 32, 2, 7, 64, 1)
    gotc = {k: list(v.shape) for k, v in conf.state_dict().items()}
    assert gotc == want["conformer"]


def test_synth_weights_load_strict_and_are_deterministic():
    cfg = EncoderConfig(80, 2, 32, 2, 7)
    a, b = synth_encoder_state_dict(cfg, 3), synth_encoder_state_dict(cfg, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["linear.weight"], synth_encoder_state_dict(cfg, 4)["linear.weight"])
    enc = cb.Encoder(80, 2, 32, 2, 7)
    enc.load_state_dict(a, strict=True)
    conf = cb.Conformer(370, 80, 2, 32, 2, 7, 640, 1)
    sd = synth_encoder_state_dict(cfg, 3, prefix="encoder.")
    sd.update(synth_decoder_state_dict(32, 370))
    conf.load_state_dict(sd, strict=True)
    copy.deepcopy(enc)                                   # native handles are never copied


def test_shape_rules():
    assert subsampled_frames(1501) == 374 and subsampled_frames(2001) == 499 and subsampled_frames(6001) == 1499
    assert subsampled_frames(1001) == 249
    assert CONFORMER_M.freq_bins == 19
    l = synth_lengths(32, 2001)
    assert l[0] == 2001 and (l[:-1] >= l[1:]).all() and l[-1] >= 1000
    assert torch.equal(O.subsampled_lengths(torch.tensor([1501, 9, 7])), torch.tensor([374, 1, 1]))
    x = synth_logmel(2, 80, 50, seed=1)
    assert x.min() >= -11.52 and torch.equal(x, synth_logmel(2, 80, 50, seed=1))


def test_greedy_ctc_edge_cases():
    pad, unk = 0, 9
    assert O.greedy_ctc_ids(torch.tensor([], dtype=torch.long), pad, unk) == []
    assert O.greedy_ctc_ids(torch.tensor([0, 0, 9, 0]), pad, unk) == []
    # the blank does NOT reset the previous token (processor.py:309-319): "a blank a" -> one "a"
    assert O.greedy_ctc_ids(torch.tensor([3, 0, 3, 3, 4, 9, 4, 3]), pad, unk) == [3, 4, 3]
    logits = torch.zeros(5, 10)
    logits[torch.arange(5), torch.tensor([2, 2, 0, 5, 5])] = 1.0
    assert O.greedy_ctc_batch(logits[None], pad, unk) == [[2, 5]]


def test_flop_count_matches_baseline_md():
    from conformer_b200.synth import encoder_flops
    for f in (O.encoder_flops, encoder_flops):                               # the oracle's count and the product's agree
        assert abs(f(64, 1501, 256, 16) / 1e9 - 1986.9) < 0.5
        assert abs(f(32, 2001, 512, 17) / 1e9 - 5309.2) < 1.0                # C3
        assert abs(f(4, 1001, 144, 16) / 1e9 - 27.2) < 0.1                   # C1
        assert abs(f(16, 6001, 256, 16) / 1e9 - 2658.7) < 1.0                # C5, B = 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_has_no_cpu_fallback():
    enc = cb.Encoder(80, 1, 32, 2, 7).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.zeros(1, 80, 40))
    with pytest.raises(RuntimeError, match="no CPU path"):
        cb.Swish()(torch.zeros(4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        cb.generate_padding_mask(torch.tensor([3, 2]))


def test_oracle_is_not_imported_by_the_product():
    import pathlib
    root = pathlib.Path(cb.__file__).parent
    for f in list(root.glob("*.py")) + list((root / "csrc").glob("*")):
        assert "oracle" not in f.read_text().replace("oracle/", "").replace("no oracle", ""), f


def test_reference_written_checkpoint_loads_and_round_trips(tmp_path):
    """tests/golden/ref_checkpoint_tiny_ddp.pt was written by the REFERENCE's CheckpointManager for a DDP-style wrapped
    model (keys prefixed 'module.', manager.py:33-41 / train.py:185-186).  It must load strictly into the drop-in
    Conformer (prefix stripped, checkpoint.py:27-35), into a DDP-style wrapper (prefix kept), and the encoder sub-tree into
    a stand-alone Encoder; a checkpoint written here must read back identically."""
    from cases import TINY_CKPT
    from conformer_b200 import checkpoint as ck
    c = TINY_CKPT
    cfg = c["cfg"]
    path = os.path.join(GOLDEN_DIR, "ref_checkpoint_tiny_ddp.pt")
    want = synth_encoder_state_dict(cfg, c["seed"], prefix="encoder.")
    want.update(synth_decoder_state_dict(cfg.d_model, c["vocab"], hidden=c["hidden"], seed=c["seed"] + 1))
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"model", "optimizer", "scheduler", "n_steps", "n_epochs"} and ck.is_ddp_state_dict(raw["model"])

    model = cb.Conformer(c["vocab"], cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, c["hidden"], 1)
    ck.load_model(path, model)                                      # from a path, prefix stripped, strict
    got = model.state_dict()
    assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)

    class Wrapped(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.module = m
    other = cb.Conformer(c["vocab"], cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, c["hidden"], 1)
    ck.load_model(ck.strip_ddp_prefix(raw["model"]), Wrapped(other), world_size=2)     # prefix added for a wrapped model
    assert all(torch.equal(other.state_dict()[k], want[k]) for k in want)

    enc = cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size)
    ck.load_model(path, enc, prefix="encoder.")
    assert all(torch.equal(v, want["encoder." + k]) for k, v in enc.state_dict().items())

    opt = torch.optim.Adam(model.parameters(), lr=2e-5)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    assert ck.load_checkpoint(path, model, opt, sched) == (7, 1)    # optimizer/scheduler state of the reference file
    mgr = ck.CheckpointManager(str(tmp_path), n_savings=2)
    for step in (1, 2, 3):
        mgr.save_checkpoint(model, opt, sched, step, 0)
    assert sorted(os.listdir(tmp_path)) == ["2.pt", "3.pt"]         # rotation keeps the last n_savings
    fresh = cb.Conformer(c["vocab"], cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, c["hidden"], 1)
    assert ck.load_checkpoint(os.path.join(tmp_path, "3.pt"), fresh) == (3, 0)
    assert all(torch.equal(fresh.state_dict()[k], want[k]) for k in want)
    with pytest.raises(RuntimeError):                               # strict: a foreign key is an error, as in the reference
        bad = dict(ck.strip_ddp_prefix(raw["model"]))
        bad["encoder.extra"] = torch.zeros(1)
        ck.load_model(bad, fresh)


def test_launch_mode_switch_needs_no_device():
    """cfm_set_pdl only records how the next kernels are launched (programmatic dependent launch on / off): callable
    on a box without a GPU, any non-zero value means on."""
    from conformer_b200._lib import lib
    assert lib().cfm_set_pdl(0) == 0
    assert lib().cfm_set_pdl(7) == 0
    cb.set_pdl(False)
    cb.set_pdl(True)
