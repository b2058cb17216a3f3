"""Module- and encoder-level parity on the GPU, through the drop-in nn.Modules (ctypes -> C ABI ->
sm_100a kernels), against (a) the golden outputs of the real reference committed under tests/golden
and (b) the oracle on the same seeded inputs."""
import pytest
import torch

import conformer_b200 as cb
from cases import ENCODER_CASES, SUBMODULE_CFG, TOKEN_CASE
from conformer_b200.synth import (CONFORMER_L, CONFORMER_M, synth_decoder_state_dict, synth_encoder_state_dict,
                                  synth_lengths, synth_logmel)
from helpers import assert_close_bf16, assert_close_fp32, bf16_report, golden, sub
from oracle import conformer_oracle as O

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"


def _load(mod, sd):
    mod.load_state_dict(sd, strict=True)
    return mod.to(DEV).eval()


@pytest.fixture(scope="module")
def subcase():
    c = SUBMODULE_CFG
    cfg = c["cfg"]
    sd = synth_encoder_state_dict(cfg, c["seed"])
    g = torch.Generator().manual_seed(c["x_seed"])
    x = torch.randn(c["B"], c["T"], cfg.d_model, generator=g).to(DEV)
    lengths = torch.tensor(c["lengths"], device=DEV)
    pos = O.rel_pos_table(c["T"], cfg.d_model, sd["rel_pe.div_term"])[None].expand(c["B"], -1, -1).contiguous().to(DEV)
    return cfg, sd, x, lengths, pos, golden("submodules")


def test_activations_mask_posenc(subcase):
    cfg, sd, x, lengths, pos, gold = subcase
    assert_close_fp32(cb.Swish()(x), O.swish(x.cpu()), "swish")
    assert_close_fp32(cb.GLU(dim=-1)(x), O.glu(x.cpu(), -1), "glu last")
    assert_close_fp32(cb.GLU(dim=1)(x[:, :40]), O.glu(x[:, :40].cpu(), 1), "glu dim1")
    valid = cb.generate_padding_mask(lengths)
    assert valid.dtype == torch.bool and torch.equal(valid.cpu(), gold["padding_mask"])
    assert torch.equal(cb.generate_padding_mask(lengths, 50).cpu()[:, :41], gold["padding_mask"])
    pe = _load(cb.RelativePositionalEncoding(cfg.d_model), sub(sd, "rel_pe."))
    out = pe(x)
    assert out.shape == (x.shape[0], 2 * x.shape[1] - 1, cfg.d_model)
    assert_close_fp32(out[0], gold["rel_pe"], "rel_pe")
    assert torch.equal(out[0], out[-1])


def test_ffn_fp32(subcase):
    cfg, sd, x, _, _, gold = subcase
    m = _load(cb.FeedForwardModule(cfg.d_model), sub(sd, "layers.0.ffn_1."))
    assert_close_fp32(m(x), gold["ffn"], "ffn")


@pytest.mark.parametrize("masked", [False, True])
def test_mhsa_fp32(subcase, masked):
    cfg, sd, x, lengths, pos, gold = subcase
    m = _load(cb.MultiHeadSelfAttentionModule(cfg.d_model, cfg.n_heads), sub(sd, "layers.0.attention."))
    mask = (~cb.generate_padding_mask(lengths))[:, None, None, :] if masked else None
    assert_close_fp32(m(x, pos, mask), gold["mhsa_mask" if masked else "mhsa_nomask"], "mhsa")


def test_rel_attention_fp32_vs_oracle(subcase):
    cfg, sd, x, lengths, pos, _ = subcase
    m = _load(cb.RelativeMultiHeadAttention(cfg.d_model, cfg.n_heads), sub(sd, "layers.0.attention.attention."))
    mask = (~cb.generate_padding_mask(lengths))[:, None, None, :]
    want = O.rel_attention(sd, "layers.0.attention.attention.", x.cpu(), pos[0].cpu(), mask[:, 0, 0].cpu(), cfg.n_heads)
    assert_close_fp32(m(x, x, x, pos, mask), want, "rel_attention")


def test_conv_module_fp32(subcase):
    cfg, sd, x, _, _, gold = subcase
    m = _load(cb.ConvolutionModule(cfg.d_model, cfg.kernel_size), sub(sd, "layers.0.conv."))
    assert_close_fp32(m(x), gold["conv"], "conv")


@pytest.mark.parametrize("masked", [False, True])
def test_block_fp32(subcase, masked):
    cfg, sd, x, lengths, pos, gold = subcase
    m = _load(cb.ConformerBlock(cfg.d_model, cfg.n_heads, cfg.kernel_size), sub(sd, "layers.0."))
    mask = (~cb.generate_padding_mask(lengths))[:, None, None, :] if masked else None
    assert_close_fp32(m(x, pos, mask), gold["block_mask" if masked else "block_nomask"], "block", scale=2.0)


def test_subsampling_fp32(subcase):
    cfg, sd, _, _, _, gold = subcase
    c = SUBMODULE_CFG
    m = _load(cb.ConvolutionSubsampling(cfg.d_model), sub(sd, "downsampling_conv."))
    mel = synth_logmel(c["B"], cfg.n_mel_channels, c["Tin"], seed=77).to(DEV)
    y, l = m(mel, torch.tensor(c["mel_lengths"], device=DEV))
    assert_close_fp32(y, gold["subsampling"], "subsampling")
    assert torch.equal(l.cpu(), gold["subsampling_lengths"])


def _encoder_case(case):
    cfg = case["cfg"]
    sd = synth_encoder_state_dict(cfg, case["seed"])
    x = synth_logmel(case["B"], cfg.n_mel_channels, case["Tin"], seed=1000 + case["seed"]).to(DEV)
    lengths = synth_lengths(case["B"], case["Tin"]).to(DEV) if case["with_lengths"] else None
    enc = _load(cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size), sd)
    return enc, x, lengths, golden("encoder_" + case["name"])


@pytest.mark.parametrize("case", ENCODER_CASES, ids=[c["name"] for c in ENCODER_CASES])
def test_encoder_fp32_vs_reference_golden(case):
    enc, x, lengths, gold = _encoder_case(case)
    y, out_len = enc(x, lengths)
    # two fp32 implementations with different summation orders: 4x the single-op gate over 16 layers
    assert_close_fp32(y, gold["y"], "encoder fp32", scale=4.0)
    if lengths is not None:
        assert torch.equal(out_len.cpu(), gold["out_lengths"])
    else:
        assert out_len is None


@pytest.mark.parametrize("case", [c for c in ENCODER_CASES if c["cfg"].d_model % 64 == 0],
                         ids=[c["name"] for c in ENCODER_CASES if c["cfg"].d_model % 64 == 0])
def test_encoder_bf16_vs_reference_golden(case):
    enc, x, lengths, gold = _encoder_case(case)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y, out_len = enc(x, lengths)
    assert y.dtype == torch.float32
    r = assert_close_bf16(y, gold["y"], "encoder bf16")
    print("bf16 vs reference fp32:", case["name"], r)
    y2, _ = enc.set_compute("bf16")(x, lengths)
    assert torch.equal(y, y2)


def test_greedy_ctc_tokens_bit_exact_fp32():
    c = TOKEN_CASE
    cfg = c["cfg"]
    gold = golden("tokens")
    sd = synth_encoder_state_dict(cfg, c["seed"], prefix="encoder.")
    sd.update(synth_decoder_state_dict(cfg.d_model, c["vocab"], seed=c["seed"] + 1, peaky=c["peaky"]))
    model = _load(cb.Conformer(c["vocab"], cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size), sd)
    x = synth_logmel(c["B"], cfg.n_mel_channels, c["Tin"], seed=2000 + c["seed"]).to(DEV)
    lengths = synth_lengths(c["B"], c["Tin"]).to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    logits, out_len = model(x, lengths)
    assert torch.equal(out_len.cpu(), gold["out_lengths"])
    toks = O.greedy_ctc_batch(logits.cpu(), 0, c["vocab"] - 1)
    want = [[int(t) for t in row if t >= 0] for row in gold["tokens"]]
    assert toks == want
    # the same through the device decode (no logits.cpu() round trip)
    tok_dev, cnt_dev = cb.greedy_ctc_decode(logits)
    assert [tok_dev[b, :cnt_dev[b]].tolist() for b in range(tok_dev.shape[0])] == want
    safe = gold["margin"] > 1e-3
    assert torch.equal(logits.cpu().argmax(-1)[safe], gold["ids"][safe])


def test_cpu_tensors_and_training_mode_fail_loudly():
    enc = cb.Encoder(80, 1, 64, 4, 7).to(DEV).eval()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 80, 64))
    enc.train()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 80, 64, device=DEV))


def test_reload_weights_repacks():
    cfg = ENCODER_CASES[0]["cfg"]
    enc, x, lengths, gold = _encoder_case(ENCODER_CASES[0])
    enc.load_state_dict(synth_encoder_state_dict(cfg, 999))
    y_other, _ = enc(x, lengths)
    assert (y_other.cpu() - gold["y"]).abs().max() > 1e-2
    enc.load_state_dict(synth_encoder_state_dict(cfg, ENCODER_CASES[0]["seed"]))
    y, _ = enc(x, lengths)
    assert_close_fp32(y, gold["y"], "reloaded", scale=4.0)


@pytest.mark.parametrize("name,cfg,B,Tin", [("C2-M", CONFORMER_M, 8, 1501), ("C3-L", CONFORMER_L, 4, 2001)])
def test_encoder_bf16_full_length_vs_oracle(name, cfg, B, Tin):
    """Full utterance lengths of BASELINE configs 2/3 (reduced batch so the fp32 oracle stays quick);
    oracle = the restated reference run in fp32 on the GPU with TF32 disabled."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = synth_encoder_state_dict(cfg, 7)
    x = synth_logmel(B, cfg.n_mel_channels, Tin, seed=8).to(DEV)
    lengths = synth_lengths(B, Tin).to(DEV)
    enc = _load(cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size), sd)
    y, out_len = enc.set_compute("bf16")(x, lengths)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    want, want_len = O.encoder(sd_dev, x, lengths, cfg.n_heads)
    assert torch.equal(out_len, want_len)
    print(name, assert_close_bf16(y, want, name))


@pytest.mark.parametrize("T", [1, 17, 63, 64, 65, 130, 374])
@pytest.mark.parametrize("masked", [False, True])
def test_rel_attention_bf16_tensor_core_vs_oracle(T, masked):
    """Head size 64 routes to the tensor-core flash kernel (in-tile relative shift); tile edges at 64."""
    d, H, B = 128, 2, 3
    g = torch.Generator().manual_seed(100 + T)
    sd = {k[len("layers.0.attention.attention."):]: v for k, v in
          synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, H, 31), 40).items()
          if k.startswith("layers.0.attention.attention.")}
    # larger biases/weights so that the positional term matters
    sd = {k: (v * 3.0 if "proj.weight" in k and "out" not in k else v) for k, v in sd.items()}
    x = torch.randn(B, T, d, generator=g)
    pos = O.rel_pos_table(T, d)
    lengths = torch.tensor([T, max(1, T // 2), 1])
    key_is_pad = ~O.padding_mask(lengths, T) if masked else None
    want = O.rel_attention(sd, "", x, pos, key_is_pad, H)
    m = _load(cb.RelativeMultiHeadAttention(d, H), sd).set_compute("bf16")
    mask = key_is_pad[:, None, None, :].to(DEV) if masked else None
    xd = x.to(DEV)
    got = m(xd, xd, xd, pos[None].expand(B, -1, -1).to(DEV), mask)
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 2e-2, rel
    m32 = m.set_compute("fp32")
    assert_close_fp32(m32(xd, xd, xd, pos[None].expand(B, -1, -1).to(DEV), mask), want, "fp32 attention", scale=2.0)


@pytest.mark.parametrize("B,Tin", [(1, 7), (2, 64), (3, 101), (2, 303), (5, 1501)])
def test_subsampling_bf16_tensor_core_front_end(B, Tin):
    """channels = 256 in bf16 mode: conv_1 runs as a K=16 tensor-core GEMM over 3x3 patches (inputs rounded to bf16,
    as the reference's autocast conv2d does), conv_2 as the implicit GEMM; odd and even T1, tile tails, tiny inputs."""
    d = 256
    sd = sub(synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, 4, 31), 90), "downsampling_conv.")
    x = synth_logmel(B, 80, Tin, seed=91 + Tin)
    lengths = torch.tensor([Tin] + [max(7, Tin // 2)] * (B - 1))
    want, want_len = O.conv_subsampling(sd, "", x, lengths)
    m = _load(cb.ConvolutionSubsampling(d), sd).set_compute("bf16")
    got, got_len = m(x.to(DEV), lengths.to(DEV))
    assert torch.equal(got_len.cpu(), want_len)
    assert torch.isfinite(got).all()
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 1e-2, rel


@pytest.mark.parametrize("T", [1, 17, 63, 64, 65, 127, 128, 129, 200, 374, 700])
@pytest.mark.parametrize("masked", [False, True])
def test_rel_attention_bf16_factorised_vs_oracle(T, masked):
    """d_model 256 / head size 64 called WITHOUT a positional table uses the module's own sinusoids through the
    factorised kernel (rotated query x absolute key table, no relative shift); the oracle gets the reference's
    [2T-1, d] table.  Tile edges at 64 keys / 128 queries."""
    d, H, B = 256, 4, 3
    g = torch.Generator().manual_seed(300 + T)
    sd = {k[len("layers.0.attention.attention."):]: v for k, v in
          synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, H, 31), 41).items()
          if k.startswith("layers.0.attention.attention.")}
    # larger projections so that both the content and the positional term matter (scores of a few units)
    sd = {k: (v * 3.0 if "proj.weight" in k and "out" not in k else v) for k, v in sd.items()}
    x = torch.randn(B, T, d, generator=g)
    pos = O.rel_pos_table(T, d)
    lengths = torch.tensor([T, max(1, T // 2), 1])
    key_is_pad = ~O.padding_mask(lengths, T) if masked else None
    want = O.rel_attention(sd, "", x, pos, key_is_pad, H)
    m = _load(cb.RelativeMultiHeadAttention(d, H), sd).set_compute("bf16")
    mask = key_is_pad[:, None, None, :].to(DEV) if masked else None
    xd = x.to(DEV)
    from conformer_b200._lib import check, lib
    check(lib().cfm_set_attention_kernel(1))
    try:
        got = m(xd, xd, xd, None, mask)
    finally:
        check(lib().cfm_set_attention_kernel(0))
    assert torch.isfinite(got).all()
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    # the relative-shift kernel on the same inputs (explicit table) is the yardstick
    got_skew = m(xd, xd, xd, pos[None].expand(B, -1, -1).to(DEV), mask)
    rel_skew = ((got_skew.cpu() - want).norm() / want.norm()).item()
    print(f"T={T} factorised rel {rel:.2e}  shift-kernel rel {rel_skew:.2e}")
    assert rel < 2e-2, (rel, rel_skew)


@pytest.mark.parametrize("d,T", [(256, 150), (512, 70), (64, 40)])
def test_conv_module_bf16_vs_oracle(d, T):
    """bf16 ConvolutionModule: GLU fused into the pointwise GEMM epilogue + packed-fp32x2 depthwise kernel
    (d % 64 == 0, k = 31), against the fp32 oracle."""
    B = 3
    sd = sub(synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, 4, 31), 50 + d), "layers.0.conv.")
    g = torch.Generator().manual_seed(d + T)
    x = torch.randn(B, T, d, generator=g)
    want = O.conv_module(sd, "", x)
    m = _load(cb.ConvolutionModule(d, 31), sd).set_compute("bf16")
    got = m(x.to(DEV))
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 1.5e-2, rel


@pytest.mark.parametrize("T", [1, 15, 16, 95, 96, 97, 111, 112, 192, 193, 374, 499])
def test_conv_module_bf16_fused_tile_edges(T):
    """d_model = 256, k = 31 in bf16 mode routes ConvolutionModule to the single fused kernel (LN -> pw1 -> GLU ->
    depthwise+BN+Swish -> pw2): 96-frame tiles with a 15-frame halo, so lengths around 96k and 96k+-15 are the
    edges; the stand-alone module has no residual, the block test below covers the residual epilogue."""
    B, d = 3, 256
    sd = sub(synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, 4, 31), 51), "layers.0.conv.")
    g = torch.Generator().manual_seed(7000 + T)
    x = torch.randn(B, T, d, generator=g) * 1.5 + 0.25
    want = O.conv_module(sd, "", x)
    m = _load(cb.ConvolutionModule(d, 31), sd).set_compute("bf16")
    got = m(x.to(DEV))
    assert torch.isfinite(got).all()
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 1.5e-2, rel
    assert (got.cpu() - want).abs().max() < 0.1


@pytest.mark.parametrize("T", [40, 200])
def test_block_bf16_fused_conv_residual(T):
    """ConformerBlock in bf16 at d_model = 256: fused FFN, tcgen05 attention and the fused conv module writing
    res + conv into the second residual buffer."""
    B, d, H = 2, 256, 4
    sd = sub(synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, H, 31), 52), "layers.0.")
    g = torch.Generator().manual_seed(7100 + T)
    x = torch.randn(B, T, d, generator=g)
    pos = O.rel_pos_table(T, d)
    lengths = torch.tensor([T, max(1, T // 3)])
    pad = ~O.padding_mask(lengths, T)
    want = O.conformer_block(sd, "", x, pos, pad, H)
    m = _load(cb.ConformerBlock(d, H, 31), sd).set_compute("bf16")
    got = m(x.to(DEV), pos[None].expand(B, -1, -1).to(DEV), pad[:, None, None, :].to(DEV))
    print(assert_close_bf16(got, want, "block bf16"))


@pytest.mark.parametrize("rows", [(1, 1), (2, 100), (3, 374), (5, 128), (64, 374), (40, 500), (149, 128), (300, 127)])
def test_ffn_module_bf16_fused_vs_oracle(rows):
    """d_model = 256 in bf16 mode routes FeedForwardModule to the fused LN->GEMM->Swish->GEMM kernel.  More than 148
    row tiles (one per SM) exercises the balanced schedule: left-over tiles are split along the hidden dimension
    across CTAs and summed by the reduction store."""
    B, T = rows
    d = 256
    sd = sub(synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, 4, 31), 60), "layers.0.ffn_1.")
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, d, generator=g) * 1.5 + 0.3
    want = O.feed_forward(sd, "", x)
    m = _load(cb.FeedForwardModule(d), sd).set_compute("bf16")
    got = m(x.to(DEV))
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 1e-2, rel
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("B,T", [(1, 5), (2, 130), (4, 374)])
def test_mhsa_module_bf16_ln_fused_vs_oracle(B, T):
    """d_model = 256 in bf16 mode: LayerNorm is fused into the q|k|v projection kernel (ln_gemm)."""
    d, H = 256, 4
    sd = sub(synth_encoder_state_dict(cb.EncoderConfig(80, 1, d, H, 31), 70), "layers.0.attention.")
    g = torch.Generator().manual_seed(B * 100 + T)
    x = torch.randn(B, T, d, generator=g) * 2.0 - 0.5
    pos = O.rel_pos_table(T, d)
    lengths = torch.tensor([T] + [max(1, T // 2)] * (B - 1))
    pad = ~O.padding_mask(lengths, T)
    want = O.mhsa_module(sd, "", x, pos, pad, H)
    m = _load(cb.MultiHeadSelfAttentionModule(d, H), sd).set_compute("bf16")
    got = m(x.to(DEV), pos[None].expand(B, -1, -1).to(DEV), pad[:, None, None, :].to(DEV))
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 2e-2, rel


@pytest.mark.timeout(240)
def test_repeated_two_stream_forwards_do_not_hang_or_drift():
    """Regression for a barrier-protocol deadlock seen once per ~1e5 attention CTAs (softmax warps lapping the MMA warp):
    many Conformer-M forwards on two streams must all complete and give the same result (up to the order of the L2
    reduction adds of the residual stream)."""
    cfg = cb.EncoderConfig(80, 4, 256, 4, 31)
    enc = _load(cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size),
                synth_encoder_state_dict(cfg, 5)).set_compute("bf16")
    from conformer_b200.synth import synth_logmel
    x = synth_logmel(32, 80, 1501, seed=9).to(DEV)
    lengths = torch.full((32,), 1501, dtype=torch.int64, device=DEV)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ref = None
    with torch.no_grad():
        for i in range(600):
            with torch.cuda.stream(streams[i & 1]):
                y, _ = enc(x, lengths)
            if i % 100 == 99:
                torch.cuda.synchronize()
                assert torch.isfinite(y).all()
                s = y.float().abs().sum().item()
                ref = s if ref is None else ref
                assert abs(s - ref) <= 1e-3 * abs(ref), (i, s, ref)
    torch.cuda.synchronize()


def _tokens_to_lists(tokens, counts):
    tokens, counts = tokens.cpu(), counts.cpu()
    for b in range(tokens.shape[0]):
        assert (tokens[b, counts[b]:] == -1).all()
    return [tokens[b, :counts[b]].tolist() for b in range(tokens.shape[0])]


@pytest.mark.parametrize("B,T,V", [(1, 1, 5), (3, 33, 7), (4, 374, 370), (2, 1499, 370), (2, 64, 33), (5, 100, 2)])
def test_greedy_ctc_decode_device_bit_exact(B, T, V):
    """cfm_ctc_greedy_decode vs the oracle's restatement of processor.py:301-328: coarse logits force many argmax ties
    (first maximum must win), long runs of pad / unk (skipped without resetting the previous id) and repeats."""
    g = torch.Generator().manual_seed(1000 * B + T + V)
    logits = torch.randint(0, 4, (B, T, V), generator=g).float()               # heavy ties
    hot = torch.randint(0, V, (B, T), generator=g)
    runs = torch.repeat_interleave(hot[:, ::3], 3, dim=1)[:, :T]                # repeated ids in runs of 3
    boost = (torch.rand(B, T, generator=g) < 0.7).float() * 5.0
    logits.scatter_add_(2, runs[..., None], boost[..., None])
    got = _tokens_to_lists(*cb.greedy_ctc_decode(logits.to(DEV)))
    want = O.greedy_ctc_batch(logits, 0, V - 1)
    assert got == want
    # explicit pad / unk ids
    got2 = _tokens_to_lists(*cb.greedy_ctc_decode(logits.to(DEV), pad_id=1, unk_id=0))
    assert got2 == O.greedy_ctc_batch(logits, 1, 0)


def test_greedy_ctc_decode_device_matches_reference_golden_tokens():
    """The reference Conformer's logits (golden fixture) decoded on the device give the reference's token ids."""
    c = TOKEN_CASE
    gold = golden("tokens")
    logits = gold["logits"]
    got = _tokens_to_lists(*cb.greedy_ctc_decode(logits.to(DEV)))
    assert got == O.greedy_ctc_batch(logits, 0, c["vocab"] - 1)
    assert got == [[int(t) for t in row if t >= 0] for row in gold["tokens"]]


def test_encoder_cuda_graph_replay_matches_eager():
    """Encoder.capture: the forward captured in a CUDA graph gives the eager result, also after new inputs are copied into
    its static buffers (fp32 path: deterministic kernels, so equality is exact up to nothing)."""
    cfg = cb.EncoderConfig(80, 2, 256, 4, 31)
    enc = _load(cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size),
                synth_encoder_state_dict(cfg, 11))
    from conformer_b200.synth import synth_logmel
    x1 = synth_logmel(3, 80, 301, seed=21).to(DEV)
    x2 = synth_logmel(3, 80, 301, seed=22).to(DEV)
    lengths = torch.tensor([301, 250, 120], device=DEV)
    for mode in ("fp32", "bf16"):
        enc.set_compute(mode)
        cap = enc.capture(x1, lengths)
        for x in (x1, x2, x1):
            cap.x.copy_(x)
            y_g, l_g = cap.replay()
            y_e, l_e = enc(x, lengths)
            torch.cuda.synchronize()
            assert torch.equal(l_g, l_e)
            if mode == "fp32":
                assert torch.allclose(y_g, y_e, rtol=1e-5, atol=1e-6)
            else:                                   # the residual stream is updated by L2 reduction adds: order may differ
                assert ((y_g - y_e).norm() / y_e.norm()).item() < 2e-3


@pytest.mark.timeout(300)
@pytest.mark.parametrize("B,Tin,with_lengths", [(1, 7, False), (200, 99, True), (103, 50, True), (3, 4100, False),
                                                 (5, 1541, True), (37, 389, True), (64, 771, False), (2, 9821, False),
                                                 (149, 261, True), (12, 2053, True)])
def test_encoder_bf16_shape_sweep_vs_fp32_path(B, Tin, with_lengths):
    """Every shape-dependent schedule (FFN tile sharing for any tile count, attention ring / S-buffer rotation for 1..38
    key tiles, conv-module tile edges, front-end odd/even lengths): the bf16 tensor-core path stays within 2e-2 of the
    fp32 CUDA-core path, which the tests above pin to the reference."""
    cfg = cb.EncoderConfig(80, 3, 256, 4, 31)
    enc = _load(cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size),
                synth_encoder_state_dict(cfg, 3))
    from conformer_b200.synth import synth_logmel
    x = synth_logmel(B, 80, Tin, seed=B + Tin).to(DEV)
    lengths = None
    if with_lengths:
        g = torch.Generator().manual_seed(Tin)
        ls = torch.randint(max(7, Tin // 4), Tin + 1, (B,), generator=g).sort(descending=True).values
        ls[0] = Tin
        lengths = ls.to(DEV)
    y32, l32 = enc.set_compute("fp32")(x, lengths)
    y16, l16 = enc.set_compute("bf16")(x, lengths)
    torch.cuda.synchronize()
    assert torch.isfinite(y16).all()
    if lengths is not None:
        assert torch.equal(l32, l16)
    rel = ((y16 - y32).norm() / y32.norm()).item()
    assert rel < 2e-2, rel


def test_forward_host_pipelined_matches_forward():
    """cfm_encoder_forward_host_pipelined: batches rotating over 3 slots / 2 compute streams with copies on their own
    streams land the same results on the host as the device-resident forward of the same inputs (fp32 path)."""
    cfg = cb.EncoderConfig(80, 2, 256, 4, 31)
    enc = _load(cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size),
                synth_encoder_state_dict(cfg, 17))
    from conformer_b200.synth import synth_logmel
    B, Tin, slots = 4, 401, 3
    xs = [synth_logmel(B, 80, Tin, seed=50 + i).pin_memory() for i in range(7)]
    lens = torch.tensor([401, 333, 200, 77]).pin_memory()
    T2 = enc.output_frames(Tin)
    y_hosts = [torch.empty(B, T2, 256).pin_memory() for _ in range(slots)]
    l_hosts = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(slots)]
    compute = [torch.cuda.Stream(), torch.cuda.Stream()]
    cin, cout = torch.cuda.Stream(), torch.cuda.Stream()
    want = [enc(x.to(DEV), lens.to(DEV)) for x in xs]
    torch.cuda.synchronize()
    got = []
    for k, x in enumerate(xs):
        if k >= slots:                       # the slot's previous result must be consumed before it is overwritten
            cout.synchronize()
            got.append((y_hosts[k % slots].clone(), l_hosts[k % slots].clone())) if len(got) == k - slots else None
        with torch.cuda.stream(compute[k % 2]):
            enc.forward_host_pipelined(x, lens, y_hosts[k % slots], l_hosts[k % slots], k % slots, cin, cout)
    torch.cuda.synchronize()
    # results still sitting in the slots: the last `slots` batches
    for k in range(len(xs) - slots, len(xs)):
        got.append((y_hosts[k % slots].clone(), l_hosts[k % slots].clone()))
    assert len(got) == len(xs)
    for (y_g, l_g), (y_w, l_w) in zip(got, want):
        assert torch.equal(l_g, l_w.cpu())
        assert torch.allclose(y_g, y_w.cpu(), rtol=1e-5, atol=1e-6)
