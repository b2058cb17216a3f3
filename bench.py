#!/usr/bin/env python
"""Headline benchmark: Conformer-M encoder forward, audio-seconds per second (RTFx).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3|C4|C5] [--pdl 0|1]

One "step" = one encoder forward over one synthetic batch of the named workload (BASELINE.json
configs[1] = "C2": Conformer-M, d=256, 16 blocks, 4 heads, bf16, 64 x 15 s of 80-mel log-fbank,
lengths all 1501 so the padding-mask path is inside the timed forward).
For N > 1 launch under torchrun (one rank per GPU); each rank runs the same per-GPU batch
(weak scaling, utterances are independent - no collective on the data path), and the time is the
max over ranks of a CUDA-event interval bracketed by barriers.

Printed JSON (rank 0, one line):
  value                whole-job audio-s/s with the mel batch already resident in HBM (K steps / total time)
  ms_per_step_median   median over the K steps of the interval between consecutive step completions
  e2e                  the same through the host-buffer entry point (cfm_encoder_forward_host_io): pinned HOST mel in,
                       HOST result out, H2D/D2H copies inside the timed region
  roofline             the dominant single kernel, timed per launch with CUDA events on the launching stream in a
                       separate profiled pass of the same steps; `kernels` lists all classes; `roofline_second` is the
                       runner-up (at C2 the fused FFN and the attention kernel are within 2 % of each other)
  gpu_eager_baseline   the reference's own modules (baseline/_ref when it travelled, else the oracle port) run EAGERLY
                       on this GPU under torch.autocast(bf16) at the same batch: the cuBLAS/cuDNN bar
  configs              the other BASELINE configs (C1 fp32, C3, C5) at their own shape: ms_per_step, RTFx, dominant kernel
  strong               strong scaling: the named batch split over the N ranks
  next_rows            (N = 1) the rows either side of the encoder: log-mel front-end and CTC head + greedy
                       decode at the shapes they have around this batch, next to the same computation in eager PyTorch
  cpu_baseline         the reference's CPU path on this box's host cores, bounded sample
  config.pdl           programmatic dependent launch (cfm_set_pdl): without --pdl the multi-stream legs (value, e2e,
                       configs.*.ms_per_step) run with it off and the single-stream / graph-replay legs with it on - what a
                       caller tuning for throughput resp. latency would set
  --impl reference     times that CPU path alone, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from conformer_b200.synth import (CONFORMER_L, CONFORMER_M, CONFORMER_S, EncoderConfig, encoder_flops,  # noqa: E402
                                  synth_encoder_state_dict, synth_lengths, synth_logmel, subsampled_frames)

WORKLOADS = {
    # name: (cfg, batch per GPU, mel frames, lengths: "full" | "var", compute, description)
    "C1": (CONFORMER_S, 4, 1001, "full", "fp32", "Conformer-S encoder fwd, 4x10s, fp32 (BASELINE configs[0])"),
    "C2": (CONFORMER_M, 64, 1501, "full", "bf16", "Conformer-M encoder fwd, 64x15s, bf16 (BASELINE configs[1])"),
    "C3": (CONFORMER_L, 32, 2001, "var", "bf16", "Conformer-L encoder fwd, 32x20s var-len, bf16 (BASELINE configs[2])"),
    "C5": (CONFORMER_M, 16, 6001, "full", "bf16", "Conformer-M encoder fwd, 16x60s long-form, bf16 (BASELINE configs[4])"),
    # training step: Conformer(370, M) fwd + CTC + bwd + Adam, 32 x 15 s per GPU, gradient all-reduce over NCCL
    "C4": (CONFORMER_M, 32, 1501, "full", "bf16", "Conformer-M CTC training step, 32x15s per GPU, bf16, DP all-reduce (BASELINE configs[3])"),
    # tiny shape used only by the CPU multi-process tests of the harness itself
    "T0": (EncoderConfig(80, 2, 64, 4, 31), 2, 101, "var", "fp32", "harness self-test (not a benchmark workload)"),
}
FRAME_SEC = 0.01   # hop 160 @ 16 kHz (train.py:105,109)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t_begin = index, [], None, None

    def start(self):
        """Spawn the sampler (before the warm-up: nvidia-smi needs a few hundred ms to produce its first row)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def begin_timed(self):
        """Call right before the first timed region: waits (bounded) for the sampler's first row."""
        if self.proc is not None:
            t0 = time.monotonic()
            while not self.rows and time.monotonic() - t0 < 3.0:
                time.sleep(0.01)
        self.t_begin = time.monotonic()

    def stop(self):
        """Call right after the last timed region; only rows sampled inside the timed regions are reported."""
        t_end = time.monotonic()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t_begin is not None and self.t_begin <= t <= t_end + 0.03]
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def rank_seed(rank: int) -> int:
    """Each rank benchmarks its own utterances (the batch shards by utterance)."""
    return 1234 + rank


def max_over_ranks(ms: float, dist, device) -> float:
    """Step time of the job = the slowest rank's device-timed interval."""
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rtfx(world: int, secs_per_rank_step: float, ms_per_step: float) -> float:
    """Weak scaling: every rank processes secs_per_rank_step of audio per step."""
    return world * secs_per_rank_step / (ms_per_step * 1e-3)


def strong_scaling_batches(batch: int, world: int):
    """Strong scaling: the named batch split over the ranks (first ranks take the remainder)."""
    return [batch // world + (1 if r < batch % world else 0) for r in range(world)]


def make_lengths(mode: str, batch: int, t_in: int):
    if mode == "full":
        return torch.full((batch,), t_in, dtype=torch.long)
    return synth_lengths(batch, t_in)


def audio_seconds(lengths, batch, t_in):
    return float(lengths.sum().item() if lengths is not None else batch * t_in) * FRAME_SEC


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as fh:
            return next((l.split(":", 1)[1].strip() for l in fh if l.startswith("model name")), "?")
    except OSError:
        return "?"


# ----------------------------------------------------------------------------------------- the reference itself
def reference_encoder(cfg, sd):
    """The reference's own Encoder (baseline/_ref, installed by baseline/install_ref.py; it does not exist in git and
    /root/reference does not exist on the GPU box) with the seeded weights, or None when it did not travel."""
    if not os.path.isdir(os.path.join(REF_DIR, "model")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    from model.modules.encoder import Encoder as RefEncoder
    enc = RefEncoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size).eval()
    enc.load_state_dict(sd, strict=True)
    return enc


def reference_forward_fn(cfg, sd, device):
    """(callable(x, lengths) -> y, kind): the reference module when available, else the oracle port of it."""
    enc = reference_encoder(cfg, sd)
    if enc is not None:
        enc = enc.to(device)

        def run(x, lengths):
            for m in enc.modules():                       # attention.py:60-61 caches a dtype-specific constant
                if hasattr(m, "mask_value"):
                    m.mask_value = None
            return enc(x, lengths)[0]
        return run, "reference"
    from oracle import conformer_oracle as O
    sd_dev = {k: v.to(device) for k, v in sd.items()}
    return (lambda x, lengths: O.encoder(sd_dev, x, lengths, cfg.n_heads)[0]), "port"


def cpu_reference_rtfx(cfg, t_in, len_mode, budget_s=20.0, sample_batch=4):
    """The reference's CPU path on the host cores: fp32, all threads, bounded sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_encoder_state_dict(cfg, 0)
    run, kind = reference_forward_fn(cfg, sd, "cpu")
    x = synth_logmel(sample_batch, cfg.n_mel_channels, t_in, seed=1234)
    lengths = make_lengths(len_mode, sample_batch, t_in)
    secs = audio_seconds(lengths, sample_batch, t_in)
    times = []
    with torch.inference_mode():
        t0 = time.perf_counter()
        run(x, lengths)                                          # warm-up
        first = time.perf_counter() - t0
        reps = max(1, min(5, int(budget_s / max(first, 1e-3)) - 1))
        for _ in range(reps):
            t0 = time.perf_counter()
            run(x, lengths)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times) if times else first
    what = "reference Encoder (baseline/_ref)" if kind == "reference" else "oracle port of the reference Encoder"
    return dict(value=secs / med, unit="audio-s/s", cores=torch.get_num_threads(), kind=kind,
                sample=f"{what}, fp32, batch {sample_batch} x {t_in} frames, median of {max(len(times), 1)} "
                       f"after 1 warm-up, {med * 1e3:.0f} ms/forward, cpu '{cpu_model_name()}'"), med


def run_reference_arm(args, cfg, batch, t_in, len_mode, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, _ = cpu_reference_rtfx(cfg, t_in, len_mode, budget_s=15.0)
    sd = synth_encoder_state_dict(cfg, 0)
    run, kind = reference_forward_fn(cfg, sd, "cpu")
    sb = 4
    x = synth_logmel(sb, cfg.n_mel_channels, t_in, seed=1234)
    lengths = make_lengths(len_mode, sb, t_in)
    secs = audio_seconds(lengths, sb, t_in)
    steps = min(args.steps, 5)
    with torch.inference_mode():
        for _ in range(min(args.warmup, 1)):
            run(x, lengths)
        t0 = time.perf_counter()
        for _ in range(steps):
            run(x, lengths)
        dt = (time.perf_counter() - t0) / steps
    v = secs / dt
    base["value"] = v
    line = {"impl": "reference", "metric": "encoder audio-seconds per second (RTFx)", "value": v, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "sample": f"batch {sb} of the same utterance length per step (throughput metric)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(cfg, batch, t_in, len_mode, dev, steps=5, warmup=2, seed=1234):
    """The reference modules run eagerly on this GPU under bf16 autocast at the SAME batch (cuBLAS / cuDNN / ATen
    sm_100 kernels): the kernel-for-kernel bar.  CUDA-event timed, median of `steps`."""
    sd = synth_encoder_state_dict(cfg, 0)
    run, kind = reference_forward_fn(cfg, sd, dev)
    x = synth_logmel(batch, cfg.n_mel_channels, t_in, seed=seed).to(dev)
    lengths = make_lengths(len_mode, batch, t_in).to(dev)
    secs = audio_seconds(lengths, batch, t_in)
    out = {}
    for name, ctx in (("bf16_autocast", lambda: torch.autocast("cuda", dtype=torch.bfloat16)),
                      ("fp32_tf32_default", lambda: torch.autocast("cuda", enabled=False))):
        times = []
        try:
            with torch.inference_mode(), ctx():
                for _ in range(warmup):
                    run(x, lengths)
                torch.cuda.synchronize()
                for _ in range(steps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run(x, lengths)
                    e1.record()
                    e1.synchronize()
                    times.append(e0.elapsed_time(e1))
        except Exception as e:  # noqa: BLE001 - a reported baseline must never kill the bench line
            out[name] = {"error": str(e)[:200]}
            continue
        ms = statistics.median(times)
        out[name] = {"ms_per_step": ms, "value": secs / (ms * 1e-3), "unit": "audio-s/s", "steps": steps}
        torch.cuda.empty_cache()
    out["kind"] = kind
    out["what"] = (("reference Encoder from baseline/_ref" if kind == "reference" else "oracle port of the reference Encoder")
                   + f", eager PyTorch {torch.__version__} on this GPU, batch {batch} x {t_in} frames, median of {steps} "
                   "CUDA-event-timed forwards after warm-up")
    return out


def _time_ms(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def next_rows_bench(cb, dev, batch=64, seconds=15.0):
    """SURVEY.md 8f rows either side of the encoder, at the shapes they have around the C2 batch (informational, rank 0 at
    N = 1): the log-mel front-end on 64 x 15 s of 16 kHz audio and the CTC head (LSTM 256 -> 640, vocabulary 370, greedy
    decode) on the encoder's 64 x 374 x 256 output, each next to the same computation in eager PyTorch on this GPU."""
    out = {}
    try:
        fe = cb.LogMelFrontEnd()
        wav = torch.randn(batch, int(seconds * 16000), device=dev) * 0.1
        ms = _time_ms(lambda: fe.mel_spectrogram(wav))
        win = torch.hann_window(400, periodic=True, device=dev)
        fb = fe.fb.to(dev)

        def eager():
            spec = torch.stft(wav, 400, 160, 400, win, center=True, pad_mode="reflect", return_complex=True).abs() ** 2
            return torch.log(torch.clamp(torch.matmul(fb.T, spec), min=1e-5))

        out["logmel_frontend"] = {"ms_per_batch": ms, "value": batch * seconds / (ms * 1e-3), "unit": "audio-s/s",
                                  "eager_torch_ms": _time_ms(eager), "what": f"{batch} x {seconds:g} s waveforms -> log-mel "
                                  "(cfm_logmel_forward: 3xTF32 tcgen05 DFT + filterbank GEMMs); eager = torch.stft (cuFFT) + matmul"}
        del wav
    except Exception as e:  # noqa: BLE001 - informational
        out["logmel_frontend"] = {"error": str(e)[:200]}
    try:
        from conformer_b200.synth import synth_decoder_state_dict
        d, V, H, T = 256, VOCAB, 640, 374
        sd = {k[len("decoder."):]: v for k, v in synth_decoder_state_dict(d, V, hidden=H, seed=3).items()}
        dec = cb.Decoder(V, d, H, 1)
        dec.load_state_dict(sd, strict=True)
        dec = dec.to(dev).eval()
        x = torch.randn(batch, T, d, device=dev)
        lengths = torch.full((batch,), T, dtype=torch.long, device=dev)
        row = {}
        for mode in ("fp32", "bf16"):
            dec.set_compute(mode)
            row[mode + "_ms"] = _time_ms(lambda: cb.greedy_ctc_decode(dec(x, lengths)))
        ref = None
        if os.path.isdir(os.path.join(REF_DIR, "model")):
            if REF_DIR not in sys.path:
                sys.path.insert(0, REF_DIR)
            from model.modules.decoder import Decoder as RefDecoder
            ref = RefDecoder(V, d, H, 1)
            ref.load_state_dict(sd)
            ref = ref.to(dev).eval()
            row["eager_reference_ms"] = _time_ms(lambda: ref(x, lengths).argmax(-1))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                row["eager_reference_bf16_autocast_ms"] = _time_ms(lambda: ref(x, lengths).argmax(-1))
        row["what"] = (f"CTC head on {batch} x {T} x {d}: LSTM({d} -> {H}) + Swish + BatchNorm + Linear({H} -> {V}) + greedy decode "
                       "(cfm_decoder_forward + cfm_ctc_greedy_decode); eager = the reference Decoder (cuDNN LSTM) + argmax")
        out["ctc_head"] = row
    except Exception as e:  # noqa: BLE001 - informational
        out["ctc_head"] = {"error": str(e)[:200]}
    return out


# ----------------------------------------------------------------------------------------- our arm
class Bench:
    """One workload on this rank's GPU: builds the encoder and the inputs, times device-resident steps."""

    def __init__(self, cb, cfg, batch, t_in, len_mode, compute, dev, rank, dist):
        self.cfg, self.batch, self.t_in, self.dev, self.dist = cfg, batch, t_in, dev, dist
        sd = synth_encoder_state_dict(cfg, 0)
        enc = cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size)
        enc.load_state_dict(sd)
        self.enc = enc.to(dev).eval().set_compute(compute)
        self.x_host = synth_logmel(batch, cfg.n_mel_channels, t_in, seed=rank_seed(rank)).pin_memory()
        self.lengths_host = make_lengths(len_mode, batch, t_in).pin_memory()
        self.x = self.x_host.to(dev)
        self.lengths = self.lengths_host.to(dev)
        self.t_out = subsampled_frames(t_in)
        self.secs = audio_seconds(self.lengths_host, batch, t_in)
        self.n_streams = int(os.environ.get("BENCH_STREAMS", "2"))
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.n_streams)]
        self.count = 0

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def step(self):
        self.enc(self.x, self.lengths)

    def step_rotating(self):
        i = self.count % self.n_streams
        self.count += 1
        with torch.cuda.stream(self.streams[i]):
            self.step()
        return i

    def timed_serial(self, fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return max_over_ranks(e0.elapsed_time(e1), self.dist, self.dev) / steps

    def timed_rotating(self, steps, extra_streams=(), fn=None):
        """K steps issued alternately on the compute streams (independent batches; the other stream's kernels fill the
        SMs one forward leaves idle in partial waves and at kernel boundaries).  Returns (total/K, median of the
        intervals between consecutive step completions), both max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ends = []
        self.barrier()
        cur = torch.cuda.current_stream()
        e0.record(cur)
        every = list(self.streams) + list(extra_streams)
        for st in every:
            st.wait_event(e0)
        for _ in range(steps):
            st = (fn or self.step_rotating)()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(st if not isinstance(st, int) else self.streams[st])
            ends.append(ev)
        for st in every:
            cur.wait_stream(st)
        e1.record(cur)
        self.barrier()
        # step k runs on stream k % n: the interval between consecutive completions ON THE SAME STREAM covers n steps of
        # the whole pipeline, so interval / n is a per-step time that does not depend on how completions of different
        # streams bunch together
        t = [e0.elapsed_time(ev) for ev in ends]
        n = max(1, len(self.streams)) if fn is None else max(1, getattr(self, "e2e_lanes", len(self.streams)))
        deltas = [(t[k] - t[k - n]) / n for k in range(n, len(t))]
        med = statistics.median(deltas) if deltas else e0.elapsed_time(e1) / steps
        return (max_over_ranks(e0.elapsed_time(e1), self.dist, self.dev) / steps, max_over_ranks(med, self.dist, self.dev))

    def profile(self, steps):
        self.enc.profile(True)
        for _ in range(steps):
            self.step()
        torch.cuda.synchronize()
        prof = self.enc.collect_profile()
        self.enc.profile(False)
        return prof


TENSOR_CLASSES = ("gemm_tcgen05", "attention", "gemm_fp32", "ffn_fused", "conv_fused")
SINGLE_KERNEL_CLASSES = ("ffn_fused", "attention", "conv_fused", "conv1", "dwconv")


def kernel_table(prof, prof_steps, peaks):
    kernels = {}
    for name, e in prof.items():
        if e["launches"] == 0:
            continue
        ms = e["ms"] / prof_steps
        k = {"launches_per_step": e["launches"] / prof_steps, "ms_per_step": ms}
        if e["flops"] > 0:
            k["tflops"] = e["flops"] / prof_steps / (ms * 1e-3) / 1e12
        if e["bytes"] > 0:
            k["gbs"] = e["bytes"] / prof_steps / (ms * 1e-3) / 1e9
        kernels[name] = k
    total = sum(k["ms_per_step"] for k in kernels.values())
    for n, k in kernels.items():
        k["share"] = k["ms_per_step"] / total if total else None
        if "tflops" in k and n in TENSOR_CLASSES and n != "gemm_fp32":
            k["tensor_frac"] = k["tflops"] / peaks["tensor"]
        if "gbs" in k:
            k["hbm_frac"] = k["gbs"] / peaks["hbm"]
    return kernels


def dominant_roofline(kernels, peaks, rank=0):
    """The roofline block describes ONE kernel (so does its ncu traffic figure): the most expensive class that is a single
    kernel function; "gemm_tcgen05" aggregates several GEMM shapes and stays in `kernels`.  rank = 1: the runner-up (at C2
    the fused FFN and the attention kernel are within 2 % of each other, so which one is "dominant" changes from run to run)."""
    single = [n for n in kernels if n in SINGLE_KERNEL_CLASSES]
    order = sorted(single or kernels, key=lambda n: -kernels[n]["ms_per_step"])
    dom_name = order[min(rank, len(order) - 1)]
    dom = kernels[dom_name]
    if "tflops" in dom and dom_name in TENSOR_CLASSES:
        roof = {"kernel": dom_name, "bound": "tensor", "achieved": dom["tflops"], "peak": peaks["tensor"],
                "unit": "TFLOP/s", "frac": dom["tflops"] / peaks["tensor"], "traffic": None,
                "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step)"}
    else:
        roof = {"kernel": dom_name, "bound": "hbm", "achieved": dom.get("gbs"), "peak": peaks["hbm"], "unit": "GB/s",
                "frac": (dom.get("gbs") or 0.0) / peaks["hbm"], "traffic": None, "peak_source": f"{peaks['src']} hbm_gbs"}
    # roofline.traffic: DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        names = {"ffn_fused": ["ffn_fused_kernel"], "attention": ["rel_attention_fact_kernel", "rel_attention_tc_kernel"],
                 "conv_fused": ["conv_module_fused_kernel"], "gemm_tcgen05": ["gemm_tc_kernel"], "conv1": ["conv1_tc_kernel"]}
        for n in names.get(dom_name, []):
            hit = [k for k in tr["kernels"] if k.startswith(n)]
            if hit:
                roof["traffic"] = tr["kernels"][hit[0]]["dram_bytes_per_launch"]
                roof["traffic_source"] = "profiles/ncu_traffic.json (" + os.path.basename(tr["source"]) + "): " + hit[0]
                break
    except (OSError, ValueError, KeyError):
        pass
    return roof


_PDL_AUTO = True          # no --pdl flag: programmatic dependent launch on for single-stream legs, off for multi-stream legs


def pdl_for(multi_stream):
    """Programmatic dependent launch (cfm_set_pdl) lets a kernel's CTAs become resident while the previous kernel of ITS
    stream drains.  On one stream that hides launch ramps (C2: 4.37 -> 4.06 ms per step); with two streams alternating,
    those parked CTAs hold SMs the other stream's kernels would have used (C3: 7.9 -> 8.4 ms), so a throughput-oriented
    caller turns it off.  bench.py does what such a caller would: off for the multi-stream legs, on for the
    single-stream and graph-replay legs - unless --pdl fixes it."""
    if _PDL_AUTO:
        from conformer_b200._lib import lib as _cfm_lib
        _cfm_lib().cfm_set_pdl(0 if multi_stream else 1)


def side_config(cb, name, dev, rank, dist, world, steps, peaks):
    """One of the other BASELINE configs at its own shape: device-resident two-stream time, single-stream time, dominant
    kernel.  Fewer steps than the headline; same timing rules."""
    cfg, batch, t_in, len_mode, compute, desc = WORKLOADS[name]
    b = Bench(cb, cfg, batch, t_in, len_mode, compute, dev, rank, dist)
    with torch.inference_mode():
        for _ in range(3):
            b.step()
        for _ in range(2 * b.n_streams):
            b.step_rotating()
        pdl_for(False)
        ms_serial = b.timed_serial(b.step, steps)
        pdl_for(b.n_streams > 1)
        ms, med = b.timed_rotating(steps)
        pdl_for(False)
        prof = b.profile(min(steps, 3))
    kernels = kernel_table(prof, min(steps, 3), peaks)
    flops = encoder_flops(batch, t_in, cfg.d_model, cfg.n_blocks, cfg.n_mel_channels, cfg.kernel_size)
    roof = dominant_roofline(kernels, peaks)
    out = {"workload": desc, "dtype": "bf16" if compute == "bf16" else "f32", "per_gpu_batch": batch, "mel_frames": t_in,
           "frames_out": b.t_out, "audio_seconds_per_gpu_step": b.secs, "steps": steps, "ms_per_step": ms,
           "ms_per_step_median": med, "serial_ms_per_step": ms_serial, "value": whole_job_rtfx(world, b.secs, ms),
           "unit": "audio-s/s", "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
           "dominant_kernel": {k: roof[k] for k in ("kernel", "bound", "achieved", "unit", "frac")},
           "kernels": {n: {kk: round(v, 4) for kk, v in k.items() if v is not None} for n, k in kernels.items()}}
    if compute == "bf16":
        out["tensor_frac_whole_step"] = out["algorithmic_tflops"] / peaks["tensor"]
    del b
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------- config 4: the training step
VOCAB = 370            # tokenizer/vi.json (train.py:115)


def train_batch(cfg, batch, t_in, seed, dev):
    g = torch.Generator().manual_seed(seed)
    x = synth_logmel(batch, cfg.n_mel_channels, t_in, seed=seed).to(dev)
    xl = torch.full((batch,), t_in, dtype=torch.long, device=dev)
    yl = torch.randint(40, 81, (batch,), generator=g)
    y = torch.randint(1, VOCAB - 1, (batch, 80), generator=g).to(dev)
    return x, xl, y, yl.to(dev)


def ctc_loss_fn(outputs, targets, in_lens, tgt_lens):
    """evaluation.py:10-16"""
    return torch.nn.functional.ctc_loss(outputs.float().log_softmax(-1).transpose(0, 1), targets, in_lens, tgt_lens, blank=0,
                                        zero_infinity=True)


def run_train_bench(args, cfg, batch, t_in, desc):
    """One step = forward (training mode, dropout 0.1) + CTC loss + backward + gradient all-reduce + Adam, as train.py:232-243."""
    import conformer_b200 as cb
    from conformer_b200.synth import synth_decoder_state_dict
    from conformer_b200.training import GradientAllReducer
    if args.pdl is not None:
        from conformer_b200._lib import lib as _cfm_lib
        _cfm_lib().cfm_set_pdl(args.pdl)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    sd = synth_encoder_state_dict(cfg, 0, prefix="encoder.")
    sd.update(synth_decoder_state_dict(cfg.d_model, VOCAB, seed=1))
    model = cb.Conformer(VOCAB, cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, dropout_rate=0.1)
    model.load_state_dict(sd)
    model = model.to(dev).train().set_compute("bf16")
    opt = torch.optim.Adam(model.parameters(), lr=2e-5)
    reducer = GradientAllReducer(model)
    x, xl, y, yl = train_batch(cfg, batch, t_in, rank_seed(rank), dev)
    secs = audio_seconds(xl.cpu(), batch, t_in)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, ol = model(x, xl)
        loss = ctc_loss_fn(out, y, ol, yl)
        opt.zero_grad()
        reducer.arm()
        loss.backward()
        reducer.finish()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        loss = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.begin_timed()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dist, dev) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    loss_v = float(loss.item())
    # forward-only and forward+backward-only splits, and the all-reduce alone (bus bandwidth against NVLink)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    ev[0].record()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out, ol = model(x, xl)
    ev[1].record()
    l2 = ctc_loss_fn(out, y, ol, yl)
    opt.zero_grad()
    l2.backward()
    ev[2].record()
    torch.cuda.synchronize()
    fwd_ms, bwd_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    allreduce = None
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if dist is not None:
        flat = torch.zeros(n_params, device=dev)
        for _ in range(3):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        barrier()
        ev[0].record()
        for _ in range(5):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        ev[1].record()
        barrier()
        t = max_over_ranks(ev[0].elapsed_time(ev[1]), dist, dev) / 5
        allreduce = {"bytes_per_step": n_params * 4, "ms_alone": t, "bus_gbs": 2 * (world - 1) / world * n_params * 4 / (t * 1e-3) / 1e9,
                     "nvlink_ref_gbs": 725.0, "overlap": "per-block slices of the flat fp32 gradient buffer are all-reduced "
                     "(NCCL, async) while the backward of the earlier blocks runs; exposed cost = step time at N minus step time at 1"}
    eager = None
    if rank == 0 and not args.no_extras:
        eager = gpu_eager_train_baseline(cfg, batch, t_in, dev)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    fl = 3.0 * encoder_flops(batch, t_in, cfg.d_model, cfg.n_blocks, cfg.n_mel_channels, cfg.kernel_size)
    from conformer_b200.training import encoder_train_state
    enc_mod = next(m for m in model.modules() if hasattr(m, "_train_state"))
    graph_stats = encoder_train_state(enc_mod).graph_stats()       # native calls launched directly / captured / replayed
    line = {"metric": "training-step audio-seconds per second", "value": world * secs / (ms * 1e-3), "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": batch, "mel_frames": t_in, "vocab": VOCAB, "dropout": 0.1,
                       "optimizer": "Adam lr 2e-5 (train.py:188), fp32 master weights", "targets": "80 ids per utterance, lengths 40-80",
                       "parallelism": f"dp{world}: one process per GPU, gradient all-reduce (AVG) over NCCL, no SyncBN (as the reference)",
                       "step": "encoder fwd/bwd in the native library (training-mode BatchNorm, dropout), CTC head + CTC loss + Adam in PyTorch (outside the path)"},
            "forward_ms": fwd_ms, "backward_ms": bwd_ms, "final_loss": loss_v,
            "algorithmic_tflops_encoder": fl / (ms * 1e-3) / 1e12, "allreduce": allreduce, "clocks": clocks,
            "cuda_graphs": graph_stats, "gpu_launches": None}
    if eager is not None:
        line["gpu_eager_baseline"] = eager
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def gpu_eager_train_baseline(cfg, batch, t_in, dev, steps=3):
    """The reference Conformer (baseline/_ref) training step eagerly on this GPU under bf16 autocast, same batch."""
    if not os.path.isdir(os.path.join(REF_DIR, "model")):
        return {"unavailable": "baseline/_ref did not travel"}
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        from conformer_b200.synth import synth_decoder_state_dict
        from model.conformer import Conformer as RefConformer
        sd = synth_encoder_state_dict(cfg, 0, prefix="encoder.")
        sd.update(synth_decoder_state_dict(cfg.d_model, VOCAB, seed=1))
        m = RefConformer(VOCAB, cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size, dropout_rate=0.1)
        m.load_state_dict(sd)
        m = m.to(dev).train()
        opt = torch.optim.Adam(m.parameters(), lr=2e-5)
        x, xl, y, yl = train_batch(cfg, batch, t_in, 1234, dev)
        times = []
        for i in range(steps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out, ol = m(x, xl)
            loss = ctc_loss_fn(out, y, ol, yl)
            opt.zero_grad()
            loss.backward()
            opt.step()
            e1.record()
            e1.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        return {"ms_per_step": ms, "value": batch * t_in * FRAME_SEC / (ms * 1e-3), "unit": "audio-s/s", "kind": "reference",
                "what": f"reference Conformer training step, eager PyTorch {torch.__version__} bf16 autocast, batch {batch}, median of {steps}"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--compute", default=None, choices=["bf16", "fp32"])
    ap.add_argument("--pdl", type=int, default=None, choices=[0, 1],
                    help="programmatic dependent launch of the forward kernels (cfm_set_pdl); default: library default (on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_eager_baseline, the other configs and strong scaling")
    args = ap.parse_args()
    cfg, batch, t_in, len_mode, compute, desc = WORKLOADS[args.config]
    if args.batch:
        batch = args.batch
    if args.compute:
        compute = args.compute
    if args.impl == "reference":
        run_reference_arm(args, cfg, batch, t_in, len_mode, desc)
        return
    if args.config == "C4":
        run_train_bench(args, cfg, batch, t_in, desc)
        return

    import conformer_b200 as cb
    global _PDL_AUTO
    if args.pdl is not None:
        from conformer_b200._lib import lib as _cfm_lib
        _cfm_lib().cfm_set_pdl(args.pdl)
        _PDL_AUTO = False
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    peaks = load_peaks()

    b = Bench(cb, cfg, batch, t_in, len_mode, compute, dev, rank, dist)
    enc, t_out = b.enc, b.t_out
    n_dev_streams = b.n_streams
    y_host = torch.empty((batch, t_out, cfg.d_model), dtype=torch.float32).pin_memory()
    lo_host = torch.empty(batch, dtype=torch.int64).pin_memory()
    # end to end: the mel crosses PCIe as bf16 when the bf16 tensor-core front-end takes it (its first kernel rounds the
    # mel to bf16, so the result is bit-identical to an fp32 transfer - tests/test_gpu_parity.py), the result as fp32
    # (the reference's output dtype under autocast).  The fp32-in form is timed next to it.
    bf16_in = compute == "bf16" and cfg.d_model == 256
    x_e2e = b.x_host.bfloat16().pin_memory() if bf16_in else b.x_host

    with torch.inference_mode():
        for _ in range(args.warmup):
            b.step()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(2 * n_dev_streams):
            b.step_rotating()
        step_e2e_serial = lambda: enc.forward_host(x_e2e, b.lengths_host, y_host, lo_host)     # noqa: E731
        for _ in range(2):
            step_e2e_serial()
        if rank == 0:
            sampler.begin_timed()
        pdl_for(False)
        ms_dev_serial = b.timed_serial(b.step, args.steps)
        pdl_for(n_dev_streams > 1)
        ms_dev, ms_dev_median = b.timed_rotating(args.steps)
        pdl_for(False)
        launches = enc.last_launch_count
        # informational: the same forward replayed from a CUDA graph on one stream (Encoder.capture)
        ms_graph_serial = None
        try:
            cap = enc.capture(b.x, b.lengths)
            for _ in range(3):
                cap.replay()
            ms_graph_serial = b.timed_serial(cap.replay, args.steps)
            del cap
        except Exception as e:  # noqa: BLE001 - optional measurement, never fatal
            sys.stderr.write(f"[bench] CUDA-graph replay skipped: {e}\n")
        ms_e2e_serial = b.timed_serial(step_e2e_serial, args.steps)
        # The end-to-end loop through cfm_encoder_forward_host_io: BENCH_E2E_SLOTS (default 4) in-flight batches rotate over
        # the same compute streams; each step's H2D copy goes to a copy-in stream and its D2H copy to a copy-out stream,
        # ordered by the library's per-slot events, so the compute streams never wait for a copy.  Every step still copies
        # its own input from pinned host memory and lands its own result on the host inside the timed region.
        n_slots = int(os.environ.get("BENCH_E2E_SLOTS", "4")) if n_dev_streams > 1 else 1
        copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        y_hosts = [y_host] + [torch.empty_like(y_host).pin_memory() for _ in range(n_slots - 1)]
        lo_hosts = [lo_host] + [torch.empty_like(lo_host).pin_memory() for _ in range(n_slots - 1)]
        e2e_count = [0]

        def make_e2e_step(xh, yhs=None):
            yhs = y_hosts if yhs is None else yhs

            def step():
                k = e2e_count[0]
                e2e_count[0] += 1
                st = b.streams[k % n_dev_streams]
                with torch.cuda.stream(st):
                    enc.forward_host_pipelined(xh, b.lengths_host, yhs[k % n_slots], lo_hosts[k % n_slots], k % n_slots,
                                               copy_in, copy_out)
                return copy_out                       # a step is complete when its result has landed on the host
            return step

        b.e2e_lanes = n_slots                         # a slot's results complete in order; n_slots steps per slot cycle
        step_e2e = make_e2e_step(x_e2e)
        for _ in range(2 * n_slots):
            step_e2e()
        torch.cuda.synchronize()
        pdl_for(n_dev_streams > 1)
        ms_e2e, ms_e2e_median = b.timed_rotating(args.steps, extra_streams=(copy_in, copy_out), fn=step_e2e)
        ms_e2e_fp32in = ms_e2e_bf16out = None
        if bf16_in:
            step_f32 = make_e2e_step(b.x_host)
            for _ in range(n_slots):
                step_f32()
            torch.cuda.synchronize()
            ms_e2e_fp32in, _ = b.timed_rotating(args.steps, extra_streams=(copy_in, copy_out), fn=step_f32)
            # informational: the result also crossing PCIe as bf16 (a caller option of cfm_encoder_forward_host_io; the
            # headline keeps the reference's fp32 result).  Shows how much of e2e is the host link at N GPUs.
            y16 = [torch.empty(y_host.shape, dtype=torch.bfloat16).pin_memory() for _ in range(n_slots)]
            step_16 = make_e2e_step(x_e2e, y16)
            for _ in range(2 * n_slots):
                step_16()
            torch.cuda.synchronize()
            ms_e2e_bf16out, _ = b.timed_rotating(args.steps, extra_streams=(copy_in, copy_out), fn=step_16)
            del y16
        pdl_for(False)
        clocks = sampler.stop() if rank == 0 else None       # sampled across the timed regions above
        # profiled pass (per-launch CUDA events on the launching stream), not part of the headline
        prof_steps = min(args.steps, 5)
        prof = b.profile(prof_steps)

        # strong scaling (SURVEY.md 8e "alongside"): the named batch split over the ranks
        strong = None
        if not args.no_extras:
            mine = strong_scaling_batches(batch, world)[rank]
            if world == 1:
                strong = {"per_rank_batches": [batch], "ms_per_step": ms_dev, "value": whole_job_rtfx(1, b.secs, ms_dev),
                          "unit": "audio-s/s", "note": "N = 1: identical to the weak-scaling line"}
            else:
                sb = Bench(cb, cfg, max(mine, 1), t_in, len_mode, compute, dev, rank, dist)
                for _ in range(3):
                    sb.step()
                for _ in range(2 * sb.n_streams):
                    sb.step_rotating()
                pdl_for(sb.n_streams > 1)
                ms_s, med_s = sb.timed_rotating(args.steps)
                pdl_for(False)
                total_secs = audio_seconds(make_lengths(len_mode, batch, t_in), batch, t_in)
                strong = {"per_rank_batches": strong_scaling_batches(batch, world), "ms_per_step": ms_s,
                          "ms_per_step_median": med_s, "value": total_secs / (ms_s * 1e-3), "unit": "audio-s/s",
                          "note": f"the {batch}-utterance batch split over {world} ranks; max-over-ranks device time"}
                del sb
                torch.cuda.empty_cache()

    # the other BASELINE configs at their own shape, and the eager-PyTorch bar, on rank 0's GPU after the headline numbers
    # (every rank runs them so the barriers match; only rank 0 reports)
    configs, eager, next_rows = {}, None, None
    if not args.no_extras and args.config == "C2":
        del b
        torch.cuda.empty_cache()
        for name in ("C1", "C3", "C5"):
            try:
                configs[name] = side_config(cb, name, dev, rank, dist, world, max(5, args.steps // 2), peaks)
            except Exception as e:  # noqa: BLE001
                configs[name] = {"error": str(e)[:300]}
        if rank == 0:
            eager = gpu_eager_baseline(cfg, batch, t_in, len_mode, dev)
            if world == 1:
                next_rows = next_rows_bench(cb, dev)

    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    kernels = kernel_table(prof, prof_steps, peaks)
    roof = dominant_roofline(kernels, peaks)
    flops = encoder_flops(batch, t_in, cfg.d_model, cfg.n_blocks, cfg.n_mel_channels, cfg.kernel_size)
    secs_per_step = audio_seconds(make_lengths(len_mode, batch, t_in), batch, t_in)
    h2d = x_e2e.numel() * x_e2e.element_size() + batch * 8
    d2h = y_host.numel() * 4 + batch * 8
    line = {
        "metric": "encoder audio-seconds per second (RTFx)", "value": whole_job_rtfx(world, secs_per_step, ms_dev),
        "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
        "ms_per_step_median": ms_dev_median, "serial_ms_per_step": ms_dev_serial,
        "serial_cuda_graph_ms_per_step": ms_graph_serial, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if compute == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": desc, "per_gpu_batch": batch, "mel_frames": t_in, "frames_out": t_out,
                   "lengths": "all = mel_frames (padding-mask path inside the timed forward)" if len_mode == "full"
                              else "sorted descending, linspace(Tin, Tin/2)",
                   "audio_seconds_per_gpu_step": secs_per_step, "parallelism": f"dp{world} (independent replicas)",
                   "issue": f"steps rotate over {n_dev_streams} CUDA streams per GPU (independent batches overlap at kernel "
                            "boundaries); serial_ms_per_step is the single-stream time; ms_per_step = total / steps, "
                            "ms_per_step_median = median interval between consecutive step completions",
                   "pdl": ("programmatic dependent launch (cfm_set_pdl): off for the multi-stream legs (value, e2e, configs.*.ms_per_step), "
                           "on for the single-stream and graph-replay legs - what a caller tuning for throughput / latency sets")
                          if _PDL_AUTO else f"programmatic dependent launch fixed by --pdl {args.pdl}",
                   "l2": "no flush needed: one step streams ~1.3 GB of activations (conv1 output alone 958 MB) "
                         "through the 126 MB L2"},
        "e2e": {"value": whole_job_rtfx(world, secs_per_step, ms_e2e), "unit": "audio-s/s", "ms_per_step": ms_e2e,
                "ms_per_step_median": ms_e2e_median, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "pcie_gbs_per_rank": {"h2d": h2d / (ms_e2e * 1e-3) / 1e9, "d2h": d2h / (ms_e2e * 1e-3) / 1e9},
                "host_dtypes": {"mel": "bf16" if bf16_in else "f32", "result": "f32"},
                "fp32_mel_ms_per_step": ms_e2e_fp32in, "bf16_result_ms_per_step": ms_e2e_bf16out,
                "serial_ms_per_step": ms_e2e_serial,
                "api": "cfm_encoder_forward_host_io (Encoder.forward_host_pipelined): pinned host mel -> device -> forward -> "
                       f"host; {n_slots} in-flight slots over {n_dev_streams} compute streams, copies on a copy-in and a "
                       "copy-out stream ordered by the library's per-slot events; serial_ms_per_step is the same call "
                       "(copy, forward, copy) on a single stream",
                "slots": n_slots, "compute_streams": n_dev_streams},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "algorithmic_tflops": flops / (ms_dev * 1e-3) / 1e12,
        "tensor_frac_whole_step": flops / (ms_dev * 1e-3) / 1e12 / peaks["tensor"],
        "roofline": roof,
        "roofline_second": dominant_roofline(kernels, peaks, rank=1),
        "kernels": kernels,
    }
    if strong is not None:
        line["strong"] = strong
    if configs:
        line["configs"] = configs
    if eager is not None:
        line["gpu_eager_baseline"] = eager
        ours = ms_dev_serial
        bf = eager.get("bf16_autocast", {})
        if "ms_per_step" in bf:
            eager["speedup_vs_bf16_autocast"] = {"single_stream": bf["ms_per_step"] / ours, "two_streams": bf["ms_per_step"] / ms_dev}
    if next_rows is not None:
        line["next_rows"] = next_rows
    if not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference_rtfx(cfg, t_in, len_mode)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
