#!/usr/bin/env python
"""Headline benchmark: Conformer-M encoder forward, audio-seconds per second (RTFx).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C5]

One "step" = one encoder forward over one synthetic batch of the named workload (BASELINE.json
configs[1] = "C2": Conformer-M, d=256, 16 blocks, 4 heads, bf16, 64 x 15 s of 80-mel log-fbank).
For N > 1 launch under torchrun (one rank per GPU); each rank runs the same per-GPU batch
(weak scaling, utterances are independent - no collective on the data path), and the time is the
max over ranks of a CUDA-event interval bracketed by barriers.

Printed JSON (rank 0, one line):
  value          whole-job audio-s/s with the mel batch already resident in HBM
  e2e            the same through cfm_encoder_forward_host: pinned HOST mel in, HOST result out,
                 H2D/D2H copies inside the timed region
  roofline       the dominant kernel class (tcgen05 GEMMs), timed per launch with CUDA events on the
                 launching stream in a separate profiled pass of the same steps; `kernels` lists all
  cpu_baseline   the oracle (a port of the reference's algorithm; the reference itself is a Python
                 package that does not exist on the GPU box) on this box's host cores, bounded sample
  --impl reference  times that CPU path alone, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from conformer_b200.synth import (CONFORMER_L, CONFORMER_M, EncoderConfig, synth_encoder_state_dict,  # noqa: E402
                                  synth_lengths, synth_logmel, subsampled_frames)

WORKLOADS = {
    # name: (cfg, batch per GPU, mel frames, variable lengths, description)
    "C2": (CONFORMER_M, 64, 1501, False, "Conformer-M encoder fwd, 64x15s, bf16 (BASELINE configs[1])"),
    "C3": (CONFORMER_L, 32, 2001, True, "Conformer-L encoder fwd, 32x20s var-len, bf16 (BASELINE configs[2])"),
    "C5": (CONFORMER_M, 16, 6001, False, "Conformer-M encoder fwd, 16x60s long-form, bf16 (BASELINE configs[4])"),
    # tiny shape used only by the CPU multi-process tests of the harness itself
    "T0": (EncoderConfig(80, 2, 64, 4, 31), 2, 101, True, "harness self-test (not a benchmark workload)"),
}
FRAME_SEC = 0.01   # hop 160 @ 16 kHz (train.py:105,109)


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


def audio_seconds(lengths, batch, t_in):
    return float(lengths.sum().item() if lengths is not None else batch * t_in) * FRAME_SEC


def cpu_reference_rtfx(cfg, t_in, var_len, budget_s=20.0, sample_batch=4):
    """The reference's algorithm (oracle port) on the host cores: fp32, all threads, bounded sample."""
    from oracle import conformer_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_encoder_state_dict(cfg, 0)
    x = synth_logmel(sample_batch, cfg.n_mel_channels, t_in, seed=1234)
    lengths = synth_lengths(sample_batch, t_in) if var_len else None
    secs = audio_seconds(lengths, sample_batch, t_in)
    times = []
    with torch.inference_mode():
        t0 = time.perf_counter()
        O.encoder(sd, x, lengths, cfg.n_heads)                 # warm-up
        first = time.perf_counter() - t0
        reps = max(1, min(5, int(budget_s / max(first, 1e-3)) - 1))
        for _ in range(reps):
            t0 = time.perf_counter()
            O.encoder(sd, x, lengths, cfg.n_heads)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times) if times else first
    model = "?"
    try:
        with open("/proc/cpuinfo") as fh:
            model = next((l.split(":", 1)[1].strip() for l in fh if l.startswith("model name")), "?")
    except OSError:
        pass
    return dict(value=secs / med, unit="audio-s/s", cores=torch.get_num_threads(), kind="port",
                sample=f"oracle encoder fp32, batch {sample_batch} x {t_in} frames, median of {max(len(times), 1)} "
                       f"after 1 warm-up, {med * 1e3:.0f} ms/forward, cpu '{model}'"), med


def run_reference_arm(args, cfg, batch, t_in, var_len, desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, _ = cpu_reference_rtfx(cfg, t_in, var_len, budget_s=15.0)
    # K timed "steps", each one bounded-sample forward
    from oracle import conformer_oracle as O
    sd = synth_encoder_state_dict(cfg, 0)
    sb = 4
    x = synth_logmel(sb, cfg.n_mel_channels, t_in, seed=1234)
    lengths = synth_lengths(sb, t_in) if var_len else None
    secs = audio_seconds(lengths, sb, t_in)
    steps = min(args.steps, 5)
    with torch.inference_mode():
        for _ in range(min(args.warmup, 1)):
            O.encoder(sd, x, lengths, cfg.n_heads)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.encoder(sd, x, lengths, cfg.n_heads)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--compute", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg, batch, t_in, var_len, desc = WORKLOADS[args.config]
    if args.batch:
        batch = args.batch
    if args.impl == "reference":
        run_reference_arm(args, cfg, batch, t_in, var_len, desc)
        return

    import conformer_b200 as cb
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

    sd = synth_encoder_state_dict(cfg, 0)
    enc = cb.Encoder(cfg.n_mel_channels, cfg.n_blocks, cfg.d_model, cfg.n_heads, cfg.kernel_size)
    enc.load_state_dict(sd)
    enc = enc.to(dev).eval().set_compute(args.compute)
    x_host = synth_logmel(batch, cfg.n_mel_channels, t_in, seed=rank_seed(rank)).pin_memory()
    lengths_host = synth_lengths(batch, t_in).pin_memory() if var_len else None
    x = x_host.to(dev)
    lengths = lengths_host.to(dev) if var_len else None
    t_out = subsampled_frames(t_in)
    y_host = torch.empty((batch, t_out, cfg.d_model), dtype=torch.float32).pin_memory()
    lo_host = torch.empty(batch, dtype=torch.int64).pin_memory() if var_len else None
    secs_per_step = audio_seconds(lengths_host, batch, t_in)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dist, dev) / steps

    with torch.inference_mode():
        step_dev = lambda: enc(x, lengths)                                              # noqa: E731
        step_e2e = lambda: enc.forward_host(x_host, lengths_host, y_host, lo_host)     # noqa: E731
        for _ in range(args.warmup):
            step_dev()
        n_streams = int(os.environ.get("BENCH_STREAMS", "2"))
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        counter = [0]

        def on_alternating_streams(fn):
            def step():
                i = counter[0] % n_streams
                counter[0] += 1
                with torch.cuda.stream(streams[i]):
                    fn()
            return step

        def timed_two_streams(fn, steps):
            """K steps issued alternately on two streams (independent batches; the second stream's kernels fill the
            SMs the first one leaves idle in partial waves and during kernel ramp-up/tail)."""
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            step = on_alternating_streams(fn)
            barrier()
            cur = torch.cuda.current_stream()
            e0.record(cur)
            for st in streams:
                st.wait_event(e0)
            for _ in range(steps):
                step()
            for st in streams:
                cur.wait_stream(st)
            e1.record(cur)
            barrier()
            return max_over_ranks(e0.elapsed_time(e1), dist, dev) / steps

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        warm = on_alternating_streams(step_dev)
        for _ in range(2 * n_streams):
            warm()
        for _ in range(2):
            step_e2e()
        if rank == 0:
            sampler.begin_timed()
        ms_dev_serial = timed(step_dev, args.steps)
        ms_dev = timed_two_streams(step_dev, args.steps)
        launches = enc.last_launch_count
        # informational: the same forward replayed from a CUDA graph on one stream (Encoder.capture).  Replay removes the
        # per-kernel launch latency a single stream pays (4.30 vs 4.55 ms), but two overlapping graph replays pack worse
        # than two streams of plain launches (4.12 vs 3.86 ms), so the headline loop stays on stream launches.
        ms_graph_serial = None
        try:
            cap = enc.capture(x, lengths)
            for _ in range(3):
                cap.replay()
            ms_graph_serial = timed(cap.replay, args.steps)
            del cap
        except Exception as e:  # noqa: BLE001 - optional measurement, never fatal
            sys.stderr.write(f"[bench] CUDA-graph replay skipped: {e}\n")
        ms_e2e_serial = timed(step_e2e, args.steps)
        # The end-to-end loop through cfm_encoder_forward_host_pipelined: BENCH_E2E_SLOTS (default 4) in-flight batches
        # rotate over the same two compute streams; each step's H2D copy goes to a copy-in stream and its D2H copy to a
        # copy-out stream, ordered by the library's per-slot events, so the compute streams never wait for a copy
        # (with forward_host on rotating streams every stream spent ~1 ms of its cycle in its own copies: 4.34 ms/step;
        # pipelined: 4.01 ms with 3 slots, 3.94 ms with 4).
        # Every step still copies its own input from pinned host memory and lands its own result on the host inside
        # the timed region.
        n_dev_streams = n_streams
        n_slots = int(os.environ.get("BENCH_E2E_SLOTS", "4")) if n_dev_streams > 1 else 1
        copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        y_hosts = [y_host] + [torch.empty_like(y_host).pin_memory() for _ in range(n_slots - 1)]
        lo_hosts = [lo_host] + [torch.empty_like(lo_host).pin_memory() if lo_host is not None else None for _ in range(n_slots - 1)]
        e2e_count = [0]

        def step_e2e_pipelined():
            k = e2e_count[0]
            e2e_count[0] += 1
            with torch.cuda.stream(streams[k % n_dev_streams]):
                enc.forward_host_pipelined(x_host, lengths_host, y_hosts[k % n_slots], lo_hosts[k % n_slots], k % n_slots,
                                           copy_in, copy_out)

        def timed_pipelined(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            cur = torch.cuda.current_stream()
            e0.record(cur)
            for st in streams + [copy_in, copy_out]:
                st.wait_event(e0)
            for _ in range(steps):
                step_e2e_pipelined()
            for st in streams + [copy_in, copy_out]:
                cur.wait_stream(st)
            e1.record(cur)
            barrier()
            return max_over_ranks(e0.elapsed_time(e1), dist, dev) / steps

        for _ in range(2 * n_slots):
            step_e2e_pipelined()
        torch.cuda.synchronize()
        ms_e2e = timed_pipelined(args.steps)
        n_streams = n_slots
        clocks = sampler.stop() if rank == 0 else None       # sampled across the four timed regions above
        # profiled pass (per-launch CUDA events on the launching stream), not part of the headline
        enc.profile(True)
        prof_steps = min(args.steps, 5)
        for _ in range(prof_steps):
            step_dev()
        torch.cuda.synchronize()
        prof = enc.collect_profile()
        enc.profile(False)

    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
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
    prof_total = sum(k["ms_per_step"] for k in kernels.values())
    for k in kernels.values():
        k["share"] = k["ms_per_step"] / prof_total if prof_total else None
    # the roofline block describes ONE kernel (so does its ncu traffic figure): the most expensive class that is a single
    # kernel function; "gemm_tcgen05" aggregates four GEMM shapes (conv2, linear, q|k|v, out_proj) and stays in `kernels`
    single = [n for n in kernels if n in ("ffn_fused", "attention", "conv_fused", "conv1", "dwconv", "ln_gemm")]
    dom_name = max(single or kernels, key=lambda n: kernels[n]["ms_per_step"])
    dom = kernels[dom_name]
    if "tflops" in dom and dom_name in ("gemm_tcgen05", "attention", "gemm_fp32", "ffn_fused", "ln_gemm"):
        roof = {"kernel": dom_name, "bound": "tensor", "achieved": dom["tflops"], "peak": peaks["tensor"],
                "unit": "TFLOP/s", "frac": dom["tflops"] / peaks["tensor"], "traffic": None,
                "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step)"}
    else:
        roof = {"kernel": dom_name, "bound": "hbm", "achieved": dom.get("gbs"), "peak": peaks["hbm"], "unit": "GB/s",
                "frac": (dom.get("gbs") or 0.0) / peaks["hbm"], "traffic": None, "peak_source": f"{peaks['src']} hbm_gbs"}
    # roofline.traffic: DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/)
    try:
        tr = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))
        names = {"ffn_fused": ["ffn_fused_kernel<1>", "ffn_fused_kernel"], "attention": ["rel_attention_fact_kernel", "rel_attention_tc_kernel"],
                 "conv_fused": ["conv_module_fused_kernel"], "gemm_tcgen05": ["gemm_tc_kernel"], "conv1": ["conv1_tc_kernel"]}
        for n in names.get(dom_name, []):
            hit = [k for k in tr["kernels"] if k.startswith(n)]
            if hit:
                roof["traffic"] = tr["kernels"][hit[0]]["dram_bytes_per_launch"]
                roof["traffic_source"] = "profiles/ncu_traffic.json (" + os.path.basename(tr["source"]) + "): " + hit[0]
                break
    except (OSError, ValueError, KeyError):
        pass
    if "dwconv" in kernels and "gbs" in kernels["dwconv"]:
        kernels["dwconv"]["hbm_frac"] = kernels["dwconv"]["gbs"] / peaks["hbm"]
    for n in ("gemm_tcgen05", "attention", "ffn_fused", "ln_gemm"):
        if n in kernels and "tflops" in kernels[n]:
            kernels[n]["tensor_frac"] = kernels[n]["tflops"] / peaks["tensor"]

    from oracle.conformer_oracle import encoder_flops
    flops = encoder_flops(batch, t_in, cfg.d_model, cfg.n_blocks, cfg.n_mel_channels, cfg.kernel_size)
    line = {
        "metric": "encoder audio-seconds per second (RTFx)", "value": whole_job_rtfx(world, secs_per_step, ms_dev),
        "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
        "serial_ms_per_step": ms_dev_serial, "serial_cuda_graph_ms_per_step": ms_graph_serial, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.compute == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": desc, "per_gpu_batch": batch, "mel_frames": t_in, "frames_out": t_out,
                   "audio_seconds_per_gpu_step": secs_per_step, "parallelism": f"dp{world} (independent replicas)",
                   "issue": f"steps rotate over {n_dev_streams} CUDA streams per GPU (independent batches overlap at kernel "
                            "boundaries); serial_ms_per_step is the single-stream time",
                   "l2": "no flush needed: one step streams ~1.3 GB of activations (conv1 output alone 958 MB) "
                         "through the 126 MB L2"},
        "e2e": {"value": whole_job_rtfx(world, secs_per_step, ms_e2e), "unit": "audio-s/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": x_host.numel() * 4 + (lengths_host.numel() * 8 if var_len else 0),
                "d2h_bytes_per_step": y_host.numel() * 4 + (lo_host.numel() * 8 if var_len else 0),
                "serial_ms_per_step": ms_e2e_serial,
                "api": "cfm_encoder_forward_host_pipelined (Encoder.forward_host_pipelined): pinned host mel -> device -> "
                       f"forward -> host; {n_slots} in-flight slots over {n_dev_streams} compute streams, copies on a copy-in and "
                       "a copy-out stream ordered by the library's per-slot events; serial_ms_per_step is "
                       "cfm_encoder_forward_host (copy, forward, copy) on a single stream",
                "slots": n_slots, "compute_streams": n_dev_streams},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "algorithmic_tflops": flops / (ms_dev * 1e-3) / 1e12,
        "tensor_frac_whole_step": flops / (ms_dev * 1e-3) / 1e12 / peaks["tensor"],
        "roofline": roof,
        "kernels": kernels,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference_rtfx(cfg, t_in, var_len)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
