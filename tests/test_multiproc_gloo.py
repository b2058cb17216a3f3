"""world_size-2 CPU (gloo) tests of the multi-rank harness logic: max-over-ranks timing, per-rank shards,
whole-job aggregation, and the reference arm under torchrun (rank 0 alone runs and prints)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = 10.0 + 5.0 * rank                         # rank 1 is the slow one
        job_ms = bench.max_over_ranks(mine, dist, "cpu")
        from conformer_b200.synth import synth_logmel
        x = synth_logmel(2, 80, 50, seed=bench.rank_seed(rank))
        sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(sums, x.double().sum().reshape(1))
        out.put((rank, job_ms, [float(s) for s in sums], bench.whole_job_rtfx(world, 9.6, job_ms)))
    finally:
        dist.destroy_process_group()


def test_max_over_ranks_and_distinct_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, job_ms, sums, value in res:
        assert job_ms == 15.0                            # the slowest rank defines the step
        assert sums[0] != sums[1]                        # ranks benchmark different utterances
        assert abs(value - 2 * 9.6 / 0.015) < 1e-6       # whole-job aggregate, not per-GPU


def test_strong_scaling_split_covers_the_batch():
    import bench
    for batch, world in ((64, 1), (64, 2), (64, 8), (32, 3), (5, 8)):
        parts = bench.strong_scaling_batches(batch, world)
        assert len(parts) == world and sum(parts) == batch and max(parts) - min(parts) <= 1


def test_reference_arm_under_torchrun_rank0_only():
    port = 31500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--config", "T0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    # "reference" when baseline/_ref travelled with the repo (baseline/install_ref.py), else the oracle port
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0


def _reducer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import conformer_b200 as cb
    from conformer_b200.training import GradientAllReducer, TrainState
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = cb.Conformer(11, 80, 2, 32, 2, 7, 16, 1)
        enc = model.encoder
        st = TrainState(enc, torch.device("cpu"))
        object.__setattr__(enc, "_train_state", st)
        # what the native backward leaves behind: per-rank gradients in the flat buffer, .grad aliasing its slices
        st.flat_grad.copy_(torch.arange(st.flat_grad.numel(), dtype=torch.float32) * (rank + 1))
        st.begin_backward.__func__  # noqa: B018 - the attribute exists
        for (_, p), v in zip(st.params, st.views):
            p.grad = v
        for p in model.decoder.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        red = GradientAllReducer(model)
        red.arm()
        for i in (1, 0, -1):                               # the order the backward finishes them in
            st.on_block_done(i, *st.block_slice(i))
        red.finish()
        want = torch.arange(st.flat_grad.numel(), dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok_enc = torch.allclose(st.flat_grad, want)
        ok_dec = all(torch.allclose(p.grad, torch.full_like(p, sum(range(1, world + 1)) / world)) for p in model.decoder.parameters())
        slices = [st.block_slice(-1), st.block_slice(0), st.block_slice(1)]
        out.put((rank, ok_enc, ok_dec, slices, st.flat_grad.numel(), red.bytes_reduced))
    finally:
        dist.destroy_process_group()


def test_gradient_all_reducer_averages_block_slices_and_head():
    """World-size-2 gloo run of the data-parallel gradient exchange of config 4: per-block slices of the encoder's flat
    gradient buffer (in the order the backward finishes them) and the CTC head's gradients are averaged over the ranks;
    the slices partition the buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_enc, ok_dec, slices, total, nbytes in res:
        assert ok_enc and ok_dec
        assert slices[0][0] == 0 and slices[0][1] == slices[1][0] and slices[1][1] == slices[2][0] and slices[2][1] == total
        assert nbytes >= total * 4
