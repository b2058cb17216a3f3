"""Training step of the drop-in Encoder (BASELINE config 4; the reference's loop is train.py:219-249).

``model.train(); out, out_len = model(x, lengths); loss.backward(); optimizer.step()`` works as with the reference:
in training mode ``Encoder.forward`` runs cfm_encoder_train_forward (dropout, batch-statistics BatchNorm, activations
kept on a tape) and returns a tensor whose autograd node runs cfm_encoder_train_backward, which writes the fp32 gradient
of every parameter straight into ``param.grad``.  All ``param.grad`` tensors of an encoder are views of ONE flat fp32
buffer in parameter order, so a block's gradients are one contiguous slice: ``GradientAllReducer`` hands such slices to
NCCL while the backward of the earlier blocks is still running - the reference gets the same overlap from
DistributedDataParallel's bucketed hooks (train.py:185-186,239).

Host-side plumbing only: every FLOP of the encoder is in the native library; there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib
from .modules import _on_device, _ptr, _stream


class TrainState:
    """Native training handle of one Encoder on one device + the flat gradient buffer its parameters' .grad alias."""

    def __init__(self, enc, device) -> None:
        self.device = device
        self.handle: Optional[C.c_void_p] = None
        self.key = None
        self.compute = None
        self.workspace: Optional[torch.Tensor] = None
        self.seed = 0x5EED
        self.step = 0
        self.graphs = True               # replay the step's launch sequences as CUDA graphs (cfm_train_set_graphs)
        self.on_block_done: Optional[Callable[[int, int, int], None]] = None      # (block index or -1 for the front-end, lo, hi)
        self.params = [(n, p) for n, p in enc.named_parameters() if p.requires_grad]
        sizes = [p.numel() for _, p in self.params]
        self.flat_grad = torch.zeros(sum(sizes), device=device, dtype=torch.float32)
        self.views, self.offsets, off = [], {}, 0
        for (n, p), sz in zip(self.params, sizes):
            self.views.append(self.flat_grad[off:off + sz].view_as(p))
            self.offsets[n] = (off, off + sz)
            off += sz

    def block_slice(self, i: int) -> Tuple[int, int]:
        """[lo, hi) of block i's parameters in the flat gradient buffer; i = -1: the front-end (convs + linear)."""
        pre = f"layers.{i}." if i >= 0 else None
        sel = [v for k, v in self.offsets.items() if (k.startswith(pre) if pre else not k.startswith("layers."))]
        return min(a for a, _ in sel), max(b for _, b in sel)

    def begin_backward(self) -> None:
        """The native backward ACCUMULATES into the flat buffer.  A parameter whose .grad is None (the optimizer's
        zero_grad(set_to_none=True), which the reference calls between forward and backward, train.py:238) or a foreign
        tensor starts from a zeroed slice; a parameter whose .grad already is its slice keeps accumulating, like autograd."""
        self._fresh = [i for i, ((_, p), v) in enumerate(zip(self.params, self.views))
                       if p.grad is None or p.grad.data_ptr() != v.data_ptr()]
        if len(self._fresh) == len(self.views):
            self.flat_grad.zero_()
        else:
            for i in self._fresh:
                self.views[i].zero_()

    def end_backward(self) -> None:
        for i in self._fresh:
            p, v = self.params[i][1], self.views[i]
            if p.grad is None:
                p.grad = v                       # .grad now aliases the flat buffer (what the all-reduce works on)
            else:
                p.grad.add_(v)                   # a gradient tensor someone else installed: add ours into it

    def bind(self, enc, compute: int) -> None:
        named = {k: v for k, v in enc.state_dict(keep_vars=True).items() if v.is_floating_point()}
        key = (compute,) + tuple(t.data_ptr() for t in named.values())
        if self.handle is not None and key == self.key:
            return
        if self.handle is not None:
            lib().cfm_module_destroy(self.handle)
            self.handle = None
        for name, t in named.items():
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError(f"conformer_b200: training needs contiguous fp32 parameters on {self.device} ('{name}')")
        h = C.c_void_p()
        cfg = enc._config()
        check(lib().cfm_module_create(enc._KIND, C.byref(cfg), C.byref(h)))
        try:
            for name, t in named.items():
                check(lib().cfm_module_set_param(h, name.encode(), t.data_ptr(), t.numel()))
            for (name, _), v in zip(self.params, self.views):
                check(lib().cfm_module_set_grad(h, name.encode(), v.data_ptr(), v.numel()))
            check(lib().cfm_train_bind(h, compute))
            check(lib().cfm_train_set_graphs(h, 1 if self.graphs else 0))
        except Exception:
            lib().cfm_module_destroy(h)
            raise
        self.handle, self.key, self.compute = h, key, compute

    def set_graphs(self, enable: bool) -> None:
        """CUDA-graph replay of the forward / per-block backward / weight refresh (on by default)."""
        self.graphs = bool(enable)
        if self.handle is not None:
            check(lib().cfm_train_set_graphs(self.handle, 1 if self.graphs else 0))

    def graph_stats(self) -> dict:
        """{'direct', 'captured', 'replayed', 'failed'} call counts of the native handle (cfm_train_graph_stats)."""
        out = (C.c_int64 * 4)()
        if self.handle is not None:
            check(lib().cfm_train_graph_stats(self.handle, out))
        return dict(zip(("direct", "captured", "replayed", "failed"), (int(v) for v in out)))

    def __del__(self):
        try:
            if self.handle is not None:
                lib().cfm_module_destroy(self.handle)
        except Exception:
            pass


class _EncoderTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook: torch.Tensor, enc, state: TrainState, x: torch.Tensor, lengths: Optional[torch.Tensor]):
        B, _, Tin = x.shape
        dev = x.device
        with _on_device(dev):
            need = lib().cfm_encoder_train_workspace_bytes(state.handle, B, Tin)
            if need < 0:
                check(1)
            if state.workspace is None or state.workspace.numel() < need:
                state.workspace = torch.empty(int(need), device=dev, dtype=torch.uint8)
            y = torch.empty((B, enc.output_frames(Tin), enc._cfg[2]), device=dev, dtype=torch.float32)
            lin = lout = None
            if lengths is not None:
                lin = lengths.to(device=dev, dtype=torch.int64).contiguous()
                lout = torch.empty_like(lin)
            state.step += 1
            check(lib().cfm_train_refresh(state.handle, _stream()))
            check(lib().cfm_encoder_train_forward(state.handle, x.data_ptr(), _ptr(lin), y.data_ptr(), _ptr(lout), B, Tin,
                                                  enc._dropout_rate, state.seed + state.step, state.workspace.data_ptr(),
                                                  state.workspace.numel(), _stream()))
        ctx.enc, ctx.state, ctx.x = enc, state, x          # x (the mel batch) is read again by conv_1's weight gradient
        ctx.mark_non_differentiable(lout) if lout is not None else None
        return (y, lout) if lout is not None else (y,)

    @staticmethod
    def backward(ctx, dy, *unused):
        enc, state = ctx.enc, ctx.state
        L = enc._cfg[1]
        dy = dy.to(torch.float32).contiguous()
        ws = state.workspace
        state.begin_backward()
        with _on_device(dy.device):
            # block by block, last to first: after each call that block's parameter gradients are final
            for i in range(L - 1, -1, -1):
                check(lib().cfm_encoder_train_backward(state.handle, dy.data_ptr() if i == L - 1 else None, i + 1, i, 0,
                                                       ws.data_ptr(), ws.numel(), _stream()))
                if state.on_block_done is not None:
                    state.on_block_done(i, *state.block_slice(i))
            check(lib().cfm_encoder_train_backward(state.handle, dy.data_ptr() if L == 0 else None, 0, 0, 1, ws.data_ptr(),
                                                   ws.numel(), _stream()))
            if state.on_block_done is not None:
                state.on_block_done(-1, *state.block_slice(-1))
        state.end_backward()
        return None, None, None, None, None


def encoder_train_state(enc, device=None) -> TrainState:
    """The native training state of an Encoder on `device` (default: where its parameters are), created on first use."""
    dev = torch.device(device) if device is not None else next(enc.parameters()).device
    st = enc._train_state
    if st is None or st.device != dev:
        st = TrainState(enc, dev)
        object.__setattr__(enc, "_train_state", st)
    return st


def encoder_train_forward(enc, x: torch.Tensor, lengths: Optional[torch.Tensor]):
    """Encoder.forward in training mode (called by modules.Encoder.forward)."""
    dev = x.device
    st = encoder_train_state(enc, dev)
    compute = enc._compute()
    with _on_device(dev):
        st.bind(enc, compute)
    if not hasattr(st, "hook") or st.hook.device != dev:
        st.hook = torch.zeros(1, device=dev, requires_grad=True)     # makes autograd call our backward
    out = _EncoderTrainFn.apply(st.hook, enc, st, x, lengths)
    for m in enc.modules():                                           # nn.BatchNorm1d bookkeeping (convolution.py:16)
        if isinstance(m, torch.nn.BatchNorm1d) and m.num_batches_tracked is not None:
            m.num_batches_tracked += 1
    return (out[0], out[1]) if lengths is not None else (out[0], None)


class GradientAllReducer:
    """Data-parallel gradient averaging for a drop-in Conformer / Encoder (the reference: DistributedDataParallel,
    train.py:185-186).  The encoder's gradients are averaged block by block WHILE its backward runs: as soon as the native
    backward of block i has been enqueued, the contiguous slice of the flat gradient buffer holding block i's parameters
    is handed to an asynchronous NCCL all-reduce (NCCL works on its own stream and waits for the compute stream through
    an event, so communication of block i overlaps the backward kernels of block i-1).  ``finish()`` all-reduces what is
    left (the CTC head's parameters, which autograd fills in) and waits for everything."""

    def __init__(self, model: torch.nn.Module, process_group=None) -> None:
        import torch.distributed as dist
        self.dist, self.group, self.model = dist, process_group, model
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.pending: List = []
        self.encoder = next((m for m in model.modules() if hasattr(m, "_train_state")), None)
        self.bytes_reduced = 0

    def _enc_state(self) -> Optional[TrainState]:
        return None if self.encoder is None else self.encoder._train_state

    def arm(self) -> None:
        """Call after the forward of a step (the encoder's training state exists from its first forward on)."""
        st = self._enc_state()
        if st is not None and self.world > 1:
            st.on_block_done = self._on_block
        self.pending, self.bytes_reduced = [], 0

    def _on_block(self, i: int, lo: int, hi: int) -> None:
        st = self._enc_state()
        buf = st.flat_grad[lo:hi]
        self.bytes_reduced += buf.numel() * 4
        self.pending.append(self.dist.all_reduce(buf, op=self.dist.ReduceOp.AVG, group=self.group, async_op=True))

    def finish(self) -> None:
        if self.world > 1:
            st = self._enc_state()
            mine = set() if st is None else {id(p) for _, p in st.params}
            rest = [p.grad for p in self.model.parameters() if p.grad is not None and id(p) not in mine]
            if rest:
                flat = torch.cat([g.reshape(-1) for g in rest])
                self.bytes_reduced += flat.numel() * 4
                self.dist.all_reduce(flat, op=self.dist.ReduceOp.AVG, group=self.group)
                off = 0
                for g in rest:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
            for w in self.pending:
                w.wait()
        self.pending = []
