"""Drop-in nn.Module mirrors of the reference's encoder-path modules.

Same class names, constructor arguments, attribute names and ``state_dict`` keys as
Alan-404/Conformer, so ``load_state_dict(reference.state_dict())`` is
strict-clean and a reference script can swap its imports.  The sub-modules created in each
``__init__`` (nn.Linear, nn.LayerNorm, nn.Conv1d, ...) are *parameter containers only*: no
forward of theirs is ever called.  Every ``forward`` here unwraps device pointers and calls the
C-ABI library (include/conformer_b200.h) that runs hand-written sm_100a kernels.  There is no
CPU path and no PyTorch fallback: CPU tensors raise.

Arithmetic: fp32 by default (CUDA-core kernels, reference-fp32 accuracy); under
``torch.autocast('cuda', ...)`` - the way the reference selects low precision, train.py:232,
test.py:147 - or after ``module.set_compute('bf16')``, GEMMs run in bf16 on the tcgen05 tensor
cores with fp32 accumulate and an fp32 residual stream.

Inference only in this round: ``forward`` in training mode raises (eval-mode BatchNorm, no
dropout, no backward).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import Config, check, lib

__all__ = ["Swish", "GLU", "generate_padding_mask", "RelativePositionalEncoding", "FeedForwardModule",
           "RelativeMultiHeadAttention", "MultiHeadSelfAttentionModule", "ConvolutionModule",
           "ConvolutionSubsampling", "ConformerBlock", "Encoder", "Decoder", "Conformer"]


def _stream(device=None) -> int:
    """The current stream OF THE TENSORS' DEVICE (the reference's ATen ops launch on the device their inputs live on)."""
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """Make ``device`` current around a native call: the library launches on (and its handles are tied to) the current
    CUDA device, while PyTorch callers expect ``module(x)`` to work whatever device is current."""

    def __init__(self, device) -> None:
        self.device, self.prev = torch.device(device), None

    def __enter__(self):
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if idx != torch.cuda.current_device():
            self.prev = torch.cuda.current_device()
            torch.cuda.set_device(idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _dev_f32(x: torch.Tensor, what: str) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError(f"conformer_b200: {what} must be a CUDA tensor (there is no CPU path)")
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.to(torch.float32).contiguous()
    return x


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# =============================================================================== stateless ops
class Swish(nn.Module):
    """model/utils/activation.py:4-8"""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xin = _dev_f32(x, "x")
        y = torch.empty_like(xin)
        if xin.numel():
            with _on_device(xin.device):
                check(lib().cfm_swish(xin.data_ptr(), y.data_ptr(), xin.numel(), _stream(xin.device)))
        return y.to(x.dtype)


class GLU(nn.Module):
    """model/utils/activation.py:10-17 - first half * sigmoid(second half) along ``dim``."""

    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xin = _dev_f32(x.movedim(self.dim, -1), "x")
        half = xin.shape[-1] // 2
        y = torch.empty(xin.shape[:-1] + (half,), device=xin.device, dtype=torch.float32)
        rows = y.numel() // max(half, 1)
        if y.numel():
            with _on_device(xin.device):
                check(lib().cfm_glu(xin.data_ptr(), y.data_ptr(), rows, half, _stream(xin.device)))
        return y.movedim(-1, self.dim).to(x.dtype)


def generate_padding_mask(lengths: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    """model/utils/masking.py:4-13 - bool [B, max_length], True where t < lengths[b]."""
    if not lengths.is_cuda:
        raise RuntimeError("conformer_b200: lengths must be a CUDA tensor (there is no CPU path)")
    if max_length is None:
        max_length = int(lengths.max())          # same host sync as the reference (masking.py:9-10)
    max_length = int(max_length)
    l64 = lengths.to(torch.int64).contiguous()
    out = torch.empty((l64.numel(), max_length), device=lengths.device, dtype=torch.uint8)
    if out.numel():
        with _on_device(l64.device):
            check(lib().cfm_padding_mask(l64.data_ptr(), l64.numel(), max_length, out.data_ptr(), _stream(l64.device)))
    return out.view(torch.bool)


def greedy_ctc_decode(logits: torch.Tensor, pad_id: int = 0, unk_id: Optional[int] = None):
    """processing/processor.py:301-328 at token-id level, on the device: argmax over the vocabulary, skip pad (the CTC
    blank) and unk without resetting the previous id, collapse repeats; every row over the full padded length.
    logits [B, T, V] (CUDA).  Returns (tokens int32 [B, T] padded with -1, counts int32 [B]); `unk_id` defaults to the
    last vocabulary entry (processor.py:114)."""
    if not logits.is_cuda:
        raise RuntimeError("conformer_b200: logits must be a CUDA tensor (there is no CPU path)")
    if logits.dim() != 3:
        raise ValueError("logits must be [batch, frames, vocab]")
    B, T, V = logits.shape
    lg = logits.detach().to(torch.float32).contiguous()
    tokens = torch.empty((B, T), device=logits.device, dtype=torch.int32)
    counts = torch.zeros((B,), device=logits.device, dtype=torch.int32)
    if B and T:
        with _on_device(lg.device):
            check(lib().cfm_ctc_greedy_decode(lg.data_ptr(), B, T, V, int(pad_id), int(V - 1 if unk_id is None else unk_id),
                                              tokens.data_ptr(), counts.data_ptr(), _stream(lg.device)))
    return tokens, counts


# =============================================================================== native handle plumbing
class _Native(nn.Module):
    """Owns the cfm_module_t handles (one per compute mode) of a mirror module."""

    _KIND = -1

    def __init__(self) -> None:
        super().__init__()
        object.__setattr__(self, "_handles", {})       # compute -> (handle, version_key)
        object.__setattr__(self, "_ws", {})            # (device, stream) -> uint8 workspace tensor
        object.__setattr__(self, "_private_ws", None)  # set only while Encoder.capture records its graph
        self._compute_override: Optional[int] = None

    # -- configuration each subclass provides
    def _config(self) -> Config:
        raise NotImplementedError

    # -- compute mode
    def set_compute(self, mode: Optional[str]) -> "_Native":
        """'fp32', 'bf16' or None (= fp32 unless torch.autocast is active)."""
        table = {None: None, "fp32": _lib.COMPUTE_F32, "float32": _lib.COMPUTE_F32,
                 "bf16": _lib.COMPUTE_BF16, "bfloat16": _lib.COMPUTE_BF16}
        if mode not in table:
            raise ValueError(f"unknown compute mode {mode!r}")
        self._compute_override = table[mode]
        return self

    def _compute(self) -> int:
        if self._compute_override is not None:
            return self._compute_override
        return _lib.COMPUTE_BF16 if torch.is_autocast_enabled() else _lib.COMPUTE_F32

    # -- handle lifecycle
    def _tensors(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.state_dict(keep_vars=True).items() if v.is_floating_point()}

    def _version_key(self, tensors) -> Tuple:
        # (address, autograd version) per tensor: an in-place update or a re-assignment triggers a re-pack.  Tensors
        # created under torch.inference_mode() carry no version counter; for those only a new address is noticed.
        return tuple((t.data_ptr(), 0 if t.is_inference() else t._version) for t in tensors)

    def _drop_handles(self) -> None:
        for h, _, _ in self._handles.values():
            lib().cfm_module_destroy(h)
        self._handles.clear()
        self._ws.clear()

    def _apply(self, fn, *a, **k):                     # .to()/.cuda()/.float() move the parameters
        self._drop_handles()
        return super()._apply(fn, *a, **k)

    def __del__(self):
        try:
            self._drop_handles()
        except Exception:
            pass

    def __getstate__(self):                            # handles are process-local; never copy/pickle them
        state = dict(self.__dict__)
        state["_handles"], state["_ws"], state["_private_ws"] = {}, {}, None
        if "_train_state" in state:
            state["_train_state"] = None
        return state

    def _native(self) -> int:
        if self.training:
            raise RuntimeError("conformer_b200: training-mode forward is not implemented (inference only); "
                               "call .eval()")
        compute = self._compute()
        hkey = (compute, torch.cuda.current_device())
        entry = self._handles.get(hkey)
        if entry is not None:
            h, tensors, key = entry
            if key == self._version_key(tensors):
                return h
            lib().cfm_module_destroy(h)
            del self._handles[hkey]
        named = self._tensors()
        staged = {}
        for name, t in named.items():
            if not t.is_cuda:
                raise RuntimeError("conformer_b200: parameters must live on a CUDA device (there is no CPU path); "
                                   "call .cuda() on the module")
            if t.device.index != torch.cuda.current_device():
                raise RuntimeError(f"conformer_b200: parameter '{name}' is on {t.device} but the input is on "
                                   f"cuda:{torch.cuda.current_device()}; move the module and its inputs to one device")
            staged[name] = t.detach().to(torch.float32).contiguous()
        h = C.c_void_p()
        cfg = self._config()
        check(lib().cfm_module_create(self._KIND, C.byref(cfg), C.byref(h)))
        try:
            for name, t in staged.items():
                check(lib().cfm_module_set_param(h, name.encode(), t.data_ptr(), t.numel()))
            check(lib().cfm_module_finalize(h, compute, _stream()))
        except Exception:
            lib().cfm_module_destroy(h)
            raise
        tensors = list(named.values())
        self._handles[hkey] = (h, tensors, self._version_key(tensors))
        return h

    def _workspace(self, h, batch: int, frames: int, device) -> torch.Tensor:
        need = lib().cfm_workspace_bytes(h, batch, frames)
        if need < 0:
            check(1)
        if self._private_ws is not None:   # Encoder.capture: the graph owns its scratch (see CapturedEncoderForward)
            assert self._private_ws.numel() >= need and self._private_ws.device == device
            return self._private_ws
        stream = torch.cuda.current_stream(device)
        key = (device, stream.cuda_stream)  # forwards on different streams may overlap on the device
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            if ws is not None:
                ws.record_stream(stream)   # work already queued on this stream may still use the old block
            ws = torch.empty(int(need), device=device, dtype=torch.uint8)
            self._ws[key] = ws
        return ws

    @property
    def last_launch_count(self) -> int:
        """Kernel launches enqueued by the most recent forward (any compute mode)."""
        return max([int(lib().cfm_last_launch_count(h)) for h, _, _ in self._handles.values()] or [0])

    KERNEL_CLASSES = ("gemm_tcgen05", "gemm_fp32", "layer_norm", "attention", "dwconv", "conv1", "misc", "ffn_fused", "conv_fused")

    def profile(self, on: bool) -> None:
        """Bracket every kernel launch of subsequent forwards with CUDA events (see cfm_profile_enable)."""
        self._native()
        for h, _, _ in self._handles.values():
            check(lib().cfm_profile_enable(h, 1 if on else 0))

    def collect_profile(self) -> Dict[str, Dict[str, float]]:
        """Per kernel class: launches, milliseconds, algorithmic flops and bytes since the last collect."""
        n = len(self.KERNEL_CLASSES)
        out: Dict[str, Dict[str, float]] = {}
        for h, _, _ in self._handles.values():
            buf = (C.c_double * (n * 4))()
            check(lib().cfm_profile_collect(h, buf, n))
            for i, name in enumerate(self.KERNEL_CLASSES):
                e = out.setdefault(name, dict(launches=0.0, ms=0.0, flops=0.0, bytes=0.0))
                e["launches"] += buf[i * 4]
                e["ms"] += buf[i * 4 + 1]
                e["flops"] += buf[i * 4 + 2]
                e["bytes"] += buf[i * 4 + 3]
        return out


def _key_pad(mask: Optional[torch.Tensor], B: int, T: int) -> Optional[torch.Tensor]:
    """The reference passes mask as bool [B,1,1,T], True = padded key (encoder.py:30)."""
    if mask is None:
        return None
    m = mask.reshape(mask.shape[0], -1, mask.shape[-1])
    if m.shape[1] != 1 or m.shape[-1] != T:
        raise RuntimeError("conformer_b200: only key-padding masks of shape [B,1,1,T] are supported")
    m = m[:, 0, :].expand(B, T)
    if not m.is_cuda:
        raise RuntimeError("conformer_b200: mask must be a CUDA tensor")
    return m.to(torch.uint8).contiguous()


def _pos_table(pos_embedding: Optional[torch.Tensor], T: int, d: int) -> Optional[torch.Tensor]:
    """The reference replicates the [2T-1,d] table over the batch (position.py:26); one copy is used."""
    if pos_embedding is None:
        return None
    p = pos_embedding[0] if pos_embedding.dim() == 3 else pos_embedding
    if p.shape != (2 * T - 1, d):
        raise RuntimeError(f"conformer_b200: pos_embedding must be [B, 2T-1, d] = [*, {2 * T - 1}, {d}], "
                           f"got {tuple(pos_embedding.shape)}")
    return _dev_f32(p, "pos_embedding")


# =============================================================================== mirrors
class RelativePositionalEncoding(nn.Module):
    """model/utils/position.py:5-28"""

    def __init__(self, d_model: int) -> None:
        super().__init__()
        self.d_model = d_model
        self.div_term = nn.Parameter(
            torch.exp(torch.arange(0, self.d_model, 2) * -(math.log(10000.0) / d_model)).unsqueeze(0),
            requires_grad=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, T = x.size(0), x.size(1)
        dt = _dev_f32(self.div_term.detach(), "div_term")
        out = torch.empty((B, 2 * T - 1, self.d_model), device=dt.device, dtype=torch.float32)
        with _on_device(dt.device):
            check(lib().cfm_rel_pos_encoding(dt.data_ptr(), B, T, self.d_model, out.data_ptr(), _stream()))
        return out


class FeedForwardModule(_Native):
    """model/utils/ffn.py:5-23"""
    _KIND = _lib.KIND_FFN

    def __init__(self, dim: int, dropout_rate: float = 0.0) -> None:
        super().__init__()
        self.layer_norm = nn.LayerNorm(normalized_shape=dim)
        self.hidden_linear = nn.Linear(in_features=dim, out_features=4 * dim)
        self.swish = Swish()
        self.dropout_1 = nn.Dropout(p=dropout_rate)
        self.out_linear = nn.Linear(in_features=4 * dim, out_features=dim)
        self.dropout_2 = nn.Dropout(p=dropout_rate)
        self._dim = dim

    def _config(self) -> Config:
        return Config(0, 0, self._dim, 1, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xin = _dev_f32(x, "x")
        rows, T = xin.numel() // self._dim // max(xin.shape[-2], 1), xin.shape[-2]
        y = torch.empty_like(xin)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, rows, T, xin.device)
            check(lib().cfm_ffn_forward(h, xin.data_ptr(), y.data_ptr(), rows, T, ws.data_ptr(), ws.numel(), _stream()))
        return y


class RelativeMultiHeadAttention(_Native):
    """model/utils/attention.py:20-102 (self-attention use only: q is k is v, as at attention.py:16)."""
    _KIND = _lib.KIND_REL_ATTENTION

    def __init__(self, d_model: int, n_heads: int, dropout_rate: float = 0.0):
        super().__init__()
        assert d_model % n_heads == 0
        self.d_model = d_model
        self.n_heads = n_heads
        self.head_samples = d_model // n_heads
        self.sqrt_dim = math.sqrt(self.head_samples)
        self.query_proj = nn.Linear(d_model, d_model)
        self.key_proj = nn.Linear(d_model, d_model)
        self.value_proj = nn.Linear(d_model, d_model)
        self.pos_proj = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.content_bias = nn.Parameter(torch.empty(n_heads, self.head_samples))
        self.position_bias = nn.Parameter(torch.empty(n_heads, self.head_samples))
        self.out_proj = nn.Linear(d_model, d_model)
        self.mask_value = None
        nn.init.xavier_uniform_(self.content_bias)
        nn.init.xavier_uniform_(self.position_bias)

    def _config(self) -> Config:
        return Config(0, 0, self.d_model, self.n_heads, 1)

    def _run(self, x, pos_embedding, mask):
        xin = _dev_f32(x, "x")
        B, T, d = xin.shape
        y = torch.empty_like(xin)
        pos = _pos_table(pos_embedding, T, d)
        kp = _key_pad(mask, B, T)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, B, T, xin.device)
            check(lib().cfm_mhsa_forward(h, xin.data_ptr(), _ptr(pos), _ptr(kp), y.data_ptr(), B, T, ws.data_ptr(),
                                         ws.numel(), _stream()))
        return y

    def forward(self, q, k, v, pos_embedding, mask=None):
        if not (q is k and k is v) and not (q.data_ptr() == k.data_ptr() == v.data_ptr() and q.shape == k.shape == v.shape):
            raise NotImplementedError("conformer_b200: only self-attention (q is k is v) is on the encoder path")
        return self._run(q, pos_embedding, mask)


class MultiHeadSelfAttentionModule(_Native):
    """model/utils/attention.py:7-18"""
    _KIND = _lib.KIND_MHSA

    def __init__(self, d_model: int, n_heads: int, dropout_rate: float = 0.0) -> None:
        super().__init__()
        self.layer_norm = nn.LayerNorm(normalized_shape=d_model)
        self.attention = RelativeMultiHeadAttention(d_model=d_model, n_heads=n_heads, dropout_rate=dropout_rate)
        self.dropout = nn.Dropout(p=dropout_rate)
        self._d, self._h = d_model, n_heads

    def _config(self) -> Config:
        return Config(0, 0, self._d, self._h, 1)

    def forward(self, x: torch.Tensor, pos_embedding: torch.Tensor, mask: Optional[torch.Tensor] = None):
        return RelativeMultiHeadAttention._run(self, x, pos_embedding, mask)


class ConvolutionModule(_Native):
    """model/utils/convolution.py:7-32"""
    _KIND = _lib.KIND_CONV_MODULE

    def __init__(self, channels: int, kernel_size: int, dropout_rate: float = 0.0) -> None:
        super().__init__()
        padding = (kernel_size - 1) // 2
        self.layer_norm = nn.LayerNorm(normalized_shape=channels)
        self.pointwise_conv_1 = nn.Conv1d(channels, channels * 2, kernel_size=1)
        self.glu = GLU(dim=1)
        self.deepwise_conv = nn.Conv1d(channels, channels, kernel_size=kernel_size, padding=padding, groups=channels)
        self.batch_norm = nn.BatchNorm1d(num_features=channels)
        self.swish = Swish()
        self.pointwise_conv_2 = nn.Conv1d(channels, channels, kernel_size=1)
        self.dropout = nn.Dropout(p=dropout_rate)
        self._c, self._k = channels, kernel_size

    def _config(self) -> Config:
        return Config(0, 0, self._c, 1, self._k)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xin = _dev_f32(x, "x")
        B, T, _ = xin.shape
        y = torch.empty_like(xin)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, B, T, xin.device)
            check(lib().cfm_conv_module_forward(h, xin.data_ptr(), y.data_ptr(), B, T, ws.data_ptr(), ws.numel(), _stream()))
        return y


class ConvolutionSubsampling(_Native):
    """model/utils/convolution.py:34-57"""
    _KIND = _lib.KIND_SUBSAMPLING

    def __init__(self, channels: int) -> None:
        super().__init__()
        self.conv_1 = nn.Conv2d(in_channels=1, out_channels=channels, kernel_size=3, stride=2)
        self.act_1 = nn.ReLU()
        self.conv_2 = nn.Conv2d(in_channels=channels, out_channels=channels, kernel_size=3, stride=2)
        self.act_2 = nn.ReLU()
        self._c = channels
        self._n_mel = 80

    def _config(self) -> Config:
        return Config(self._n_mel, 0, self._c, 1, 1)

    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        xin = _dev_f32(x, "x")
        B, n_mel, Tin = xin.shape
        if n_mel != self._n_mel:                       # the module itself is agnostic to n_mel
            self._n_mel = n_mel
            self._drop_handles()
        T2 = ((Tin - 1) // 2 - 1) // 2
        F2 = ((n_mel - 1) // 2 - 1) // 2
        y = torch.empty((B, T2, self._c * F2), device=xin.device, dtype=torch.float32)
        lin = lout = None
        if lengths is not None:
            lin = lengths.to(device=xin.device, dtype=torch.int64).contiguous()
            lout = torch.empty_like(lin)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, B, Tin, xin.device)
            check(lib().cfm_subsampling_forward(h, xin.data_ptr(), _ptr(lin), y.data_ptr(), _ptr(lout), B, Tin,
                                                ws.data_ptr(), ws.numel(), _stream()))
        return y, lout


class ConformerBlock(_Native):
    """model/utils/block.py:8-29"""
    _KIND = _lib.KIND_BLOCK

    def __init__(self, d_model: int, n_heads: int, kernel_size: int, dropout_rate: float = 0.0) -> None:
        super().__init__()
        self.ffn_1 = FeedForwardModule(dim=d_model, dropout_rate=dropout_rate)
        self.attention = MultiHeadSelfAttentionModule(n_heads=n_heads, d_model=d_model, dropout_rate=dropout_rate)
        self.conv = ConvolutionModule(channels=d_model, kernel_size=kernel_size, dropout_rate=dropout_rate)
        self.ffn_2 = FeedForwardModule(dim=d_model, dropout_rate=dropout_rate)
        self.layer_norm = nn.LayerNorm(normalized_shape=d_model)
        self._cfg = (d_model, n_heads, kernel_size)

    def _config(self) -> Config:
        d, h, k = self._cfg
        return Config(0, 1, d, h, k)

    def forward(self, x: torch.Tensor, pos_embedding: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        xin = _dev_f32(x, "x")
        B, T, d = xin.shape
        y = torch.empty_like(xin)
        pos = _pos_table(pos_embedding, T, d)
        kp = _key_pad(mask, B, T)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, B, T, xin.device)
            check(lib().cfm_block_forward(h, xin.data_ptr(), _ptr(pos), _ptr(kp), y.data_ptr(), B, T, ws.data_ptr(),
                                          ws.numel(), _stream()))
        return y


class Encoder(_Native):
    """model/modules/encoder.py:9-37"""
    _KIND = _lib.KIND_ENCODER

    def __init__(self, n_mel_channels: int, n_blocks: int, d_model: int, n_heads: int, kernel_size: int,
                 dropout_rate: float = 0.0) -> None:
        super().__init__()
        self.downsampling_conv = ConvolutionSubsampling(channels=d_model)
        self.linear = nn.Linear(in_features=d_model * (((n_mel_channels - 1) // 2 - 1) // 2), out_features=d_model)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.rel_pe = RelativePositionalEncoding(d_model=d_model)
        self.layers = nn.ModuleList([ConformerBlock(d_model=d_model, n_heads=n_heads, kernel_size=kernel_size,
                                                    dropout_rate=dropout_rate) for _ in range(n_blocks)])
        self._cfg = (n_mel_channels, n_blocks, d_model, n_heads, kernel_size)
        self._dropout_rate = float(dropout_rate)
        object.__setattr__(self, "_train_state", None)     # conformer_b200.training.TrainState, built lazily

    def _config(self) -> Config:
        return Config(*self._cfg)

    def output_frames(self, t_in: int) -> int:
        return ((t_in - 1) // 2 - 1) // 2

    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        xin = _dev_f32(x, "x")
        B, n_mel, Tin = xin.shape
        if n_mel != self._cfg[0]:
            raise RuntimeError(f"conformer_b200: expected {self._cfg[0]} mel channels, got {n_mel}")
        if self.training:
            from .training import encoder_train_forward
            return encoder_train_forward(self, xin, lengths)
        y = torch.empty((B, self.output_frames(Tin), self._cfg[2]), device=xin.device, dtype=torch.float32)
        lin = lout = None
        if lengths is not None:
            lin = lengths.to(device=xin.device, dtype=torch.int64).contiguous()
            lout = torch.empty_like(lin)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, B, Tin, xin.device)
            check(lib().cfm_encoder_forward(h, xin.data_ptr(), _ptr(lin), y.data_ptr(), _ptr(lout), B, Tin,
                                            ws.data_ptr(), ws.numel(), _stream()))
        return y, lout

    def capture(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> "CapturedEncoderForward":
        """Capture one forward for this batch shape in a CUDA graph.  The forward is static - no host synchronisation, no
        allocation, caller-owned workspace - so its ~130 kernel launches replay as one graph launch: the per-kernel launch
        latency a single stream otherwise pays at every kernel boundary disappears.  Returns an object whose ``x`` /
        ``lengths`` are the static input buffers (copy new batches into them), ``replay()`` enqueues the forward on the
        current stream and ``y`` / ``out_lengths`` receive the result.  Re-capture after loading new weights."""
        xin = _dev_f32(x, "x").clone()
        lin = None if lengths is None else lengths.to(device=xin.device, dtype=torch.int64).clone()
        with _on_device(xin.device):
            need = lib().cfm_workspace_bytes(self._native(), xin.shape[0], xin.shape[2])
            if need < 0:
                check(1)
            # the graph bakes in the scratch pointer: give it a block of its own, kept alive by the returned object, so
            # that no later forward (same stream id, larger shape, another thread) can free or share it
            ws = torch.empty(int(need), device=xin.device, dtype=torch.uint8)
            cur = torch.cuda.current_stream(xin.device)
            side = torch.cuda.Stream(device=xin.device)
            side.wait_stream(cur)
            self._private_ws = ws
            try:
                with torch.cuda.stream(side):         # warm-up on the capture stream: weight packing and kernel
                    for _ in range(2):                # attributes exist before the capture starts
                        self.forward(xin, lin)
                cur.wait_stream(side)
                torch.cuda.synchronize(xin.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    y, lout = self.forward(xin, lin)
            finally:
                self._private_ws = None
        return CapturedEncoderForward(graph, xin, lin, y, lout, ws)

    def _host_call(self, x_host, lengths_host, y_host, lengths_out_host, ws_key, device, copy_in_stream, copy_out_stream):
        if x_host.is_cuda or y_host.is_cuda or not x_host.is_contiguous() or not y_host.is_contiguous():
            raise RuntimeError("conformer_b200: forward_host takes contiguous HOST tensors (pinned for asynchronous copies)")
        dt = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}
        if x_host.dtype not in dt or y_host.dtype not in dt:
            raise RuntimeError("conformer_b200: host buffers must be float32 or bfloat16")
        B, _, Tin = x_host.shape
        device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        if tuple(y_host.shape) != (B, self.output_frames(Tin), self._cfg[2]):
            raise RuntimeError(f"conformer_b200: y_host must be [B, T', d] = {(B, self.output_frames(Tin), self._cfg[2])}")
        with _on_device(device):
            h = self._native()
            need = lib().cfm_encoder_host_workspace_bytes(h, B, Tin)
            if need < 0:
                check(1)
            key = ws_key + (device,) if ws_key[0] == "host-slot" else ws_key + (device, _stream())
            ws = self._ws.get(key)
            if ws is None or ws.numel() < need:
                if ws is not None:
                    # copies and kernels queued against the old block (on up to three streams) must finish before it
                    # is freed, and the library must forget the slot's events, which are keyed by its address
                    for hh, _, _ in self._handles.values():
                        check(lib().cfm_encoder_host_slot_release(hh, ws.data_ptr()))
                    torch.cuda.current_stream(device).synchronize()
                ws = torch.empty(int(need), device=device, dtype=torch.uint8)
                self._ws[key] = ws
            check(lib().cfm_encoder_forward_host_io(
                h, x_host.data_ptr(), dt[x_host.dtype], _ptr(lengths_host), y_host.data_ptr(), dt[y_host.dtype],
                _ptr(lengths_out_host), B, Tin, ws.data_ptr(), ws.numel(), _stream(),
                None if copy_in_stream is None else copy_in_stream.cuda_stream,
                None if copy_out_stream is None else copy_out_stream.cuda_stream))

    def forward_host(self, x_host: torch.Tensor, lengths_host: Optional[torch.Tensor], y_host: torch.Tensor,
                     lengths_out_host: Optional[torch.Tensor] = None, device=None) -> None:
        """End-to-end call with HOST (ideally pinned) buffers through cfm_encoder_forward_host_io: the H2D copy of the mel
        batch, the forward and the D2H copy of the result are enqueued on the current stream.  ``x_host`` / ``y_host`` are
        float32 (the reference's tensors) or bfloat16 (half the PCIe bytes); ``y_host`` is [B, T', d].  Everything is
        stream-ordered and asynchronous; each stream gets its own device staging workspace."""
        self._host_call(x_host, lengths_host, y_host, lengths_out_host, ("host",), device, None, None)

    def forward_host_pipelined(self, x_host: torch.Tensor, lengths_host: Optional[torch.Tensor], y_host: torch.Tensor,
                               lengths_out_host: Optional[torch.Tensor], slot: int, copy_in_stream, copy_out_stream,
                               device=None) -> None:
        """Like forward_host, but the H2D copy is enqueued on ``copy_in_stream``, the forward on the current stream and the
        D2H copy on ``copy_out_stream``; the library orders them (and successive uses of the same ``slot``, which owns the
        device staging workspace) with events.  Rotate more slots than compute streams and the compute streams never wait
        for a copy.  The result is valid once ``copy_out_stream`` has drained."""
        self._host_call(x_host, lengths_host, y_host, lengths_out_host, ("host-slot", int(slot)), device, copy_in_stream,
                        copy_out_stream)


class CapturedEncoderForward:
    """One Encoder.forward captured in a CUDA graph (Encoder.capture)."""

    def __init__(self, graph, x, lengths, y, out_lengths, workspace) -> None:
        self.graph, self.x, self.lengths, self.y, self.out_lengths = graph, x, lengths, y, out_lengths
        self.workspace = workspace                    # scratch the captured kernels use; lives as long as the graph

    def replay(self):
        self.graph.replay()
        return self.y, self.out_lengths


class Decoder(_Native):
    """CTC head, model/modules/decoder.py:7-27: LSTM over packed sequences -> Swish -> BatchNorm1d -> Linear.  The step
    right after the encoder path, on the device through cfm_decoder_forward: the input and
    vocabulary projections are GEMMs, the recurrence one persistent cooperative kernel per layer (decoder.cu).  The
    sub-modules are parameter containers with the reference's names (``lstm.weight_ih_l0`` ..., ``norm.*``,
    ``linear.*``).  Unlike the reference there is no ``lengths.cpu()`` host synchronisation (decoder.py:17): the kernel
    reads the lengths on the device; like the reference's ``pad_packed_sequence`` the output keeps all T frames when
    ``max(lengths) == T``, which the encoder's collate contract guarantees."""
    _KIND = _lib.KIND_DECODER

    def __init__(self, vocab_size: int, d_model: int, hidden_dim: int, n_layers: int) -> None:
        super().__init__()
        self.lstm = nn.LSTM(input_size=d_model, hidden_size=hidden_dim, num_layers=n_layers, batch_first=True)
        self.activation = Swish()
        self.norm = nn.BatchNorm1d(num_features=hidden_dim)
        self.linear = nn.Linear(in_features=hidden_dim, out_features=vocab_size)
        self._dims = (vocab_size, d_model, hidden_dim, n_layers)

    def _config(self) -> Config:
        v, d, h, n = self._dims
        return Config(0, 0, d, 1, 1, v, h, n)

    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.training:
            # Training mode: the head is outside the path and is differentiated by autograd, so
            # it runs the reference's own torch ops (cuDNN LSTM, batch-statistics BatchNorm), decoder.py:15-27.
            packed = lengths is not None
            if packed:
                x = nn.utils.rnn.pack_padded_sequence(x, lengths.cpu(), batch_first=True, enforce_sorted=True)
            x, _ = self.lstm(x)
            if packed:
                x, _ = nn.utils.rnn.pad_packed_sequence(x, batch_first=True)
            x = x * torch.sigmoid(x)
            return self.linear(self.norm(x.transpose(1, 2)).transpose(1, 2))
        xin = _dev_f32(x, "x")
        B, T, _ = xin.shape
        lin = None if lengths is None else lengths.to(device=xin.device, dtype=torch.int64).contiguous()
        y = torch.empty((B, T, self._dims[0]), device=xin.device, dtype=torch.float32)
        with _on_device(xin.device):
            h = self._native()
            ws = self._workspace(h, B, T, xin.device)
            check(lib().cfm_decoder_forward(h, xin.data_ptr(), _ptr(lin), y.data_ptr(), B, T, ws.data_ptr(), ws.numel(),
                                            _stream()))
        return y


class Conformer(nn.Module):
    """model/conformer.py:9-27"""

    def __init__(self, vocab_size: int, n_mel_channels: int = 80, n_conformer_blocks: int = 16, d_model: int = 256,
                 n_heads: int = 4, kernel_size: int = 31, lstm_hidden_dim: int = 640, n_lstm_layers: int = 1,
                 dropout_rate: float = 0.0) -> None:
        super().__init__()
        self.encoder = Encoder(n_mel_channels=n_mel_channels, n_blocks=n_conformer_blocks, d_model=d_model,
                               n_heads=n_heads, kernel_size=kernel_size, dropout_rate=dropout_rate)
        self.decoder = Decoder(vocab_size=vocab_size, d_model=d_model, hidden_dim=lstm_hidden_dim, n_layers=n_lstm_layers)

    def set_compute(self, mode: Optional[str]) -> "Conformer":
        """'fp32', 'bf16' or None (= follow torch.autocast) for the encoder and the CTC head."""
        self.encoder.set_compute(mode)
        self.decoder.set_compute(mode)
        return self

    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        x, lengths = self.encoder(x, lengths)
        x = self.decoder(x, lengths)
        return x, lengths
