"""Log-mel front-end on the device: the step right before the encoder path.

Mirror of the audio half of the reference's ``ConformerProcessor`` (processing/processor.py): ``mel_spectrogram``
(:155-158, torchaudio ``MelSpectrogram`` configured at :53-63) and the padding / length rule of ``__call__`` (:373-394).
Same constructor argument names and defaults; the tokenizer half of the processor is string work and stays in the
reference.  Everything runs through the C ABI (cfm_logmel_*): there is no CPU path and no torchaudio call.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib
from .modules import _on_device, _stream


def _hz_to_mel(f: float, scale: str) -> float:
    if scale == "htk":
        return 2595.0 * math.log10(1.0 + f / 700.0)
    if f >= 1000.0:
        return 15.0 + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return f * 3.0 / 200.0


def _mel_filterbank_f32(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int, norm: Optional[str],
                        mel_scale: str) -> torch.Tensor:
    """The module's constant filterbank buffer [n_freqs, n_mels], built in fp32 on the host the way
    torchaudio.functional.melscale_fbanks publishes it (triangles between n_mels + 2 points equally spaced on the mel
    scale, slaney area normalisation), so that the filters carry the same fp32 construction rounding as the reference's."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2)
    if mel_scale == "htk":
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    else:
        logstep = math.log(6.4) / 27.0
        f_pts = torch.where(m_pts >= 15.0, 1000.0 * torch.exp(logstep * (m_pts - 15.0)), (200.0 / 3.0) * m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb.contiguous()


class LogMelFrontEnd:
    """``ConformerProcessor(sample_rate, n_fft, win_length, hop_length, n_mels, fmin, fmax, norm, mel_scale)`` - audio part."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, win_length: int = 400, hop_length: int = 160,
                 n_mels: int = 80, fmin: float = 0.0, fmax: float = 8000.0, norm: Optional[str] = "slaney",
                 mel_scale: str = "slaney") -> None:
        assert mel_scale in ("htk", "slaney"), "Invalid Mel Scale, Only HTK or Slaney"        # processor.py:45
        assert norm in (None, "slaney"), "Invalid Norm, we only support Slaney Norm"          # processor.py:46-47
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        self._cfg = _lib.LogMelConfig(sample_rate, n_fft, win_length, hop_length, n_mels, float(fmin), float(fmax),
                                      1 if norm == "slaney" else 0, 1 if mel_scale == "slaney" else 0, 1e-5)
        # constant buffers of the module, as torchaudio's MelSpectrogram holds them (spectrogram.window, mel_scale.fb)
        win = torch.hann_window(win_length)
        left = (n_fft - win_length) // 2
        self.window = torch.zeros(n_fft)
        self.window[left:left + win_length] = win
        self.fb = _mel_filterbank_f32(n_fft // 2 + 1, float(fmin), float(fmax) if fmax else sample_rate / 2.0, n_mels,
                                      sample_rate, norm, mel_scale)
        self._handles, self._ws = {}, {}

    def __del__(self):
        try:
            for h in self._handles.values():
                lib().cfm_logmel_destroy(h)
        except Exception:
            pass

    def _handle(self, device) -> C.c_void_p:
        h = self._handles.get(device)
        if h is None:
            h = C.c_void_p()
            check(lib().cfm_logmel_create(C.byref(self._cfg), self.window.data_ptr(), self.fb.data_ptr(), C.byref(h)))
            self._handles[device] = h
        return h

    def mel_spectrogram(self, signal: torch.Tensor) -> torch.Tensor:
        """processor.py:155-158: signal [B, samples] (or [samples]) on a CUDA device -> log-mel [B, n_mels, samples // hop + 1]."""
        if not signal.is_cuda:
            raise RuntimeError("conformer_b200: signal must be a CUDA tensor (there is no CPU path)")
        squeeze = signal.dim() == 1
        x = signal.reshape(1, -1) if squeeze else signal
        x = x.to(torch.float32).contiguous()
        B, L = x.shape
        out = torch.empty((B, self.n_mels, L // self.hop_length + 1), device=x.device, dtype=torch.float32)
        with _on_device(x.device):
            h = self._handle(x.device)
            need = lib().cfm_logmel_workspace_bytes(h, B, L)
            if need < 0:
                check(1)
            stream = torch.cuda.current_stream(x.device)
            key = (x.device, stream.cuda_stream)
            ws = self._ws.get(key)
            if ws is None or ws.numel() < need:
                if ws is not None:
                    ws.record_stream(stream)
                ws = torch.empty(int(need), device=x.device, dtype=torch.uint8)
                self._ws[key] = ws
            check(lib().cfm_logmel_forward(h, x.data_ptr(), out.data_ptr(), B, L, ws.data_ptr(), ws.numel(), _stream()))
        return out[0] if squeeze else out

    def frame_lengths(self, sample_lengths: torch.Tensor) -> torch.Tensor:
        """processor.py:392: samples // hop_length + 1, on the device."""
        s = sample_lengths.to(torch.int64).contiguous()
        if not s.is_cuda:
            raise RuntimeError("conformer_b200: lengths must be a CUDA tensor (there is no CPU path)")
        out = torch.empty_like(s)
        with _on_device(s.device):
            check(lib().cfm_logmel_lengths(s.data_ptr(), out.data_ptr(), s.numel(), self.hop_length, _stream()))
        return out

    def __call__(self, audios: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """processor.py:373-394 without augmentation: zero-pad to the longest utterance, log-mel, frame lengths."""
        device = audios[0].device
        lengths = torch.tensor([len(a) for a in audios], dtype=torch.int64)
        batch = torch.zeros((len(audios), int(lengths.max())), device=device, dtype=torch.float32)
        for i, a in enumerate(audios):
            batch[i, :len(a)] = a
        return self.mel_spectrogram(batch), self.frame_lengths(lengths.to(device))
