"""CPU oracle for the log-mel front-end.  TEST INFRASTRUCTURE ONLY (same rules as conformer_oracle.py).

The reference computes its encoder input as ``log(clamp(MelSpectrogram(...)(signal), min=1e-5))``
(processing/processor.py:53-63,155-158).  The arithmetic lives in a third-party dependency that is not part of
/root/reference: **torchaudio** (requirements.txt:12 pins 2.1.0; this image has 2.11.0 with the same transform).  This
file restates torchaudio's published algorithm in plain torch/numpy:

  * ``torch.stft(center=True, pad_mode="reflect", window=hann_window(win_length) [periodic], onesided=True)``
    -> frames t = 0 .. L // hop, frame t covers samples [t*hop - n_fft/2, t*hop + n_fft/2) of the reflect-padded signal;
  * ``Spectrogram(power=2.0, normalized=False)`` -> |X|^2;
  * ``torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm, mel_scale)``:
    triangles on ``linspace(0, sr // 2, n_freqs)`` between ``n_mels + 2`` points equally spaced on the mel scale
    (slaney: linear below 1 kHz, log above with step log(6.4)/27), slaney norm = 2 / (f[m+2] - f[m]);
  * mel = fb^T . power, then the reference's log-clamp.

Parity status: PINNED against torchaudio itself.  tests/golden/make_golden.py runs
``torchaudio.transforms.MelSpectrogram`` with the reference's constructor arguments (processor.py:53-63) on seeded
waveforms and commits its output (tests/golden/logmel.npz); tests/test_oracle_golden.py checks this file against it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch


def hz_to_mel(f: float, scale: str) -> float:
    if scale == "htk":
        return 2595.0 * math.log10(1.0 + f / 700.0)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    if f >= min_log_hz:
        return min_log_hz / f_sp + math.log(f / min_log_hz) / (math.log(6.4) / 27.0)
    return f / f_sp


def mel_to_hz(m: torch.Tensor, scale: str) -> torch.Tensor:
    if scale == "htk":
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
    return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int, norm: Optional[str],
                   mel_scale: str, dtype=torch.float64) -> torch.Tensor:
    """[n_freqs, n_mels] triangular filters."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=dtype)
    m_pts = torch.linspace(hz_to_mel(f_min, mel_scale), hz_to_mel(f_max, mel_scale), n_mels + 2, dtype=dtype)
    f_pts = mel_to_hz(m_pts, mel_scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0.0)
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb


def log_mel(signal: torch.Tensor, sample_rate: int = 16000, n_fft: int = 400, win_length: int = 400, hop_length: int = 160,
            n_mels: int = 80, f_min: float = 0.0, f_max: float = 8000.0, norm: Optional[str] = "slaney",
            mel_scale: str = "slaney", floor: float = 1e-5) -> torch.Tensor:
    """signal [B, L] -> [B, n_mels, L // hop + 1] in the dtype of ``signal`` (float64 gives the truth)."""
    dt = signal.dtype
    B, L = signal.shape
    half = n_fft // 2
    idx = torch.arange(-half, L + half)
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx >= L, 2 * (L - 1) - idx, idx)                       # reflect padding
    padded = signal[:, idx]
    T = L // hop_length + 1
    frames = padded.unfold(1, n_fft, hop_length)[:, :T]                       # [B, T, n_fft]
    n = torch.arange(win_length, dtype=torch.float64)
    win = (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / win_length)).to(dt)      # periodic Hann
    left = (n_fft - win_length) // 2
    window = torch.zeros(n_fft, dtype=dt)
    window[left:left + win_length] = win
    spec = torch.fft.rfft(frames * window, dim=-1)                            # [B, T, n_fft/2+1]
    power = spec.real ** 2 + spec.imag ** 2
    fb = mel_filterbank(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate, norm, mel_scale, dtype=dt)
    mel = power @ fb                                                          # [B, T, n_mels]
    return torch.log(torch.clamp(mel, min=floor)).transpose(1, 2).contiguous()


def synth_waveform(batch: int, samples: int, seed: int = 0) -> torch.Tensor:
    """Speech-like test signal: a few drifting harmonics with an amplitude envelope plus a noise floor, |x| < 1 -
    loud and quiet bands in the same frame, which is what separates an fp32-grade spectrogram from a tf32/bf16 one."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(samples, dtype=torch.float64) / 16000.0
    out = []
    for _ in range(batch):
        f0 = 90.0 + 120.0 * torch.rand(1, generator=g).item()
        x = torch.zeros(samples, dtype=torch.float64)
        for h in range(1, 9):
            drift = 1.0 + 0.05 * torch.sin(2 * math.pi * (0.7 + 0.3 * h) * t)
            x += (0.5 / h) * torch.sin(2 * math.pi * f0 * h * t * drift + 6.28 * torch.rand(1, generator=g).item())
        env = 0.55 + 0.45 * torch.sin(2 * math.pi * 1.3 * t + torch.rand(1, generator=g).item())
        x = x * env * 0.3 + 1e-3 * torch.randn(samples, generator=g, dtype=torch.float64)
        out.append(x)
    return torch.stack(out).to(torch.float32)
