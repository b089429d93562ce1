"""CPU restatement of the TacotronSTFT mel front-end (TEST INFRASTRUCTURE).

Follows /root/reference/audio_processing.py:96-134 (TacotronSTFT), :172-235 (STFT.__init__ /
transform) and :78-84 (dynamic_range_compression).

Third-party arithmetic not in the reference tree: ``librosa.filters.mel`` (pinned
librosa==0.6.3 in requirements.txt:4 / 0.8.0 in Dockerfile:6; call site audio_processing.py:104-105)
and ``librosa.util.pad_center``.  Their published algorithms are restated in ``slaney_mel_basis``
and ``pad_center``.  The reference holds no test or golden vector for the filterbank:
PARITY UNPINNED at that boundary (the restatement is the spec).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def hz_to_mel_slaney(f):
    """librosa.core.hz_to_mel(htk=False): linear below 1 kHz (200/3 Hz per mel), log above."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int = 80, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1): triangular filters on
    the Slaney mel scale, each scaled by 2/(f[i+2]-f[i]).  Returns float32 [n_mels, 1+n_fft//2]."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, sr / 2.0, n_bins)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def pad_center(data: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center for 1-D data."""
    n = data.shape[-1]
    lpad = (size - n) // 2
    return np.pad(data, (lpad, size - n - lpad), mode="constant")


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) (audio_processing.py:196)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def forward_basis(filter_length: int = 1024, win_length: int = 1024) -> np.ndarray:
    """Windowed real/imag DFT rows, float32 [2*(N/2+1), N] (audio_processing.py:183-204)."""
    fb = np.fft.fft(np.eye(filter_length))
    cutoff = filter_length // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff]), np.imag(fb[:cutoff])])
    basis = torch.FloatTensor(fb)
    window = torch.from_numpy(pad_center(hann_periodic(win_length), filter_length)).float()
    return (basis * window).numpy()


def mel_spectrogram(y: torch.Tensor, filter_length=1024, hop_length=256, win_length=1024,
                    n_mel_channels=80, sampling_rate=22050, mel_fmin=0.0, mel_fmax=8000.0,
                    mel_basis: np.ndarray | None = None) -> torch.Tensor:
    """TacotronSTFT.mel_spectrogram (audio_processing.py:117-134): y [B,N] in [-1,1] -> [B,80,1+N//hop]."""
    assert float(y.min()) >= -1 and float(y.max()) <= 1
    basis = torch.from_numpy(forward_basis(filter_length, win_length))[:, None, :]
    x = F.pad(y[:, None, None, :], (filter_length // 2, filter_length // 2, 0, 0), mode="reflect")[:, 0]
    ft = F.conv1d(x, basis, stride=hop_length)                          # :221-225
    cutoff = filter_length // 2 + 1
    mag = torch.sqrt(ft[:, :cutoff] ** 2 + ft[:, cutoff:] ** 2)        # :231
    if mel_basis is None:
        mel_basis = slaney_mel_basis(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)
    mel = torch.matmul(torch.from_numpy(mel_basis), mag)               # :132
    return torch.log(torch.clamp(mel, min=1e-5))                       # :84


def stft_transform(y: torch.Tensor, filter_length=1024, hop_length=256, win_length=1024):
    """STFT.transform (audio_processing.py:207-235): y [B,N] -> (magnitude, phase) each [B, N/2+1, 1+N//hop]."""
    basis = torch.from_numpy(forward_basis(filter_length, win_length))[:, None, :]
    x = F.pad(y[:, None, None, :], (filter_length // 2, filter_length // 2, 0, 0), mode="reflect")[:, 0]
    ft = F.conv1d(x, basis, stride=hop_length)
    cutoff = filter_length // 2 + 1
    re, im = ft[:, :cutoff], ft[:, cutoff:]
    return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)
