"""Host side of the sr_inpainting front end (A12): the mel filterbank handed to ``aldm_stft_mel`` and the waveform
preparation of ``wav_to_fbank`` / ``read_wav_file`` (utilities/audio/tools.py:21-104).

The filterbank restates what ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` (called at utilities/audio/
stft.py:145-147; librosa==0.9.2 defaults htk=False, norm="slaney") evaluates: triangular filters on the Slaney mel
scale with Slaney area normalisation.  tests/test_frontend_cpu.py pins it against golden values produced by two
independent published restatements of that algorithm (torchaudio ``melscale_fbanks`` and transformers
``mel_filter_bank`` in their librosa-compatible modes; tests/golden/make_mel_golden.py)."""
from __future__ import annotations

import contextlib
import wave
from typing import Optional, Tuple

import numpy as np
import torch

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    m = f / _F_SP
    big = f >= _MIN_LOG_HZ
    m[big] = _MIN_LOG_MEL + np.log(f[big] / _MIN_LOG_HZ) / _LOGSTEP
    return m


def mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    f = _F_SP * m
    big = m >= _MIN_LOG_MEL
    f[big] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m[big] - _MIN_LOG_MEL))
    return f


def mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: Optional[float] = None) -> torch.Tensor:
    """-> float32 [n_mels, 1 + n_fft // 2]"""
    fmax = sr / 2.0 if fmax is None else fmax
    fft_f = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin)[0], hz_to_mel(fmax)[0], n_mels + 2))
    width = np.diff(edges)
    ramp = edges[:, None] - fft_f[None, :]
    lower = -ramp[:-2] / width[:-1, None]
    upper = ramp[2:] / width[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return torch.from_numpy(w.astype(np.float32))


def mel_basis_for(cfg: dict) -> torch.Tensor:
    v = cfg["vocoder"]
    return mel_basis(v["sampling_rate"], v["n_fft"], v["num_mels"], v["fmin"], v["fmax"])


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """PCM WAV -> (mono float32 in [-1, 1], sample rate); first channel as ``torchaudio.load(...)[0]`` (tools.py:31-33)."""
    with contextlib.closing(wave.open(path, "rb")) as f:
        n, ch, sw, sr = f.getnframes(), f.getnchannels(), f.getsampwidth(), f.getframerate()
        raw = f.readframes(n)
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported sample width {sw}")
    return x.reshape(-1, ch)[:, 0].copy(), sr


def prepare_waveform(x: np.ndarray, sr: int, target_sr: int, segment_length: int) -> np.ndarray:
    """read_wav_file (tools.py:28-40): resample, remove DC, peak-normalise to 0.5, pad / crop to ``segment_length``,
    re-normalise.  Returns [1, segment_length] float32."""
    x = np.asarray(x, dtype=np.float32)
    if sr != target_sr:
        import torchaudio.functional as AF                                         # tools.py:32
        x = AF.resample(torch.from_numpy(x)[None], orig_freq=sr, new_freq=target_sr)[0].numpy()
    x = x - np.mean(x)
    x = x / (np.max(np.abs(x)) + 1e-8) * 0.5                                       # normalize_wav (tools.py:21-25)
    assert x.shape[-1] > 100, "Waveform is too short, %s" % x.shape[-1]             # pad_wav (tools.py:8-18)
    if x.shape[-1] > segment_length:
        x = x[:segment_length]
    elif x.shape[-1] < segment_length:
        x = np.concatenate([x, np.zeros(segment_length - x.shape[-1], dtype=x.dtype)])
    x = x / np.max(np.abs(x)) * 0.5                                                # tools.py:37-38
    return x[None].astype(np.float32)
