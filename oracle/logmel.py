"""Oracle (test infrastructure): Whisper log-mel front end, numpy float64/float32.

Restates the third-party arithmetic the reference reaches from
``ref:ultravox/model/ultravox_processing.py:295-303``:

* ``transformers.WhisperFeatureExtractor.__call__``
  (``hf:models/whisper/feature_extraction_whisper.py:189-342``): zero-pad the batch to
  the longest clip rounded up to a multiple of ``hop_length``; frame mask =
  sample mask ``[:, ::hop]``.
* ``_torch_extract_fbank_features`` (``:135-164``): ``torch.stft(n_fft=400, hop=160,
  periodic hann, center=True -> reflect pad 200)``, drop the last frame, ``|.|**2``,
  ``mel_filters.T @``, ``log10(clamp(1e-10))``, per-clip ``max - 8`` floor, ``(x+4)/4``.
* ``mel_filter_bank(..., norm="slaney", mel_scale="slaney")``
  (``hf:audio_utils.py:263-297,356-375,453-545``).

Transformers 5.5.0 is what is installed; the reference pins 4.51.3 - same formulas.
Pinned in ``tests/test_oracle_cpu.py`` against ``WhisperFeatureExtractor`` itself and
against ``tests/golden/logmel_*.npz``.
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
HOP = 160
SAMPLE_RATE = 16000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    with np.errstate(divide="ignore", invalid="ignore"):
        mels = np.where(hi, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, mels)
    return mels


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    f = 200.0 * m / 3.0
    hi = m >= 15.0
    return np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def mel_filter_bank(n_mels: int, n_freqs: int = N_FFT // 2 + 1, f_max: float = 8000.0,
                    sr: int = SAMPLE_RATE) -> np.ndarray:
    """float64 [n_freqs, n_mels] slaney-scale, slaney-normalised triangular filters."""
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(f_max), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_freqs)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels]))[None, :]
    return fb


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))


def pad_batch(waves: list[np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
    """Zero-pad to the longest clip rounded up to a multiple of HOP.

    Returns (padded [B, L] float32, frame_lens [B] int64) where frame_lens[i] is the number of ones
    in the frame mask ``sample_mask[:, ::HOP]`` (= ceil(len_i / HOP))."""
    lens = np.array([len(w) for w in waves], dtype=np.int64)
    longest = int(lens.max()) if len(lens) else 0
    padded_len = -(-longest // HOP) * HOP
    out = np.zeros((len(waves), padded_len), dtype=np.float32)
    for i, w in enumerate(waves):
        out[i, : len(w)] = np.asarray(w, dtype=np.float32)
    frame_lens = -(-lens // HOP)
    return out, frame_lens


def log_mel(padded: np.ndarray, n_mels: int, dtype=np.float64) -> np.ndarray:
    """[B, L] -> [B, n_mels, L // HOP] float32 (computed in ``dtype``)."""
    x = np.asarray(padded, dtype=dtype)
    B, L = x.shape
    n_frames = L // HOP  # torch.stft gives 1 + L//HOP frames; the last one is dropped
    xp = np.pad(x, ((0, 0), (N_FFT // 2, N_FFT // 2)), mode="reflect")
    idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = xp[:, idx] * hann_periodic().astype(dtype)[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)
    power = (spec.real ** 2 + spec.imag ** 2).astype(dtype)          # [B, T, 201]
    fb = mel_filter_bank(n_mels).astype(np.float32).astype(dtype)     # the extractor casts filters to f32
    mel = np.einsum("btf,fm->bmt", power, fb)
    logm = np.log10(np.maximum(mel, 1e-10))
    mx = logm.reshape(B, -1).max(axis=1)[:, None, None]
    logm = np.maximum(logm, mx - 8.0)
    return ((logm + 4.0) / 4.0).astype(np.float32)
