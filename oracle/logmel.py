"""ORACLE (test infrastructure, not product code): float64 log-mel front-end.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.  PARITY UNPINNED for F2-F4a: the reference
computes these features offline with an un-vendored, un-versioned `librosa`
(requirements.txt:5,16) and ships neither tests nor golden vectors for it, so
this file restates the published algorithm from the reference's call sites.
F1, the F4b normalisation and F5 ARE pinned, against the real
model.utils.PreEmphasis, torch.nn.InstanceNorm1d(128) and
utils.data_utils.make_audio_fixed_length (oracle/make_golden_frontend.py ->
tests/golden/frontend_pins.npz).

  F1  pre-emphasis            model/utils.py:22-38
  F2  STFT n_fft=1024 hop=512 utils/data_utils.py:36 (librosa.feature.melspectrogram, power=2)
  F3  mel projection          same call; librosa defaults: 128 Slaney mels, 0..8000 Hz, Slaney norm
  F4a power_to_db(ref=max)    utils/data_utils.py:37
  F4b log(x+1e-6)+InstanceNorm1d  model/ResNetSE34V2.py:94-98 (commented-out recipe)
  F5  fixed-length audio      utils/data_utils.py:69-75

The Slaney filterbank below is cross-checked against
torchaudio.functional.melscale_fbanks in tests/test_oracle_logmel.py.
"""
from __future__ import annotations

import numpy as np

SR = 16000
N_FFT = 1024
HOP = 512
N_MELS = 128
F_MAX = 8000.0


def make_audio_fixed_length(audio: np.ndarray, n: int) -> np.ndarray:
    """utils/data_utils.py:69-75: symmetric-pad at the end, or crop."""
    pad = n - len(audio)
    if pad > 0:
        return np.pad(audio, (0, pad), mode="symmetric")
    return audio[:n]


def preemphasis(x: np.ndarray, coef: float = 0.97) -> np.ndarray:
    """model/utils.py:33-38: reflect-pad one sample on the left, then y[t] = x[t] - coef*x[t-1].  The module keeps
    its filter as a torch.FloatTensor (:29-31), so the coefficient in force is float32(0.97) = 0.9700000286...
    (pinned against the real module by oracle/make_golden_frontend.py)."""
    x = np.asarray(x, dtype=np.float64)
    coef = float(np.float32(coef))
    prev = np.concatenate([x[..., 1:2], x[..., :-1]], axis=-1)
    return x - coef * prev


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), the window librosa.stft uses."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int = SR, n_fft: int = N_FFT, n_mels: int = N_MELS,
                   fmin: float = 0.0, fmax: float = F_MAX) -> np.ndarray:
    """librosa.filters.mel defaults (htk=False, norm='slaney'): (n_mels, 1+n_fft//2) float64."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_bins))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return w * enorm[:, None]


def stft_power(x: np.ndarray, n_cols: int | None = None, pad_mode: str = "constant") -> np.ndarray:
    """Centred STFT power spectrum, (..., 513, T) with T = 1 + N // 512 (first n_cols kept).

    pad_mode: 'constant' (librosa >= 0.10) or 'reflect' (older librosa); only the
    first and last frames differ.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    t_all = 1 + n // HOP
    t = t_all if n_cols is None else n_cols
    pad = [(0, 0)] * (x.ndim - 1) + [(N_FFT // 2, N_FFT // 2)]
    xp = np.pad(x, pad, mode=pad_mode)
    idx = np.arange(t)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = xp[..., idx] * hann_periodic()          # (..., T, 1024)
    spec = np.fft.rfft(frames, axis=-1)              # (..., T, 513)
    power = spec.real ** 2 + spec.imag ** 2
    return np.swapaxes(power, -1, -2)                # (..., 513, T)


def logmel(audio: np.ndarray, n_cols: int, mode: str = "log_in", preemph: bool = True,
           pad_mode: str = "constant") -> np.ndarray:
    """(B, N) audio -> (B, 128, n_cols) float64 log-mel.

    mode 'db'     : F4a, 10*log10(max(M,1e-10)) - 10*log10(max M), floored at -80 dB; the
                    max is per clip (the array handed to power_to_db).
    mode 'log_in' : F4b, log(M + 1e-6) then InstanceNorm1d (per clip and mel bin over the
                    n_cols kept columns, biased variance, eps 1e-5, no affine).
    """
    x = np.atleast_2d(np.asarray(audio, dtype=np.float64))
    if preemph:
        x = preemphasis(x)
    power = stft_power(x, n_cols, pad_mode)
    mel = np.einsum("mk,bkt->bmt", mel_filterbank(), power)
    if mode == "db":
        ref = mel.max(axis=(1, 2), keepdims=True)
        db = 10.0 * np.log10(np.maximum(mel, 1e-10)) - 10.0 * np.log10(np.maximum(ref, 1e-10))
        return np.maximum(db, db.max(axis=(1, 2), keepdims=True) - 80.0)
    if mode == "log_in":
        lg = np.log(mel + 1e-6)
        mu = lg.mean(axis=2, keepdims=True)
        var = lg.var(axis=2, keepdims=True)
        return (lg - mu) / np.sqrt(var + 1e-5)
    raise ValueError(mode)
