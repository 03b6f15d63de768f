"""ORACLE (test infrastructure): platform-stable synthetic weights and inputs.

Weights are a pure function of (state_dict key, shape, seed) through numpy's
PCG64, so the build container (where the real reference is importable and the
golden vectors are made) and the GPU box (where it is not) construct identical
tensors without shipping 60 MB of parameters.  Distributions follow the
reference's init (Full_model/Models.py:381-383 xavier_uniform over dim>1) with
BatchNorm/LayerNorm statistics randomised as SURVEY.md §8(d) prescribes.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

_NORM_MARKERS = (".bn1.", ".bn2.", ".downsample.1.", "layer_norm.")


def _canonical(key: str) -> str:
    # Full_model/tcn.py:31-32 registers conv1/conv2 again as net.0/net.4
    return key.replace(".net.0.", ".conv1.").replace(".net.4.", ".conv2.")


def _rng(key: str, seed: int) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(_canonical(key).encode())])


def synth_tensor(key: str, shape, seed: int, dtype=torch.float32, norm: bool = False) -> torch.Tensor:
    r = _rng(key, seed)
    shape = tuple(shape)
    if key.endswith("num_batches_tracked"):
        return torch.zeros((), dtype=torch.int64)
    is_norm = norm or any(m in key for m in _NORM_MARKERS)
    if key.endswith("running_var"):
        a = r.uniform(0.5, 1.5, shape)
    elif key.endswith("running_mean"):
        a = r.normal(0.0, 0.1, shape)
    elif is_norm and key.endswith(".weight"):
        a = r.uniform(0.5, 1.5, shape)
    elif is_norm and key.endswith(".bias"):
        a = r.normal(0.0, 0.1, shape)
    elif key.endswith("weight_g"):
        a = r.uniform(0.5, 1.5, shape)
    elif len(shape) > 1:
        recf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
        fan_in, fan_out = shape[1] * recf, shape[0] * recf
        bound = np.sqrt(6.0 / (fan_in + fan_out))
        a = r.uniform(-bound, bound, shape)
    else:
        a = r.uniform(-0.05, 0.05, shape)
    if "temporal_chunk_encoder.2." in key:
        # Models_memory.py:287-289 squares this encoding inside a softmax; at xavier scale the scores are +-100 and the
        # softmax is a constant one-hot, which would leave the temporal memory arithmetic untested
        a = a * 0.1
    return torch.tensor(a, dtype=dtype)


def synth_state_dict(template: dict, seed: int = 0) -> dict:
    """Fill every tensor of `template` (a state_dict giving keys/shapes); the sinusoid
    position tables are kept as the module built them."""
    out = {}
    for k, v in template.items():
        if "pos_table" in k:
            out[k] = v.clone()
        else:
            # a parameter whose module also owns a running_mean is a BatchNorm weight / bias
            norm = (k.rsplit(".", 1)[0] + ".running_mean") in template
            out[k] = synth_tensor(k, v.shape, seed, norm=norm)
    return out


def synth_audio(n_clips: int, n_audio: int, seed: int = 0, kind: str = "noise") -> np.ndarray:
    """noise-like 16 kHz audio, 0.1*randn clamped to [-1, 1] (float32)."""
    r = np.random.default_rng([seed, 0xA0D10])
    if kind == "noise":
        a = 0.1 * r.standard_normal((n_clips, n_audio))
    elif kind == "harmonic":
        t = np.arange(n_audio) / 16000.0
        f0 = r.uniform(90, 250, (n_clips, 1))
        a = sum(np.sin(2 * np.pi * f0 * h * t) / h for h in range(1, 12)) * 0.05
        a = a + 1e-3 * r.standard_normal((n_clips, n_audio))
    else:
        raise ValueError(kind)
    return np.clip(a, -1.0, 1.0).astype(np.float32)


def synth_prior(n_clips: int, prior_frames: int, pose_dim: int, seed: int = 0) -> np.ndarray:
    r = np.random.default_rng([seed, 0x9051])
    return r.standard_normal((n_clips, prior_frames, pose_dim)).astype(np.float32)


def synth_spec(n_clips: int, n_mels: int, w: int, seed: int = 0) -> np.ndarray:
    """A stand-in spectrogram with InstanceNorm-like statistics."""
    r = np.random.default_rng([seed, 0x59EC])
    return r.standard_normal((n_clips, n_mels, w)).astype(np.float32)


def synth_emotion(n_clips: int, frames: int, d_model: int, seed: int = 0) -> np.ndarray:
    r = np.random.default_rng([seed, 0xE307])
    return (0.5 * r.standard_normal((n_clips, frames, d_model))).astype(np.float32)
