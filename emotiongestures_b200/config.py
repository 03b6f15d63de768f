"""Geometry of the generator-inference hot path.

The reference never writes these down in one place: they are literals at the
constructor call (test_emotion_gesture_diversity_iterative.py:135) and in the
smoke block of Full_model/Models.py:516-540.  Everything the kernels need is
derived here once, from the live module, never from flags.
"""
from __future__ import annotations

import dataclasses
import math

SAMPLE_RATE = 16000
N_FFT = 1024          # utils/data_utils.py:36
HOP = 512             # utils/data_utils.py:36
N_BINS = N_FFT // 2 + 1
N_MELS = 128          # librosa default; model/hierarchy_net.py:21
F_MAX = 8000.0
PREEMPH_COEF = 0.97   # model/utils.py:24
LOG_EPS = 1e-6        # model/ResNetSE34V2.py:96
IN_EPS = 1e-5         # torch.nn.InstanceNorm1d default
DB_AMIN = 1e-10       # librosa.power_to_db default
DB_TOP = 80.0         # librosa.power_to_db default

LOGMEL_DB = 0         # F4a: power_to_db(ref=max), utils/data_utils.py:37
LOGMEL_LOG_IN = 1     # F4b: log(x+1e-6) + InstanceNorm1d, model/ResNetSE34V2.py:96-98
LOGMEL_FP16_STORAGE = 0x100   # OR-ed into LOGMEL_DB: the astype('float16') storage cast of utils/data_utils.py:38
# The features the reference's checkpoints were trained on and its loaders deliver (utils/data_utils.py:35-39, re-read as
# fp32 at data_loader/lmdb_data_loader_expressive.py:204): no pre-emphasis, power_to_db(ref=max), fp16 storage rounding.
# This is the default of every raw-audio entry point.  One documented difference remains: the reference takes ref=max
# over the whole recording it extracted features from and then slices clips out of it, while a per-clip front-end
# takes the max over the clip handed in, so a clip quieter than its recording's peak comes out shifted by a constant
# (and floored 80 dB below its own peak instead of the recording's).  Callers that have the recording's peak can
# add `10*log10(clip_max/recording_max)` to the result.
LOGMEL_REFERENCE = LOGMEL_DB | LOGMEL_FP16_STORAGE


def spectrogram_length(n_frames: int, fps: int) -> int:
    """utils/data_utils.py:42-44."""
    return int(round((n_frames / fps * SAMPLE_RATE - N_FFT) / HOP + 1))


def audio_length(n_frames: int, fps: int) -> int:
    """data_loader/lmdb_data_loader_expressive.py:95."""
    return int(round(n_frames / fps * SAMPLE_RATE))


@dataclasses.dataclass(frozen=True)
class GeneratorConfig:
    frames: int = 34
    prior_frames: int = 4
    pose_dim: int = 126
    d_model: int = 256
    d_inner: int = 1024
    n_layers: int = 3
    n_head: int = 8
    d_k: int = 64
    d_v: int = 64
    n_mels: int = N_MELS
    spec_w: int = 70            # spectrogram columns fed to the trunk
    n_audio: int = 36267        # raw samples per clip
    n_position: int = 60
    # text encoder (dead w.r.t. poses, Full_model/Models.py:400,427)
    n_words: int = 100
    wordembed_dim: int = 300
    tcn_hidden: int = 300
    tcn_layers: int = 3
    text_len: int = 60

    @property
    def trunk_hw(self):
        """(H, W) after layer3: two stride-2 3x3 pad-1 convs."""
        h = (self.n_mels + 1) // 2
        h = (h + 1) // 2
        w = (self.spec_w + 1) // 2
        w = (w + 1) // 2
        return h, w

    @property
    def fc1_in(self) -> int:
        h, w = self.trunk_hw
        return h * w

    @property
    def n_stft_frames(self) -> int:
        return 1 + self.n_audio // HOP

    def validate(self) -> None:
        if self.spec_w > self.n_stft_frames:
            raise ValueError("spec_w exceeds the STFT frame count of n_audio samples")
        if self.frames > self.n_position:
            raise ValueError("frames exceeds n_position (pos_table rows)")
        if self.n_head * self.d_k % 8 or self.d_model % 32:
            raise ValueError("d_model must be a multiple of 32 and n_head*d_k of 8")


# 34-frame TED-Emotion shape (Full_model/Models.py:522-533)
TED = GeneratorConfig()
# 60-frame BEAT shape (test_emotion_gesture_diversity_iterative.py:135,347-370)
BEAT = GeneratorConfig(frames=60, prior_frames=10, pose_dim=282, d_model=512, d_inner=2048,
                       spec_w=124, n_audio=64000)

assert spectrogram_length(34, 15) == 70 and audio_length(34, 15) == 36267
assert spectrogram_length(60, 15) == 124 and audio_length(60, 15) == 64000
assert TED.fc1_in == 32 * 18 and BEAT.fc1_in == 32 * 31
assert math.isclose(TED.n_audio / SAMPLE_RATE, 2.2667, abs_tol=1e-3)
