"""Beat-alignment metric of the evaluation loop on the GPU.

`alignment` mirrors model/Beat_score_v2.py's class of the same name — constructor `(sigma, order)`, `load_pose`,
`GAHR`, `calculate_align` with the reference's argument meaning and return shapes — so
test_emotion_gesture_diversity_iterative.py:185,243-248 runs against it unchanged, and adds `score_batch`, which does
the per-clip loop of :243-248 for a whole batch in one kernel launch (`egx_beat_align`) without the poses leaving the
device.

`load_audio` (:58-77) is three librosa onset calls.  librosa is an un-vendored, un-versioned dependency of the
reference and is not part of this image, so that half is NOT restated here (parity would be unpinned): it calls the
same librosa functions when librosa is importable and raises otherwise; `score_batch` takes the onset frames as input.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .config import BEAT


def frames_to_time(frames, sr: int = 22050, hop_length: int = 512):
    """librosa.frames_to_time as the reference calls it (model/Beat_score_v2.py:205: defaults, so 22050 Hz / 512)."""
    return np.asarray(frames, dtype=np.float64) * hop_length / float(sr)


class alignment:                                           # noqa: N801  (the reference's class name)
    def __init__(self, sigma, order, engine=None, device="cuda"):
        self.sigma = sigma
        self.order = order
        self.times = self.oenv = self.S = self.rms = None
        self.pose_data = []
        if engine is None:
            from .engine import Engine
            engine = Engine(BEAT, device)
        self.engine = engine

    # -- audio half: the reference's own librosa calls, when librosa exists -------------------------------------
    def load_audio(self, wave, t_start, without_file=False, sr_audio=16000):
        try:
            import librosa
        except ImportError as e:
            raise RuntimeError("alignment.load_audio is librosa's onset detector (model/Beat_score_v2.py:58-77); librosa is "
                               "not installed — pass onset frames from your own detector to score_batch / "
                               "calculate_align") from e
        short_y = wave[t_start * sr_audio:]
        self.oenv = librosa.onset.onset_strength(y=short_y, sr=sr_audio)
        self.times = librosa.times_like(self.oenv)
        onset_raw = librosa.onset.onset_detect(onset_envelope=self.oenv, backtrack=False)
        onset_bt = librosa.onset.onset_backtrack(onset_raw, self.oenv)
        self.S = np.abs(librosa.stft(y=short_y))
        self.rms = librosa.feature.rms(S=self.S)
        onset_bt_rms = librosa.onset.onset_backtrack(onset_raw, self.rms[0])
        return onset_raw, onset_bt, onset_bt_rms

    # -- pose half: on the device -------------------------------------------------------------------------------
    def _run(self, poses, onsets, lo, hi, pose_fps, want_mask):
        eng = self.engine
        poses = eng._f32(poses, "poses")
        if poses.dim() != 3:
            raise RuntimeError("poses must be (n_clips, n_frames, pose_dim)")
        n, f, p = poses.shape
        if len(onsets) != n or any(len(o) != 3 for o in onsets):
            raise RuntimeError("onsets must hold three onset-frame arrays (raw, backtracked, rms-backtracked) per clip")
        times = [frames_to_time(np.asarray(lst)) for clip in onsets for lst in clip]
        if any(len(t) == 0 for t in times):
            raise ZeroDivisionError("division by zero: an audio onset list is empty (GAHR divides by len(b))")
        off = np.zeros(3 * n + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(t) for t in times])
        t_dev = torch.from_numpy(np.concatenate(times) if times else np.zeros(0)).to(eng.device)
        off_dev = torch.from_numpy(off).to(eng.device)
        scores = torch.empty(n, dtype=torch.float64, device=eng.device)
        mask = torch.empty((n, 8, f), dtype=torch.uint8, device=eng.device) if want_mask else None
        with torch.cuda.device(eng.device):
            eng._check(eng.lib.egx_beat_align(
                eng._h, C.c_void_p(poses.data_ptr()), n, f, p, int(lo), int(hi), int(self.order), float(self.sigma),
                float(pose_fps), C.c_void_p(t_dev.data_ptr()), C.c_void_p(off_dev.data_ptr()),
                C.c_void_p(scores.data_ptr()), C.c_void_p(mask.data_ptr() if mask is not None else 0), eng._stream()),
                "egx_beat_align")
        return scores, mask

    def load_pose(self, pose, t_start, t_end, pose_fps, without_file=False):
        """(n_frames, pose_dim) -> the reference's 8-tuple of `(indices,)` tuples (scipy argrelextrema's return shape)."""
        pose = torch.as_tensor(np.asarray(pose) if not isinstance(pose, torch.Tensor) else pose)
        dummy = [[np.zeros(1), np.zeros(1), np.zeros(1)]]
        _, mask = self._run(pose.unsqueeze(0), dummy, t_start * pose_fps, t_end * pose_fps, pose_fps, True)
        m = mask[0].cpu().numpy()
        return tuple((np.nonzero(m[g])[0],) for g in range(8))

    @staticmethod
    def motion_frames2time(vel, offset, pose_fps):
        return vel[0] / pose_fps + offset

    @staticmethod
    def GAHR(a, b, sigma):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if len(b) == 0:
            raise ZeroDivisionError("division by zero")
        if len(a) == 0:
            return 0.0
        d = np.abs(a[None, :] - b[:, None]).min(axis=1)
        return float(np.exp(-(d ** 2) / (2 * sigma ** 2)).sum() / len(b))

    def calculate_align(self, onset_raw, onset_bt, onset_bt_rms, beat_right_arm, beat_right_shoulder, beat_right_fore_arm,
                        beat_right_wrist, beat_left_arm, beat_left_shoulder, beat_left_fore_arm, beat_left_wrist,
                        pose_fps=15):
        """The reference's signature: three onset-frame arrays and the eight `(indices,)` tuples of load_pose.  This
        host form exists for call-site compatibility (a few dozen numbers); `score_batch` is the device path."""
        avg = 0.0
        for audio_beat in (onset_raw, onset_bt, onset_bt_rms):
            for pose_beat in (beat_right_arm, beat_right_shoulder, beat_right_fore_arm, beat_right_wrist, beat_left_arm,
                              beat_left_shoulder, beat_left_fore_arm, beat_left_wrist):
                avg += self.GAHR(self.motion_frames2time(pose_beat, 0, pose_fps), frames_to_time(audio_beat), self.sigma)
        return avg / 24

    def score_batch(self, poses, onsets, t_start, t_end, pose_fps):
        """poses (n, n_frames, pose_dim) on any device; onsets: per clip the (onset_raw, onset_bt, onset_bt_rms) FRAME
        arrays of load_audio -> (n,) float64 device tensor of calculate_align scores (one launch for the batch)."""
        scores, _ = self._run(poses, onsets, t_start * pose_fps, t_end * pose_fps, pose_fps, False)
        return scores
