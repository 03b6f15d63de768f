"""ORACLE (test infrastructure, not product code): beat-alignment metric, numpy restatement.

Follows model/Beat_score_v2.py of the reference: `alignment.load_pose` (:79-127), `motion_frames2time` (:177-180),
`GAHR` (:182-196, the beat-to-audio half it returns) and `calculate_align` (:198-214).  `load_audio` (:58-77) is
three librosa onset calls (librosa is un-vendored, un-versioned and absent here): PARITY UNPINNED for that half, so
the onset frames are an INPUT of everything below.  The pose half is pinned against the real class by
oracle/make_golden_beat.py -> tests/golden/beat_align.npz.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import math

import numpy as np

# joint groups in the order load_pose RETURNS them (:127): right arm, shoulder, fore-arm, wrist, then the left ones;
# the value is the first of the group's six columns in concat(pose[:, 18:42], pose[:, 150:174]) (:101-113, :118-121)
GROUP_COL0 = (6, 0, 12, 18, 30, 24, 36, 42)


def frames_to_time(frames, sr: int = 22050, hop_length: int = 512):
    """librosa.frames_to_time with the defaults the reference calls it with (:205 passes nothing, so onset frames of
    16 kHz audio are converted at 22050 Hz; kept as the reference has it)."""
    return np.asarray(frames, dtype=np.float64) * hop_length / float(sr)


def relative_minima(x: np.ndarray, order: int) -> np.ndarray:
    """scipy.signal.argrelextrema(x, np.less, order=order)[0] (mode='clip'): strict minima against `order` neighbours
    on each side, neighbours beyond the ends clipped to the end sample (so the end samples never qualify)."""
    n = len(x)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    loc = np.arange(n)
    ok = np.ones(n, dtype=bool)
    for s in range(1, order + 1):
        ok &= x < x[np.minimum(loc + s, n - 1)]
        ok &= x < x[np.maximum(loc - s, 0)]
    return np.nonzero(ok)[0]


def velocity_norms(pose: np.ndarray) -> np.ndarray:
    """(F, P >= 174) float32 -> (8, F-1) float32: per joint group the L2 norm of the frame difference of its six
    coordinates, accumulated in float32 in column order like np.linalg.norm(np.array([...six rows...]), axis=0)."""
    pose = np.asarray(pose, dtype=np.float32)
    data = np.concatenate([pose[:, 18:42], pose[:, 150:174]], axis=1)
    vel = data[1:] - data[:-1]
    out = np.empty((8, vel.shape[0]), dtype=np.float32)
    for g, c0 in enumerate(GROUP_COL0):
        acc = vel[:, c0] * vel[:, c0]
        for k in range(1, 6):
            acc = acc + vel[:, c0 + k] * vel[:, c0 + k]
        out[g] = np.sqrt(acc)
    return out


def load_pose(pose: np.ndarray, t_start: int, t_end: int, pose_fps: int, order: int):
    """The eight beat-index arrays of alignment.load_pose, in its return order.  The right-side groups are searched in
    the window [t_start*fps, t_end*fps) (indices relative to the window), the left-side ones in the whole clip —
    as the reference has it (:114-125)."""
    v = velocity_norms(pose)
    lo, hi = t_start * pose_fps, t_end * pose_fps
    return [relative_minima(v[g][lo:hi] if g < 4 else v[g], order) for g in range(8)]


def gahr(pose_times, audio_times, sigma: float) -> float:
    """GAHR(a, b, sigma): mean over audio beats b of exp(-min_a |a - b|^2 / (2 sigma^2)); no pose beat -> 0 per term,
    no audio beat -> ZeroDivisionError, as in the reference."""
    total = 0.0
    for b in audio_times:
        l2_min = np.inf
        for a in pose_times:
            l2_min = min(l2_min, abs(a - b))
        total += math.exp(-(l2_min ** 2) / (2 * sigma ** 2))
    return total / len(audio_times)


def calculate_align(onsets_frames, pose_beats, sigma: float, pose_fps: int = 15) -> float:
    """onsets_frames: the three onset-frame arrays of load_audio (raw, backtracked, rms-backtracked); pose_beats: the
    eight arrays of load_pose -> average of the 24 GAHR scores (:198-209)."""
    avg = 0.0
    for audio_beat in onsets_frames:
        for beat in pose_beats:
            avg += gahr(np.asarray(beat) / pose_fps + 0, frames_to_time(audio_beat), sigma)
    return avg / 24
