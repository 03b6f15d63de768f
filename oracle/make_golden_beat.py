"""ORACLE tooling: pin oracle/beat.py against the REAL model.Beat_score_v2.alignment and write
tests/golden/beat_align.npz.  Build container only (needs /root/reference):
    python -m oracle.make_golden_beat

The reference file imports librosa and matplotlib at module level; both are absent here and neither is touched by
load_pose / GAHR.  calculate_align calls librosa.frames_to_time(frames) with its defaults, i.e. the documented
`frames * hop_length / sr` with sr = 22050, hop_length = 512; the stub supplies exactly that.  load_audio (the librosa
onset detector) is NOT exercised: onset frames are synthetic inputs, and that half stays unpinned.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import beat as ob  # noqa: E402

REF = "/root/reference"
N_CLIPS, FRAMES, POSE_DIM, FPS, ORDER, SIGMA = 6, 60, 282, 15, 2, 0.3


def synth_case(seed=0):
    """Smooth-ish random poses (sums of a few sinusoids + noise, so local minima of the speed exist) and three sorted
    onset-frame lists per clip; clip 4 has a flat (no-beat) right wrist, clip 5 a single audio onset per list."""
    rng = np.random.default_rng(seed)
    t = np.arange(FRAMES)[None, :, None]
    poses = np.zeros((N_CLIPS, FRAMES, POSE_DIM))
    for _ in range(4):
        poses += rng.uniform(0.05, 0.3, (N_CLIPS, 1, POSE_DIM)) * np.sin(
            rng.uniform(0.2, 1.5, (N_CLIPS, 1, POSE_DIM)) * t + rng.uniform(0, 6.28, (N_CLIPS, 1, POSE_DIM)))
    poses += 0.01 * rng.standard_normal(poses.shape)
    poses = poses.astype(np.float32)
    poses[4, :, 36:42] = 0.25                                   # right wrist never moves: no beats in that group
    onsets = []
    for b in range(N_CLIPS):
        lists = []
        for _ in range(3):
            n = 1 if b == 5 else int(rng.integers(2, 12))
            lists.append(np.sort(rng.choice(170, size=n, replace=False)).astype(np.int64))
        onsets.append(lists)
    return poses, onsets


def main():
    sys.path.insert(0, REF)
    lib = types.ModuleType("librosa")
    lib.frames_to_time = lambda frames, sr=22050, hop_length=512: np.asarray(frames) * hop_length / float(sr)
    lib.display = types.ModuleType("librosa.display")
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.figure = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules.update({"librosa": lib, "librosa.display": lib.display, "matplotlib": mpl, "matplotlib.pyplot": plt})
    from model.Beat_score_v2 import alignment

    poses, onsets = synth_case(0)
    al = alignment(SIGMA, ORDER)
    t_start, t_end = 0, FRAMES // FPS                   # test_emotion_gesture_diversity_iterative.py:187-188
    scores, masks = [], np.zeros((N_CLIPS, 8, FRAMES), np.uint8)
    for b in range(N_CLIPS):
        beats = al.load_pose(poses[b], t_start, t_end, FPS, True)
        mine = ob.load_pose(poses[b], t_start, t_end, FPS, ORDER)
        for g in range(8):
            assert np.array_equal(beats[g][0], mine[g]), (b, g)
            masks[b, g, beats[g][0]] = 1
        s_ref = al.calculate_align(*onsets[b], *beats, FPS)
        s_mine = ob.calculate_align(onsets[b], mine, SIGMA, FPS)
        assert abs(s_ref - s_mine) <= 1e-15, (b, s_ref, s_mine)
        scores.append(s_ref)
        print(f"clip {b}: beats per group {[len(x[0]) for x in beats]}, score {s_ref:.6f}")
    # a second window: the right-side groups only see frames [15, 45)
    beats = al.load_pose(poses[0], 1, 3, FPS, True)
    mine = ob.load_pose(poses[0], 1, 3, FPS, ORDER)
    for g in range(8):
        assert np.array_equal(beats[g][0], mine[g])
    s_win = al.calculate_align(*onsets[0], *beats, FPS)
    path = os.path.join(ROOT, "tests", "golden", "beat_align.npz")
    np.savez_compressed(path, seed=np.int64(0), scores=np.array(scores), beat_mask=masks, score_window_1_3=np.float64(s_win),
                        order=np.int64(ORDER), sigma=np.float64(SIGMA), fps=np.int64(FPS))
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
