"""Beat-alignment metric: oracle and CUDA path against tests/golden/beat_align.npz, which
oracle/make_golden_beat.py made with the REAL model.Beat_score_v2.alignment (load_pose + calculate_align)."""
import numpy as np
import pytest
import torch

from oracle import beat as ob
from oracle.make_golden_beat import FPS, FRAMES, ORDER, SIGMA, synth_case
from tests.helpers import load_golden


def test_oracle_matches_the_reference_made_golden():
    g = load_golden("beat_align")
    poses, onsets = synth_case(int(g["seed"]))
    for b in range(len(poses)):
        beats = ob.load_pose(poses[b], 0, FRAMES // FPS, FPS, ORDER)
        for gi in range(8):
            assert np.array_equal(np.nonzero(g["beat_mask"][b, gi])[0], beats[gi])
        assert abs(ob.calculate_align(onsets[b], beats, SIGMA, FPS) - g["scores"][b]) <= 1e-15
    win = ob.load_pose(poses[0], 1, 3, FPS, ORDER)
    assert abs(ob.calculate_align(onsets[0], win, SIGMA, FPS) - float(g["score_window_1_3"])) <= 1e-15


def test_relative_minima_is_scipy_argrelextrema():
    from scipy.signal import argrelextrema
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 7, 59):
        for order in (1, 2, 5):
            x = rng.standard_normal(n).astype(np.float32)
            x[rng.integers(0, n, 2)] = x[0]              # ties must not count (strict comparison)
            assert np.array_equal(ob.relative_minima(x, order), argrelextrema(x, np.less, order=order)[0])
    with pytest.raises(ZeroDivisionError):
        ob.gahr(np.array([0.1]), np.array([]), 0.3)
    assert ob.gahr(np.array([]), np.array([0.5, 1.0]), 0.3) == 0.0


@pytest.mark.gpu
def test_device_scores_and_beats_match_the_reference_golden():
    from emotiongestures_b200.beat import alignment
    g = load_golden("beat_align")
    poses, onsets = synth_case(int(g["seed"]))
    al = alignment(SIGMA, ORDER)
    scores, mask = al._run(torch.from_numpy(poses), onsets, 0, (FRAMES // FPS) * FPS, FPS, True)
    assert np.array_equal(mask.cpu().numpy(), g["beat_mask"]), "beat frames differ from the reference's"
    assert np.abs(scores.cpu().numpy() - g["scores"]).max() <= 1e-12
    assert np.abs(al.score_batch(torch.from_numpy(poses).cuda(), onsets, 0, FRAMES // FPS, FPS).cpu().numpy() - g["scores"]).max() <= 1e-12
    # the reference's call sequence (test_emotion_gesture_diversity_iterative.py:246-248) on one clip, windowed
    beats = al.load_pose(poses[0], 1, 3, FPS, True)
    want = ob.load_pose(poses[0], 1, 3, FPS, ORDER)
    for gi in range(8):
        assert np.array_equal(beats[gi][0], want[gi])
    assert abs(al.calculate_align(*onsets[0], *beats, FPS) - float(g["score_window_1_3"])) <= 1e-12
    with pytest.raises(ZeroDivisionError):
        al.score_batch(torch.from_numpy(poses[:1]), [[np.array([1]), np.array([], dtype=np.int64), np.array([2])]], 0, 4, FPS)
    with pytest.raises(RuntimeError, match="pose_dim is too small"):
        al.score_batch(torch.zeros(1, 34, 126), [onsets[0]], 0, 2, FPS)


@pytest.mark.gpu
def test_device_beats_on_random_clips_match_the_oracle():
    """Ragged windows / orders and 2000 random clips: beat frames bit-identical, scores to 1e-12."""
    from emotiongestures_b200.beat import alignment
    rng = np.random.default_rng(3)
    n = 2000
    poses = (rng.standard_normal((n, 34, 180)) * 0.2).astype(np.float32).cumsum(axis=1).astype(np.float32)
    onsets = [[np.sort(rng.choice(100, size=int(rng.integers(1, 9)), replace=False)) for _ in range(3)] for _ in range(n)]
    for order, (t0, t1) in ((1, (0, 2)), (3, (1, 2))):
        al = alignment(0.3, order)
        scores, mask = al._run(torch.from_numpy(poses), onsets, t0 * 15, t1 * 15, 15, True)
        scores, mask = scores.cpu().numpy(), mask.cpu().numpy()
        for b in range(0, n, 97):
            beats = ob.load_pose(poses[b], t0, t1, 15, order)
            for gi in range(8):
                assert np.array_equal(np.nonzero(mask[b, gi])[0], beats[gi]), (order, b, gi)
            assert abs(scores[b] - ob.calculate_align(onsets[b], beats, 0.3, 15)) <= 1e-12
