"""The evaluation loop of test_emotion_gesture_diversity_iterative.py:191-255 on the B200 path.

One `GestureEvaluator.step` is one iteration of that loop without the dataset:

    sampled = Emotion_VAE.sample(eid)                                   (:204)   egx_cvae3_sample
    pred_pose, _, _, _, _ = generator(in_spec, text, pre_pose, sampled) (:205)   egx_generator_forward
    logits, _ = skeleton_classifer(pred_pose)                           (:217)   egx_skeleton_forward
    acc += compute_acc(argmax(eid), logits)                             (:219-221)
    rot  = mean |target - pred| over 6-D joints                         (:223)
    _, f = FGD(pred_pose); _, g = FGD(target_pose)                      (:226-229) egx_row_features
    feature rows -> mean / covariance                                   (:230-232, 251-254) egx_fgd_accumulate
    l2  += l2_distance_pose(target, pred)                               (:236)
    BL  += alignmenter.calculate_align(onsets, load_pose(pred))         (:243-248) egx_beat_align (given the onsets)

Nothing goes through the host inside the loop: the per-frame FGD features are reduced to the
[n | sum | gram] float64 accumulator on the GPU instead of being copied into a numpy array, and
`finalize` all-reduces the accumulators (NCCL / gloo) before the Frechet tail.  The modules are the
mirrors of `emotiongestures_b200` (or a live reference generator after `install`), each already
`.cuda().eval()` with its checkpoint loaded.
"""
from __future__ import annotations

import torch

from . import fgd as _fgd


def l2_distance_pose(fake, gt):
    """test_emotion_gesture_diversity_iterative.py:46-49: mean over clips and frames of the per-frame Euclidean
    norm over the pose coordinates."""
    return (gt - fake).norm(dim=-1).mean()


def compute_acc(input_label, out):
    """test_emotion_gesture_diversity_iterative.py:35-39: top-1 accuracy in percent."""
    return 100.0 * (out.argmax(dim=1) == input_label).double().mean()


class GestureEvaluator:
    def __init__(self, generator, emotion_vae, skeleton_classifier, fgd_net, n_pre_poses: int, feature_dim: int = 512,
                 group=None, beat_sigma: float = 0.3, beat_order: int = 2, pose_fps: int = 15):
        self.generator, self.vae, self.classifier, self.fgd_net = generator, emotion_vae, skeleton_classifier, fgd_net
        self.group = group                      # process group of the clip shards (None: default group / single rank)
        self.n_pre = int(n_pre_poses)
        self.dim = int(feature_dim)
        dev = next(generator.parameters()).device
        self.device = dev
        self.acc_pred = _fgd.new_accumulator(self.dim, dev)
        self.acc_target = _fgd.new_accumulator(self.dim, dev)
        self.shift = None                       # provisional mean (first batch), keeps the f64 cancellation harmless
        self.sums = torch.zeros(6, dtype=torch.float64, device=dev)   # steps, accuracy, rotation error, l2, beat score, beat clips
        self.beat_cfg = (beat_sigma, beat_order, pose_fps)              # alignment(0.3, 2), 15 fps (:185, args)
        self._aligner = None

    def _engine(self):
        eng = getattr(self.generator, "egx_engine", None)        # a live reference module after install()
        return eng if eng is not None else self.generator.engine(getattr(self.generator, "precision", "tc"))

    @torch.no_grad()
    def step(self, in_spec, in_text_padded, pose_seq, eid_onehot, z=None, onsets=None):
        """One loop iteration; returns the predicted poses (B, F, P).  `z` (B, 32) is the sampler's Gaussian draw
        (drawn with torch.randn when omitted, like the reference; pass it for reproducible / sharded runs).
        `onsets`: per clip the (onset_raw, onset_bt, onset_bt_rms) frame arrays of alignment.load_audio; when given,
        the beat-alignment score of every predicted clip is added on the device (:243-248)."""
        dev = self.device
        pose_seq = pose_seq.to(dev, torch.float32)
        pre_pose = pose_seq[:, :self.n_pre]
        sampled = self.vae.sample(eid_onehot.to(dev, torch.float32), z=z)
        pred_pose = self.generator(in_spec.to(dev), in_text_padded.to(dev), pre_pose, sampled)[0]
        logits, _ = self.classifier(pred_pose)
        acc = compute_acc(eid_onehot.to(dev).argmax(dim=1), logits)
        b = pose_seq.shape[0]
        rot = (pose_seq.reshape(b, -1, 6) - pred_pose.reshape(b, -1, 6)).abs().mean()
        l2 = l2_distance_pose(pred_pose, pose_seq)
        eng = self._engine()
        for poses, acc_buf in ((pred_pose, self.acc_pred), (pose_seq, self.acc_target)):
            feat = self.fgd_net(poses)[1].reshape(-1, self.dim)
            if self.shift is None:
                # every rank must centre on the SAME provisional mean or the accumulators could not be summed:
                # rank 0's first-batch mean is broadcast once
                self.shift = feat.double().mean(dim=0)
                if self._world() > 1:
                    import torch.distributed as dist
                    dist.broadcast(self.shift, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                   group=self.group)
            eng.fgd_accumulate(feat, acc_buf, self.shift)
        beat = torch.zeros(2, dtype=torch.float64, device=dev)
        if onsets is not None:
            if self._aligner is None:
                from .beat import alignment
                self._aligner = alignment(self.beat_cfg[0], self.beat_cfg[1], engine=eng)
            fps = self.beat_cfg[2]
            scores = self._aligner.score_batch(pred_pose, onsets, 0, int(pred_pose.shape[1] / fps), fps)   # t_start 0 (:187-188)
            beat = torch.stack([scores.sum(), torch.tensor(float(b), dtype=torch.float64, device=dev)])
        self.sums += torch.cat([torch.stack([torch.ones((), dtype=torch.float64, device=dev), acc, rot.double(), l2.double()]), beat])
        return pred_pose

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def finalize(self):
        """All-reduce over the ranks and form the loop's summary: FGD (model/FHD_score.py:159-217 on the host),
        emotion accuracy, rotation error (degrees, x57.2958 as the reference prints it) and L2 distance."""
        if self.shift is None:
            raise RuntimeError("finalize() before any step()")
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.group)
        _fgd.all_reduce_stats(self.acc_pred, self.group)
        _fgd.all_reduce_stats(self.acc_target, self.group)
        mu_p, sig_p = _fgd.finalize_stats(self.acc_pred, self.dim, self.shift)
        mu_t, sig_t = _fgd.finalize_stats(self.acc_target, self.dim, self.shift)
        steps, acc, rot, l2, beat_sum, beat_n = (float(v) for v in self.sums.cpu())
        # the Frechet tail stays on the device (float64 eigh); the host version is the cross-check in the tests
        fgd_dev = _fgd.frechet_distance_device(*_fgd.finalize_stats_device(self.acc_pred, self.dim, self.shift),
                                               *_fgd.finalize_stats_device(self.acc_target, self.dim, self.shift))
        return {"fgd": fgd_dev, "fgd_host": _fgd.frechet_distance(mu_p, sig_p, mu_t, sig_t), "emotion_acc_percent": acc / steps,
                "rotation_error_deg": rot / steps * 57.2958, "l2_pose": l2 / steps,
                "beat_score": beat_sum / beat_n if beat_n else None,
                "pred_stats": (mu_p, sig_p), "target_stats": (mu_t, sig_t)}


def diversity_score(activations, rng=None):
    """model/FHD_score.py:244-280 (`diversity_score` + `calculate_diversity`): ten estimates of the mean L2 distance
    between five random pairs of clips' feature blocks, summarised as the centre and the bounds of the 95% normal
    interval.  `activations`: (N, 60, 512) (or anything that reshapes to it, as the reference does) tensor or array;
    `rng`: an object with numpy's `randint(low, high, size)` — the reference draws from the unseeded global
    `np.random`, so pass `np.random.RandomState(seed)` for a reproducible score.  Returns (score, (lo, hi))."""
    import numpy as np
    from scipy import stats
    rng = np.random if rng is None else rng
    act = torch.as_tensor(activations).reshape(-1, 60, 512)
    n = act.shape[0]
    window = np.empty((10, 1))
    for i in range(10):
        first = rng.randint(0, n, 5)
        second = rng.randint(0, n, 5)
        d = sum(torch.dist(act[int(a)], act[int(b)]) for a, b in zip(first, second)) / 5
        window[i] = np.float32(float(d))
    mean, std = np.mean(window, axis=0), np.std(window, axis=0)
    interval = stats.norm.interval(0.95, mean, std)
    return (interval[0] + interval[1]) / 2, interval
