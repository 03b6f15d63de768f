"""CPU, world_size 2 over gloo: the two collectives of the path (pose gather, FGD statistics
all-reduce) and the host-side FGD arithmetic against numpy / the reference formula."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emotiongestures_b200 import fgd
from emotiongestures_b200.sharding import all_gather_poses, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(x, shift):
    """What egx_fgd_accumulate produces for rows x (float64 here; the kernel is checked on the GPU)."""
    y = x.astype(np.float64) - shift
    return torch.from_numpy(np.concatenate([[float(len(x))], y.sum(0), (y.T @ y).ravel()]))


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                      # same global data on every rank
        poses = torch.from_numpy(rng.standard_normal((n_clips, 34, 126)).astype(np.float32))
        feats = (rng.standard_normal((n_clips, 32)) * 3 + 5).astype(np.float32)
        lo, hi = shard_bounds(n_clips, rank, world)
        gathered = all_gather_poses(poses[lo:hi].clone(), n_clips)
        ok_gather = torch.equal(gathered, poses)
        shift = feats[:4].astype(np.float64).mean(0)        # provisional mean, identical on all ranks
        acc = _pack(feats[lo:hi], shift)
        fgd.all_reduce_stats(acc)
        mu, sigma = fgd.finalize_stats(acc, 32, shift)
        q.put((rank, ok_gather, mu, sigma))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [64, 37])
def test_world2_gather_and_fgd_allreduce(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    rng.standard_normal((n_clips, 34, 126))
    feats = (rng.standard_normal((n_clips, 32)) * 3 + 5).astype(np.float32).astype(np.float64)
    for rank, ok_gather, mu, sigma in res:
        assert ok_gather, f"rank {rank}: gathered poses differ from the global batch"
        np.testing.assert_allclose(mu, feats.mean(0), rtol=1e-12)
        np.testing.assert_allclose(sigma, np.cov(feats, rowvar=False), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(res[0][2], res[1][2])     # every rank ends with the same statistics


def test_frechet_distance_matches_closed_form_and_failure_convention():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((500, 16)); b = rng.standard_normal((500, 16)) * 1.5 + 0.3
    m1, s1, m2, s2 = a.mean(0), np.cov(a, rowvar=False), b.mean(0), np.cov(b, rowvar=False)
    d = fgd.frechet_distance(m1, s1, m2, s2)
    # symmetric-eigendecomposition form of Tr sqrt(S1 S2)
    w, v = np.linalg.eigh(s1)
    r = (v * np.sqrt(w)) @ v.T
    tr = np.sqrt(np.clip(np.linalg.eigvalsh(r @ s2 @ r), 0, None)).sum()
    ref = ((m1 - m2) ** 2).sum() + np.trace(s1) + np.trace(s2) - 2 * tr
    assert abs(d - ref) <= 1e-8 * abs(ref)
    assert fgd.frechet_distance(m1, s1, m1, s1) == pytest.approx(0.0, abs=1e-6)
    from scipy import linalg
    assert abs(d - (((m1 - m2) ** 2).sum() + np.trace(s1) + np.trace(s2)
                    - 2 * np.trace(linalg.sqrtm(s1 @ s2)).real)) <= 1e-6 * abs(ref)
    bad = -np.eye(16)                                   # not a covariance: the reference's sentinel
    assert fgd.frechet_distance(m1, s1, m2, bad) == 100
