"""FGD sufficient statistics (mean / covariance) on the GPU with one all-reduce.

Replaces the `.cpu().numpy()` + `np.mean` / `np.cov(rowvar=False)` of
test_emotion_gesture_diversity_iterative.py:226-232,251-254 (same arithmetic in
model/FHD_score.py:240-241 and model/embedding_space_evaluator.py:132-135): every rank
accumulates [n | sum(x-s) | sum((x-s)(x-s)^T)] in float64 through egx_fgd_accumulate, the
packed buffer is all-reduced once (NCCL over NVLink on GPUs, gloo in the CPU tests), and
mu / Sigma (ddof = 1, as np.cov) are formed on every rank.  The Frechet distance tail
(scipy sqrtm, model/FHD_score.py:159-217) stays on the host.
"""
from __future__ import annotations

import numpy as np
import torch


def new_accumulator(dim: int, device) -> torch.Tensor:
    return torch.zeros(1 + dim + dim * dim, dtype=torch.float64, device=device)


def all_reduce_stats(acc: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def finalize_stats(acc, dim: int, shift=None):
    """acc [1+D+D*D] float64 -> (mu (D,), sigma (D,D)) as numpy float64, ddof=1 like np.cov."""
    a = acc.detach().cpu().numpy() if isinstance(acc, torch.Tensor) else np.asarray(acc)
    n = a[0]
    s = a[1:1 + dim]
    g = a[1 + dim:].reshape(dim, dim)
    m = s / n
    sigma = (g - n * np.outer(m, m)) / (n - 1.0)
    sigma = 0.5 * (sigma + sigma.T)
    if shift is not None:
        sh = shift.detach().cpu().numpy() if isinstance(shift, torch.Tensor) else np.asarray(shift)
        m = m + sh
    return m, sigma


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """model/FHD_score.py:159-217 (host tail), including its failure conventions: eps*I retry
    when the product is near-singular, imaginary-part check, and `return 100` on ValueError."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    try:
        covmean, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
        if not np.isfinite(covmean).all():
            offset = np.eye(sigma1.shape[0]) * eps
            covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
        if np.iscomplexobj(covmean):
            if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
                raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
            covmean = covmean.real
    except ValueError:
        return 100
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)
