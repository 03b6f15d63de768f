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


def _trace_sqrt_product(sigma1, sigma2):
    """Tr sqrt(S1 S2) through the symmetric form sqrt(S1) S2 sqrt(S1): same spectrum as S1 S2, but
    real-symmetric, so no complex square root is ever formed.  Returns (trace, min eigenvalue)."""
    w, v = np.linalg.eigh(sigma1)
    root = (v * np.sqrt(np.clip(w, 0.0, None))) @ v.T
    lam = np.linalg.eigvalsh(root @ sigma2 @ root)
    return np.sqrt(np.clip(lam, 0.0, None)).sum(), lam.min()


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """||mu1-mu2||^2 + Tr S1 + Tr S2 - 2 Tr sqrt(S1 S2)  (model/FHD_score.py:159-217; host tail).

    The reference calls scipy.linalg.sqrtm(..., disp=False), which current scipy no longer accepts;
    this keeps its observable conventions instead: a non-finite result is retried once with eps*I
    added to both covariances (:198-203), and a product whose square root would carry an imaginary
    part above 1e-3 (eigenvalue below -1e-6) returns the sentinel 100 (:206-214)."""
    mu1, mu2 = np.atleast_1d(np.asarray(mu1, np.float64)), np.atleast_1d(np.asarray(mu2, np.float64))
    sigma1 = np.atleast_2d(np.asarray(sigma1, np.float64))
    sigma2 = np.atleast_2d(np.asarray(sigma2, np.float64))
    if mu1.shape != mu2.shape or sigma1.shape != sigma2.shape:
        raise AssertionError("mean vectors / covariances have different shapes")
    tr, lam_min = _trace_sqrt_product(sigma1, sigma2)
    if not np.isfinite(tr):
        offset = np.eye(sigma1.shape[0]) * eps
        tr, lam_min = _trace_sqrt_product(sigma1 + offset, sigma2 + offset)
    if not np.isfinite(tr) or lam_min < -1e-6:
        return 100
    d = mu1 - mu2
    return float(d @ d + np.trace(sigma1) + np.trace(sigma2) - 2.0 * tr)


def finalize_stats_device(acc: torch.Tensor, dim: int, shift=None):
    """`finalize_stats` without leaving the device: (mu (D,), sigma (D,D)) as float64 tensors."""
    n = acc[0]
    m = acc[1:1 + dim] / n
    g = acc[1 + dim:].reshape(dim, dim)
    sigma = (g - n * torch.outer(m, m)) / (n - 1.0)
    sigma = 0.5 * (sigma + sigma.T)
    return (m + shift if shift is not None else m), sigma


def frechet_distance_device(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """`frechet_distance` on the GPU: the symmetric form sqrt(S1) S2 sqrt(S1) through two
    float64 `torch.linalg.eigh` calls (cuSOLVER — a library eigensolver for a once-per-evaluation D x D problem, not
    a hot-path kernel), same conventions as the host version: eps*I retry on a non-finite trace, sentinel 100 when
    the product has an eigenvalue below -1e-6.  Returns a Python float."""
    def trace_sqrt(s1, s2):
        w, v = torch.linalg.eigh(s1)
        root = (v * w.clamp_min(0).sqrt()) @ v.T
        lam = torch.linalg.eigvalsh(root @ s2 @ root)
        return lam.clamp_min(0).sqrt().sum(), lam.min()

    mu1, mu2 = mu1.double(), mu2.double()
    sigma1, sigma2 = sigma1.double(), sigma2.double()
    tr, lam_min = trace_sqrt(sigma1, sigma2)
    if not bool(torch.isfinite(tr)):
        off = torch.eye(sigma1.shape[0], dtype=torch.float64, device=sigma1.device) * eps
        tr, lam_min = trace_sqrt(sigma1 + off, sigma2 + off)
    if not bool(torch.isfinite(tr)) or float(lam_min) < -1e-6:
        return 100
    d = mu1 - mu2
    return float(d @ d + torch.trace(sigma1) + torch.trace(sigma2) - 2.0 * tr)
