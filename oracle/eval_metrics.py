"""CPU restatement of the reference's pairwise-rank evaluation metrics (TEST INFRASTRUCTURE ONLY).

Follows utils/evaluation.py of the reference:
  :146-155 compute_correlation_score      :157-173 _compute_correlation_matrix (scipy.stats.spearmanr, p <= 0.05 gate)
  :176-191 compute_sap_score              :194-214 _compute_score_matrix (np.cov, ddof=1, var_mu > 1e-12 gate)
  :217-219 _compute_avg_diff_top_two
The reference calls scipy.stats.spearmanr (scipy is a dependency of the reference, not vendored in it); its
published algorithm is restated here: average ranks, Pearson correlation of the ranks, and the two-sided p-value
of t = rho * sqrt(dof / ((1 + rho)(1 - rho))) under Student's t with dof = n - 2.  Vectorised numpy in float64,
the t-distribution tail through scipy.special.stdtr.  Pinned to tests/golden/eval_*.npz, which
tests/golden/make_golden_eval.py produced by running the reference's own functions.
"""
from __future__ import annotations

import numpy as np
from scipy import special


def average_ranks(col: np.ndarray) -> np.ndarray:
    """1-based ranks, ties share the mean of their positions (scipy.stats.rankdata 'average')."""
    col = np.asarray(col)
    order = np.argsort(col, kind="stable")
    s = col[order]
    n = len(s)
    head = np.ones(n, dtype=bool)
    head[1:] = s[1:] != s[:-1]
    start = np.flatnonzero(head)
    end = np.append(start[1:], n)
    group = np.cumsum(head) - 1
    avg = 0.5 * (start + end + 1)          # positions start..end-1 -> ranks start+1..end
    ranks = np.empty(n, dtype=np.float64)
    ranks[order] = avg[group]
    return ranks


def spearman(mus: np.ndarray, ys: np.ndarray):
    """(rho [Z,A], p [Z,A]) as scipy.stats.spearmanr(mus[:, i], ys[:, j]) returns them; NaN for constant
    or NaN-holding columns and for fewer than 3 samples."""
    mus = np.asarray(mus)
    ys = np.asarray(ys)
    n, Z = mus.shape
    A = ys.shape[1]
    rho = np.full((Z, A), np.nan)
    p = np.full((Z, A), np.nan)
    if n < 3:
        return rho, p
    bad_m = np.isnan(mus.astype(np.float64)).any(0)
    bad_y = np.isnan(ys.astype(np.float64)).any(0)
    rm = np.stack([average_ranks(mus[:, i]) for i in range(Z)], 1) - 0.5 * (n + 1)
    ry = np.stack([average_ranks(ys[:, j]) for j in range(A)], 1) - 0.5 * (n + 1)
    sxx = (rm * rm).sum(0)
    syy = (ry * ry).sum(0)
    sxy = rm.T @ ry
    dof = n - 2
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.clip(sxy / np.sqrt(sxx[:, None] * syy[None, :]), -1.0, 1.0)
        t = r * np.sqrt(np.clip(dof / ((r + 1.0) * (1.0 - r)), 0, None))
        pv = 2.0 * special.stdtr(dof, -np.abs(t))
    ok = ~(bad_m[:, None] | bad_y[None, :]) & (sxx[:, None] > 0) & (syy[None, :] > 0)
    rho[ok] = r[ok]
    p[ok] = pv[ok]
    return rho, p


def correlation_matrix(mus, ys) -> np.ndarray:
    """_compute_correlation_matrix (:157-173): |rho| where p <= 0.05, else 0 (NaN p counts as 'else')."""
    rho, p = spearman(mus, ys)
    with np.errstate(invalid="ignore"):
        return np.where(p <= 0.05, np.abs(rho), 0.0)


def correlation_score(mus, ys) -> float:
    """compute_correlation_score (:146-155)."""
    return float(np.mean(np.max(correlation_matrix(mus, ys), axis=0)))


def sap_matrix(mus, ys) -> np.ndarray:
    """_compute_score_matrix (:194-214): cov^2 / (var_mu var_y) with ddof = 1; 0 where var_mu <= 1e-12."""
    mus = np.asarray(mus, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    n = mus.shape[0]
    dm = mus - mus.mean(0)
    dy = ys - ys.mean(0)
    with np.errstate(divide="ignore", invalid="ignore"):
        cov = dm.T @ dy / (n - 1)
        var_m = (dm * dm).sum(0) / (n - 1)
        var_y = (dy * dy).sum(0) / (n - 1)
        score = cov * cov / (var_m[:, None] * var_y[None, :])
    score[~(var_m > 1e-12), :] = 0.0
    return score


def avg_diff_top_two(matrix: np.ndarray) -> float:
    """_compute_avg_diff_top_two (:217-219); np.sort puts NaN last, so one NaN in a column makes it NaN."""
    s = np.sort(matrix, axis=0)
    return float(np.mean(s[-1, :] - s[-2, :]))


def sap_score(mus, ys) -> float:
    """compute_sap_score (:176-191)."""
    return avg_diff_top_two(sap_matrix(mus, ys))
