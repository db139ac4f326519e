"""Pairwise-rank evaluation metrics on the device.

Mirrors the two rank / second-moment metrics of the reference's ``utils/evaluation.py``:

* ``compute_correlation_score`` / ``_compute_correlation_matrix`` (:146-173) -- "SCC": |Spearman rho| of every
  (latent code, attribute) pair where ``scipy.stats.spearmanr``'s p-value is <= 0.05, max over codes, mean over
  attributes.  The reference makes Z x A scipy calls in a Python double loop (16 x 6 calls on 25 728 samples).
* ``compute_sap_score`` / ``_compute_score_matrix`` (:176-219) -- "SAP": cov^2 / (var var), top-1 minus top-2.

Same names, arguments (``latent_codes [N, Z]``, ``attributes [N, A]``; numpy arrays as the trainers pass them, or
torch tensors) and return values (``{"Corr_score": float}``, ``{"SAP_score": float}``, ``[Z, A]`` float64 arrays).
One C-ABI call (``arvae_eval_metrics_f32``) produces everything; :func:`rank_metrics` returns it all at once.
``install_evaluation`` swaps the functions into the reference's modules.  The mutual-information metrics of that file
(MIG, modularity, interpretability: sklearn's k-NN estimator) are not part of this path.

Inputs must be exactly representable in float32 (the trainers' ``compute_representations`` yields float32 arrays);
anything else raises instead of being rounded silently.  CUDA is required: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import numpy as np
import torch

from . import _lib

__all__ = ["rank_metrics", "compute_correlation_score", "compute_sap_score", "_compute_correlation_matrix",
           "_compute_score_matrix", "install_evaluation", "uninstall_evaluation"]


def _to_device_f32(x, name: str, device) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"arvae_b200: {name} must be a numpy array or a torch tensor, got {type(x).__name__}")
    if t.dim() != 2:
        raise ValueError(f"arvae_b200: {name} must be [num_points, num_columns], got shape {tuple(t.shape)}")
    t = t.detach().to(device)
    if t.dtype != torch.float32:
        f = t.to(torch.float32)
        same = (f.to(torch.float64) == t.to(torch.float64)) | (torch.isnan(f) & torch.isnan(t.to(torch.float64)))
        if not bool(same.all()):
            raise TypeError(f"arvae_b200: {name} ({t.dtype}) holds values that float32 cannot represent; "
                            "the evaluation path takes float32 columns")
        t = f
    return t


def rank_metrics(latent_codes, attributes, device=None) -> Dict[str, object]:
    """All pairwise metrics in one launch sequence.

    Returns ``rho``, ``pvalue``, ``corr_matrix``, ``sap_matrix`` ([Z, A] float64 numpy arrays) and the scalars
    ``Corr_score`` / ``SAP_score`` (``SAP_score`` is NaN when Z < 2 or a score column holds NaN, as np.sort orders it).
    """
    if not torch.cuda.is_available():
        raise RuntimeError("arvae_b200: the evaluation metrics need a CUDA device (no CPU fallback exists for this path)")
    if device is None:
        device = latent_codes.device if isinstance(latent_codes, torch.Tensor) and latent_codes.is_cuda else "cuda"
    device = torch.device(device)
    mus = _to_device_f32(latent_codes, "latent_codes", device)
    ys = _to_device_f32(attributes, "attributes", device)
    B, Z = mus.shape
    A = ys.shape[1]
    if ys.shape[0] != B:
        raise ValueError(f"arvae_b200: {B} latent codes but {ys.shape[0]} attribute rows")
    lib = _lib.load()
    with torch.cuda.device(device):
        nbytes = lib.arvae_eval_metrics_workspace_bytes(B, Z, A)
        if nbytes == 0:
            raise ValueError(f"arvae_b200: evaluation metrics need N >= 1, 1 <= Z <= 1024, 1 <= A <= 64 (got N={B}, Z={Z}, A={A})")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        out = torch.empty(4 * Z * A + 2, dtype=torch.float64, device=device)
        ptr = [ctypes.c_void_p(out.data_ptr() + 8 * k * Z * A) for k in range(5)]
        rc = lib.arvae_eval_metrics_f32(
            ctypes.c_void_p(mus.data_ptr()), mus.stride(0), mus.stride(1),
            ctypes.c_void_p(ys.data_ptr()), ys.stride(0), ys.stride(1), B, Z, A,
            ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ctypes.c_void_p(ws.data_ptr()), nbytes,
            ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _lib.check(rc, "arvae_eval_metrics_f32")
        host = out.cpu().numpy()          # the one synchronisation of the call
    mats = host[:4 * Z * A].reshape(4, Z, A)
    return {"rho": mats[0].copy(), "pvalue": mats[1].copy(), "corr_matrix": mats[2].copy(),
            "sap_matrix": mats[3].copy(), "Corr_score": float(host[-2]), "SAP_score": float(host[-1])}


def _compute_correlation_matrix(mus, ys) -> np.ndarray:
    """utils/evaluation.py:157-173."""
    return rank_metrics(mus, ys)["corr_matrix"]


def compute_correlation_score(latent_codes, attributes) -> Dict[str, float]:
    """utils/evaluation.py:146-155."""
    return {"Corr_score": rank_metrics(latent_codes, attributes)["Corr_score"]}


def _compute_score_matrix(mus, ys) -> np.ndarray:
    """utils/evaluation.py:194-214."""
    return rank_metrics(mus, ys)["sap_matrix"]


def compute_sap_score(latent_codes, attributes) -> Dict[str, float]:
    """utils/evaluation.py:176-191 (two codes at least, as the reference's ``sorted_matrix[-2]`` needs)."""
    if latent_codes.shape[1] < 2:
        raise IndexError("index -2 is out of bounds for axis 0 with size 1")
    return {"SAP_score": rank_metrics(latent_codes, attributes)["SAP_score"]}


_PATCHED = ("compute_correlation_score", "compute_sap_score", "_compute_correlation_matrix", "_compute_score_matrix")
_saved = []


def install_evaluation(*modules) -> None:
    """Swap the four functions into the given modules: the reference's ``utils.evaluation`` and, because the trainers
    bind the names at import time (``from utils.evaluation import *``, imagevae/image_vae_trainer.py:15,
    measurevae/measure_vae_trainer.py:13), the trainer modules themselves."""
    g = globals()
    for mod in modules:
        for name in _PATCHED:
            if hasattr(mod, name):
                _saved.append((mod, name, getattr(mod, name)))
                setattr(mod, name, g[name])


def uninstall_evaluation() -> None:
    while _saved:
        mod, name, fn = _saved.pop()
        setattr(mod, name, fn)
