"""CPU oracle for the AR-VAE attribute-regularization hot path.

TEST INFRASTRUCTURE ONLY -- see the header of ``arvae_oracle.c``.  Importable
from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs; the product package ``arvae_b200`` never imports it.

Two restatements live here:

* ``arvae_oracle.c`` (through the ctypes wrappers below): plain C, float and
  double modes, row-block capable so it scales to batch sizes the reference
  cannot allocate.
* ``torch_port.py``: the reference's torch-CPU op chain written out again
  (same ATen ops in the same order), used to time "what the reference does on
  the host cores" and as a second opinion on the C code.

Both are pinned to the unmodified reference by ``tests/golden/*.npz``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libarvae_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i8p = ctypes.POINTER(ctypes.c_int8)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    """Compile ``libarvae_oracle.so`` with the Makefile next to this file."""
    src = os.path.join(_HERE, "arvae_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libarvae_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.arvae_oracle_version.restype = ctypes.c_int
        L.arvae_oracle_max_threads.restype = ctypes.c_int
        L.arvae_oracle_set_threads.argtypes = [ctypes.c_int]
        L.arvae_oracle_sign_matrix_f32.argtypes = [_f32p, _i64, _i64, _i8p]
        L.arvae_oracle_reg_rows_f32.argtypes = [_f32p, _i64, _f32p, _i64, _i64, _i64, _i64,
                                                ctypes.c_float, _f64p, _f64p, _f32p]
        L.arvae_oracle_reg_rows_f64.argtypes = [_f64p, _i64, _f64p, _i64, _i64, _i64, _i64,
                                                ctypes.c_double, _f64p, _f64p, _f64p]
        L.arvae_oracle_reparam_f32.argtypes = [_f32p, _f32p, _f32p, _i64, _f32p]
        L.arvae_oracle_kld_f32.argtypes = [_f32p, _f32p, _i64, _i64, ctypes.c_float,
                                           ctypes.c_float, _f32p, _f32p, _f32p, _f32p]
        L.arvae_oracle_kld_f64.argtypes = [_f64p, _f64p, _i64, _i64, ctypes.c_double,
                                           ctypes.c_double, _f64p, _f64p, _f64p, _f64p]
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().arvae_oracle_max_threads())


def set_threads(n: int) -> None:
    lib().arvae_oracle_set_threads(int(n))


def _p(arr: Optional[np.ndarray], typ):
    return None if arr is None else arr.ctypes.data_as(typ)


def sign_matrix(a: np.ndarray) -> np.ndarray:
    """int8 [B,B] matrix sign(a_i - a_j) (reference utils/trainer.py:394-395,400)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    B = a.shape[0]
    out = np.empty((B, B), dtype=np.int8)
    lib().arvae_oracle_sign_matrix_f32(_p(a, _f32p), 1, B, _p(out, _i8p))
    return out


def reg_rows(x: np.ndarray, a: np.ndarray, factor: float, row_begin: int = 0,
             row_end: Optional[int] = None, want_grad: bool = True, f64: bool = False
             ) -> Tuple[float, np.ndarray, Optional[np.ndarray]]:
    """Row block of ``reg_loss_sign`` (utils/trainer.py:378-403).

    Returns ``(loss_sum, row_loss[rows], grad[rows])``: the UNnormalised sum of
    ``|tanh(f(x_i-x_j)) - sign(a_i-a_j)|`` over the block's rows and all B
    columns, the same per row, and d(mean loss)/dx_i (already divided by B^2).
    """
    dt = np.float64 if f64 else np.float32
    x = np.ascontiguousarray(x, dtype=dt)
    a = np.ascontiguousarray(a, dtype=dt)
    B = x.shape[0]
    assert a.shape[0] == B
    if row_end is None:
        row_end = B
    n = row_end - row_begin
    loss = ctypes.c_double(0.0)
    row_loss = np.empty(n, dtype=np.float64)
    grad = np.empty(n, dtype=dt) if want_grad else None
    if f64:
        lib().arvae_oracle_reg_rows_f64(_p(x, _f64p), 1, _p(a, _f64p), 1, B, row_begin, row_end,
                                        float(factor), ctypes.byref(loss), _p(row_loss, _f64p),
                                        _p(grad, _f64p))
    else:
        lib().arvae_oracle_reg_rows_f32(_p(x, _f32p), 1, _p(a, _f32p), 1, B, row_begin, row_end,
                                        float(factor), ctypes.byref(loss), _p(row_loss, _f64p),
                                        _p(grad, _f32p))
    return loss.value, row_loss, grad


def compute_reg_loss(z: np.ndarray, labels: np.ndarray, reg_dim: int, gamma: float,
                     factor: float = 1.0, f64: bool = False) -> Tuple[float, np.ndarray]:
    """``Trainer.compute_reg_loss`` (utils/trainer.py:369-376) + its gradient.

    ``labels`` is the 1-D attribute vector the reference's callers pass
    (``labels[:, dim]``).  Returns ``(gamma * mean loss, d/dz [B,Z])``.
    """
    z = np.asarray(z)
    B, Z = z.shape
    col = z[:, reg_dim]
    loss_sum, _, g = reg_rows(col, labels, factor, f64=f64)
    dt = np.float64 if f64 else np.float32
    grad = np.zeros((B, Z), dtype=dt)
    if B:
        grad[:, reg_dim] = (dt(gamma) * g).astype(dt)
        loss = gamma * (loss_sum / (float(B) * float(B)))
    else:
        loss = float("nan")
    return (float(loss) if f64 else float(np.float32(loss))), grad


def compute_reg_loss_multi(z: np.ndarray, labels: np.ndarray, reg_dims, gamma: float,
                           factor: float = 1.0, f64: bool = False,
                           row_begin: int = 0, row_end: Optional[int] = None
                           ) -> Tuple[float, np.ndarray]:
    """The callers' per-dim loop (imagevae/image_vae_trainer.py:171-180,
    measurevae/measure_vae_trainer.py:131-142): sum over ``dim`` of
    ``compute_reg_loss(z, labels[:, dim], dim, gamma, factor)``.

    With a row range, returns that row block's share of the loss (so that the
    shares of disjoint blocks add up to the full loss) and the gradient rows.
    """
    z = np.asarray(z)
    labels = np.asarray(labels)
    B, Z = z.shape
    if row_end is None:
        row_end = B
    dt = np.float64 if f64 else np.float32
    grad = np.zeros((row_end - row_begin, Z), dtype=dt)
    total = 0.0
    for dim in reg_dims:
        loss_sum, _, g = reg_rows(z[:, dim], labels[:, dim], factor, row_begin, row_end, f64=f64)
        grad[:, dim] += (dt(gamma) * g).astype(dt)
        total += gamma * (loss_sum / (float(B) * float(B)))
    return total, grad


def reparam(loc: np.ndarray, scale: np.ndarray, eps: np.ndarray) -> np.ndarray:
    """``Normal(loc, scale).rsample()`` with given noise (imagevae/mnist_vae.py:79)."""
    loc = np.ascontiguousarray(loc, dtype=np.float32)
    scale = np.ascontiguousarray(scale, dtype=np.float32)
    eps = np.ascontiguousarray(eps, dtype=np.float32)
    z = np.empty_like(loc)
    lib().arvae_oracle_reparam_f32(_p(loc, _f32p), _p(scale, _f32p), _p(eps, _f32p), loc.size,
                                   _p(z, _f32p))
    return z


def kld(loc: np.ndarray, scale: np.ndarray, beta: float, c: float = 0.0, f64: bool = False):
    """``Trainer.compute_kld_loss`` vs the unit prior (utils/trainer.py:354-367).

    Returns ``(kld_mean, loss, dloss/dloc, dloss/dscale)``.
    """
    dt = np.float64 if f64 else np.float32
    loc = np.ascontiguousarray(loc, dtype=dt)
    scale = np.ascontiguousarray(scale, dtype=dt)
    B, Z = loc.shape
    dloc = np.empty_like(loc)
    dscale = np.empty_like(scale)
    if f64:
        k = ctypes.c_double()
        l = ctypes.c_double()
        lib().arvae_oracle_kld_f64(_p(loc, _f64p), _p(scale, _f64p), B, Z, beta, c,
                                   ctypes.byref(k), ctypes.byref(l), _p(dloc, _f64p),
                                   _p(dscale, _f64p))
    else:
        k = ctypes.c_float()
        l = ctypes.c_float()
        lib().arvae_oracle_kld_f32(_p(loc, _f32p), _p(scale, _f32p), B, Z, beta, c,
                                   ctypes.byref(k), ctypes.byref(l), _p(dloc, _f32p),
                                   _p(dscale, _f32p))
    return k.value, l.value, dloc, dscale
