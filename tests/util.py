"""Shared tolerances for the parity tests.

The gate (BASELINE.json north_star): loss and dL/dz within 1e-5 relative of the reference in
fp32; the gradient is compared per regularised column against that column's max magnitude
(element-wise relative error is meaningless where row sums cancel -- SURVEY App. B); the
attribute sign matrix must be bit-exact.
"""
import numpy as np

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-5


def assert_loss_close(got, ref, rtol=LOSS_RTOL, what="loss"):
    got = float(got)
    ref = float(ref)
    if np.isnan(ref):
        assert np.isnan(got), f"{what}: expected NaN, got {got}"
        return
    assert abs(got - ref) <= rtol * abs(ref) + 1e-12, f"{what}: {got} vs {ref} (rel {abs(got-ref)/max(abs(ref),1e-300):.3e})"


def assert_grad_close(got, ref, rtol=GRAD_RTOL, what="grad"):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    if ref.size == 0:
        return
    if got.ndim == 1:
        got, ref = got[:, None], ref[:, None]
    for c in range(ref.shape[1]):
        scale = np.max(np.abs(ref[:, c]))
        err = np.max(np.abs(got[:, c] - ref[:, c]))
        assert err <= rtol * scale + 1e-12, f"{what}[:, {c}]: max err {err:.3e} vs scale {scale:.3e} (rel {err/max(scale,1e-300):.3e})"
