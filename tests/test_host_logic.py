"""Host-side logic of the torch-facing layer that needs no GPU: dense ranks for label dtypes whose float32 cast is
not injective, python-style dim normalisation, scalar arguments given as tensors, synthetic data reproducibility."""
import numpy as np
import pytest
import torch

from arvae_b200 import ops, synth


def _sign_matrix(v):
    v = np.asarray(v, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return (v[:, None] > v[None, :]).astype(int) - (v[:, None] < v[None, :]).astype(int)


def test_rank_labels_preserve_the_sign_matrix_int64():
    a = torch.tensor([5, -3, 2**53 + 1, 2**53, 2**53 + 2, -3, 0, 2**62], dtype=torch.int64)  # float32/64 casts collide
    r = ops._rank_labels(a)
    assert r.dtype == torch.float32
    ref = (a[:, None] > a[None, :]).int() - (a[:, None] < a[None, :]).int()
    assert np.array_equal(_sign_matrix(r.numpy()), ref.numpy())
    assert r[1] == r[5]  # ties stay ties


def test_rank_labels_float64_with_nan_and_close_values():
    a = torch.tensor([1.0, 1.0 + 1e-12, float("nan"), -0.0, 0.0, float("inf"), 1.0], dtype=torch.float64)
    r = ops._rank_labels(a)
    assert torch.isnan(r[2])
    assert r[0] == r[6] and r[1] > r[0]          # 1 + 1e-12 is distinct in float64 although it rounds to 1.0f
    assert r[3] == r[4]                           # -0.0 == 0.0
    got = _sign_matrix(r.numpy())
    ref = np.sign(a.numpy()[:, None] - a.numpy()[None, :])
    ref[np.isnan(ref)] = 0
    assert np.array_equal(got, ref.astype(int))


def test_normalize_dims_python_indexing_and_errors():
    assert ops._normalize_dims((0, -1, 3, -16), 16) == (0, 15, 3, 0)
    with pytest.raises(IndexError):
        ops._normalize_dims((16,), 16)
    with pytest.raises(IndexError):
        ops._normalize_dims((-17,), 16)
    with pytest.raises(RuntimeError):
        ops._normalize_dims(tuple(range(33)), 64)


def test_scalar_accepts_numbers_and_one_element_tensors():
    assert ops._scalar(2) == 2.0 and ops._scalar(0.5) == 0.5
    assert ops._scalar(torch.tensor(1.5)) == 1.5
    assert ops._scalar(torch.FloatTensor([3.0])) == 3.0
    with pytest.raises((ValueError, RuntimeError)):
        ops._scalar(torch.tensor([1.0, 2.0]))


def test_synthetic_cases_are_reproducible_and_shaped_like_the_configs():
    a = synth.make_case("c2_dsprites_b4096", B=300)
    b = synth.make_case("c2_dsprites_b4096", B=300)
    assert torch.equal(a["z"], b["z"]) and torch.equal(a["labels"], b["labels"])
    assert a["labels"].shape == (300, 6) and set(a["labels"][:, 1].tolist()) <= {1.0, 2.0, 3.0}
    c4 = synth.make_case("c4_mnist_b65536", B=128)
    assert c4["z"].shape == (128, 16) and c4["labels"].shape == (128, 7) and c4["reg_dims"] == (1, 2, 3, 4, 5, 6)
    m = synth.make_measures(50, seed=1)
    assert m.shape == (50, 24) and m.dtype == torch.int64 and torch.equal(m, synth.make_measures(50, seed=1))


def test_labels_shape_errors_mirror_the_reference():
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            ops._prepare_labels(torch.zeros(4), (0,), 4, torch.device("cpu"))
    with pytest.raises(TypeError):
        ops._prepare_labels([1, 2, 3], (0,), 3, torch.device("cpu"))
