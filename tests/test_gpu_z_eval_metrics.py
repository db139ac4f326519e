"""GPU parity of the pairwise-rank evaluation metrics through the C ABI
(arvae_eval_metrics_f32 via arvae_b200.evaluation): golden vectors from the reference's own
utils/evaluation.py functions, the oracle on seeded inputs at sizes the goldens do not cover, reproducibility,
strides, error behaviour, and the monkey-patch into reference-shaped modules."""
import ctypes
import glob
import os
import types

import numpy as np
import pytest
import torch

from test_eval_metrics import check_against_golden, oracle_all

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "eval_*.npz")))


def _as_golden(want):
    return {"rho": want["rho"], "p": want["pvalue"], "corr_matrix": want["corr_matrix"],
            "sap_matrix": want["sap_matrix"], "corr_score": want["Corr_score"], "sap_score": want["SAP_score"]}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_matches_reference_outputs(path):
    from arvae_b200 import evaluation
    g = np.load(path)
    check_against_golden(evaluation.rank_metrics(g["mus"], g["ys"]), g)


def test_reference_named_functions_and_install():
    from arvae_b200 import evaluation
    g = np.load(os.path.join(HERE, "golden", "eval_mnist_n3000.npz"))
    mus, ys = g["mus"], g["ys"]
    assert abs(evaluation.compute_correlation_score(mus, ys)["Corr_score"] - float(g["corr_score"])) < 1e-9
    assert abs(evaluation.compute_sap_score(mus, ys)["SAP_score"] - float(g["sap_score"])) < 1e-9
    assert np.allclose(evaluation._compute_correlation_matrix(mus, ys), g["corr_matrix"], atol=1e-9, rtol=0)
    assert np.allclose(evaluation._compute_score_matrix(mus, ys), g["sap_matrix"], atol=1e-9, rtol=0)
    with pytest.raises(IndexError):
        evaluation.compute_sap_score(mus[:, :1], ys)

    ref_mod = types.ModuleType("utils_evaluation_standin")     # shaped like the reference's utils/evaluation.py
    ref_mod.compute_correlation_score = lambda a, b: {"Corr_score": -1.0}
    ref_mod.compute_sap_score = lambda a, b: {"SAP_score": -1.0}
    trainer_mod = types.ModuleType("trainer_standin")           # `from utils.evaluation import *` copies the names
    trainer_mod.compute_correlation_score = ref_mod.compute_correlation_score
    evaluation.install_evaluation(ref_mod, trainer_mod)
    try:
        assert abs(trainer_mod.compute_correlation_score(mus, ys)["Corr_score"] - float(g["corr_score"])) < 1e-9
        assert abs(ref_mod.compute_sap_score(mus, ys)["SAP_score"] - float(g["sap_score"])) < 1e-9
    finally:
        evaluation.uninstall_evaluation()
    assert ref_mod.compute_sap_score(mus, ys) == {"SAP_score": -1.0}
    assert trainer_mod.compute_correlation_score(mus, ys) == {"Corr_score": -1.0}


def test_cuda_tensor_inputs_strides_and_reproducibility():
    from arvae_b200 import evaluation
    g = np.load(os.path.join(HERE, "golden", "eval_music_n2048.npz"))
    mus = torch.from_numpy(g["mus"]).cuda()
    ys = torch.from_numpy(g["ys"]).cuda()
    a = evaluation.rank_metrics(mus, ys)
    check_against_golden(a, g)
    wide = torch.zeros(mus.shape[0], mus.shape[1] + 5, device="cuda")
    wide[:, 3:3 + mus.shape[1]] = mus
    b = evaluation.rank_metrics(wide[:, 3:3 + mus.shape[1]], ys.t().contiguous().t())   # padded rows, column-major attrs
    c = evaluation.rank_metrics(mus, ys)
    for k in ("rho", "pvalue", "corr_matrix", "sap_matrix"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
        assert np.array_equal(a[k], c[k], equal_nan=True), k                              # bitwise run to run
    assert a["Corr_score"] == c["Corr_score"] and a["SAP_score"] == c["SAP_score"]


@pytest.mark.parametrize("n,Z,A", [(1, 2, 2), (2, 3, 1), (5, 2, 2), (255, 7, 3), (2049, 33, 5), (70001, 40, 3), (300000, 4, 2)])
def test_against_oracle_at_other_sizes(n, Z, A):
    """Sizes around the sort network's classes (single chunk, cooperative, multi-launch), more than 32 codes
    (two sort passes), heavy ties."""
    from arvae_b200 import evaluation
    rng = np.random.default_rng(n + Z)
    ys = rng.integers(0, 7, (n, A)).astype(np.float32)
    mus = rng.standard_normal((n, Z)).astype(np.float32)
    k = min(Z, A)
    mus[:, :k] += 0.02 * ys[:, :k]
    mus[:, -1] = np.round(mus[:, -1], 1)
    with np.errstate(all="ignore"):
        want = oracle_all(mus, ys)
    got = evaluation.rank_metrics(mus, ys)
    check_against_golden(got, _as_golden(want))


def test_argument_errors():
    from arvae_b200 import _lib, evaluation
    x = np.zeros((10, 3), dtype=np.float32)
    with pytest.raises(ValueError):
        evaluation.rank_metrics(x, np.zeros((9, 2), dtype=np.float32))
    with pytest.raises(ValueError):
        evaluation.rank_metrics(x, np.zeros((10, 65), dtype=np.float32))
    with pytest.raises(TypeError):
        evaluation.rank_metrics(x.astype(np.float64) + 1e-12, np.zeros((10, 2), dtype=np.float32))
    ok = evaluation.rank_metrics(np.arange(30, dtype=np.float64).reshape(10, 3), np.arange(10, dtype=np.int64).reshape(10, 1))
    assert ok["rho"].shape == (3, 1) and np.all(ok["rho"] == 1.0)                         # exact casts are accepted

    lib = _lib.load()
    t = torch.zeros(10, 3, device="cuda")
    out = torch.zeros(64, dtype=torch.float64, device="cuda")
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    p = lambda v: ctypes.c_void_p(v.data_ptr())
    rc = lib.arvae_eval_metrics_f32(p(t), 3, 1, p(t), 3, 1, 10, 3, 3, p(out), p(out), p(out), p(out), p(out), p(ws), 16, None)
    assert rc != 0 and "workspace" in _lib.last_error()
    rc = lib.arvae_eval_metrics_f32(p(t), 3, 1, p(t), 3, 1, 10, 0, 3, p(out), p(out), p(out), p(out), p(out), p(ws), 16, None)
    assert rc != 0
