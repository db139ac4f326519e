"""CPU tier for the pairwise-rank evaluation metrics (reference utils/evaluation.py:146-219).

* the oracle (oracle/eval_metrics.py) against golden vectors made by the reference's own functions;
* the Student-t tail of csrc/eval_math.cuh (compiled for the host) against scipy;
* csrc/eval_metrics.cu itself -- kernels and host orchestration, compiled unchanged against the CPU-thread CUDA
  emulation in tests/native/cuda_emul.h -- against the same golden vectors and the oracle.
"""
import ctypes
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import eval_metrics as oracle_eval

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "eval_*.npz")))


def _close(a, b, atol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    m = ~np.isnan(a)
    assert np.array_equal(a[m & np.isinf(b)], b[m & np.isinf(b)]), what
    m &= ~np.isinf(b)
    if m.any():
        err = np.max(np.abs(a[m] - b[m]) / np.maximum(1.0, np.abs(b[m])))
        assert err <= atol, f"{what}: err {err:.3e} > {atol:.1e}"


def check_against_golden(got, g, tol=1e-9):
    """rho / gated matrix / SAP to 1e-9; p-values to 1e-7 absolute (p is ill-conditioned as |rho| -> 1)."""
    _close(got["rho"], g["rho"], tol, "rho")
    _close(got["pvalue"], g["p"], 1e-7, "p-value")
    _close(got["corr_matrix"], g["corr_matrix"], tol, "corr_matrix")
    _close(got["sap_matrix"], g["sap_matrix"], tol, "sap_matrix")
    _close(got["Corr_score"], g["corr_score"], tol, "Corr_score")
    _close(got["SAP_score"], g["sap_score"], tol, "SAP_score")


def oracle_all(mus, ys):
    rho, p = oracle_eval.spearman(mus, ys)
    return {"rho": rho, "pvalue": p, "corr_matrix": oracle_eval.correlation_matrix(mus, ys),
            "sap_matrix": oracle_eval.sap_matrix(mus, ys), "Corr_score": oracle_eval.correlation_score(mus, ys),
            "SAP_score": oracle_eval.sap_score(mus, ys) if mus.shape[1] >= 2 else np.nan}


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_outputs(path):
    g = np.load(path)
    with np.errstate(all="ignore"):
        check_against_golden(oracle_all(g["mus"], g["ys"]), g)


def test_oracle_average_ranks_match_scipy():
    from scipy.stats import rankdata
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 300):
        x = rng.integers(0, 5, n).astype(np.float32)
        assert np.array_equal(oracle_eval.average_ranks(x), rankdata(x))
    x = np.array([0.0, -0.0, 1.0, -1.0, 0.0], dtype=np.float32)
    assert np.array_equal(oracle_eval.average_ranks(x), rankdata(x))


# ---- host builds of the device sources ---------------------------------------------------------------------
def _compile(tmp, name, src, flags=()):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = os.path.join(str(tmp), name)
    cmd = ["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-I" + os.path.join(HERE, "native"), *flags,
           "-o", out, os.path.join(HERE, "native", src)]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env)
    return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def evalmath(tmp_path_factory):
    lib = _compile(tmp_path_factory.mktemp("evalmath"), "libevalmath.so", "eval_math_host.cpp")
    for fn in (lib.evalmath_student_t_two_sided, lib.evalmath_correlation_t):
        fn.restype = ctypes.c_double
        fn.argtypes = [ctypes.c_double, ctypes.c_double]
    return lib


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    lib = _compile(tmp_path_factory.mktemp("evalemul"), "libevalemul.so", "eval_metrics_emul.cpp")
    fp, dp, i64 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.c_int64
    lib.emul_eval_metrics.argtypes = [fp, i64, i64, fp, i64, i64, i64, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp]

    def run(mus, ys):
        """mus / ys float32 arrays of any strides (element strides are passed through, as the C ABI takes them)."""
        assert mus.dtype == np.float32 and ys.dtype == np.float32
        B, Z = mus.shape
        A = ys.shape[1]
        outs = [np.full(Z * A, -7.0) for _ in range(4)] + [np.full(2, -7.0)]
        rc = lib.emul_eval_metrics(mus.ctypes.data_as(fp), mus.strides[0] // 4, mus.strides[1] // 4,
                                   ys.ctypes.data_as(fp), ys.strides[0] // 4, ys.strides[1] // 4, B, Z, A,
                                   *[o.ctypes.data_as(dp) for o in outs])
        assert rc == 0
        m = [o.reshape(Z, A) for o in outs[:4]]
        return {"rho": m[0], "pvalue": m[1], "corr_matrix": m[2], "sap_matrix": m[3],
                "Corr_score": outs[4][0], "SAP_score": outs[4][1]}
    return run


def test_student_t_tail_matches_scipy(evalmath):
    from scipy import stats
    worst = 0.0
    for dof in (1, 2, 3, 10, 62, 126, 127, 128, 129, 255, 2046, 25726, 10 ** 6):
        for t in list(np.logspace(-3, 2.5, 60)) + [1.7, 1.73, 1.75, 1.96]:
            ref = 2.0 * stats.t.sf(t, dof)
            got = evalmath.evalmath_student_t_two_sided(t, dof)
            assert got == evalmath.evalmath_student_t_two_sided(-t, dof)
            if ref > 1e-290:
                worst = max(worst, abs(got - ref) / ref)
    assert worst < 1e-9, worst
    assert evalmath.evalmath_student_t_two_sided(float("inf"), 5) == 0.0
    assert evalmath.evalmath_student_t_two_sided(0.0, 5) == 1.0
    assert np.isnan(evalmath.evalmath_student_t_two_sided(float("nan"), 5))
    assert evalmath.evalmath_correlation_t(1.0, 10) == float("inf")       # scipy: division by zero -> inf -> p = 0
    assert evalmath.evalmath_correlation_t(-1.0, 10) == float("-inf")
    assert evalmath.evalmath_correlation_t(0.0, 10) == 0.0


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_emulated_kernels_match_reference_outputs(emul, path):
    g = np.load(path)
    check_against_golden(emul(g["mus"], g["ys"]), g)


def test_emulated_kernels_edge_values_and_strides(emul):
    rng = np.random.default_rng(3)
    n = 777
    mus = rng.standard_normal((n, 5)).astype(np.float32)
    ys = rng.integers(0, 4, (n, 3)).astype(np.float32)
    mus[:, 1] = np.where(rng.random(n) < 0.5, 0.0, -0.0)          # -0.0 ties with +0.0: constant column
    mus[::7, 2] = np.inf
    mus[::11, 2] = -np.inf                                         # infinities rank; moments turn NaN like numpy's
    mus[5, 3] = np.nan                                             # NaN propagates as in spearmanr / np.cov
    ys[:, 2] = mus[:, 0] * 2.0 + 1.0                               # perfectly monotone pair: rho = 1, p = 0
    with np.errstate(all="ignore"):
        want = oracle_all(mus, ys)
    got = emul(mus, ys)
    check_against_golden(got, {"rho": want["rho"], "p": want["pvalue"], "corr_matrix": want["corr_matrix"],
                               "sap_matrix": want["sap_matrix"], "corr_score": want["Corr_score"],
                               "sap_score": want["SAP_score"]})
    assert got["rho"][0, 2] == 1.0 and got["pvalue"][0, 2] == 0.0
    assert np.isnan(got["rho"][1]).all() and np.isnan(got["rho"][3]).all()

    # column-major and padded inputs through the stride arguments
    big = np.zeros((n, 9), dtype=np.float32)
    big[:, 1:6] = mus
    ysT = np.asfortranarray(ys)
    again = emul(big[:, 1:6], ysT)
    for k in ("rho", "pvalue", "corr_matrix", "sap_matrix"):
        assert np.array_equal(got[k], again[k], equal_nan=True), k


def test_emulated_kernels_many_columns(emul):
    """More than 32 codes (two sort passes), more pairs than one CTA row covers (grid.y > 1), one-row tiles."""
    rng = np.random.default_rng(4)
    n, Z, A = 150, 130, 9
    ys = rng.integers(0, 6, (n, A)).astype(np.float32)
    mus = rng.standard_normal((n, Z)).astype(np.float32)
    mus[:, :A] += ys
    want = oracle_all(mus, ys)
    got = emul(mus, ys)
    check_against_golden(got, {"rho": want["rho"], "p": want["pvalue"], "corr_matrix": want["corr_matrix"],
                               "sap_matrix": want["sap_matrix"], "corr_score": want["Corr_score"],
                               "sap_score": want["SAP_score"]})


def test_emulated_kernels_tiny_batches(emul):
    for n in (1, 2, 3, 4):
        rng = np.random.default_rng(n)
        mus = rng.standard_normal((n, 2)).astype(np.float32)
        ys = rng.standard_normal((n, 2)).astype(np.float32)
        got = emul(mus, ys)
        if n < 3:
            assert np.isnan(got["rho"]).all() and (got["corr_matrix"] == 0).all()
        else:
            with np.errstate(all="ignore"):
                want = oracle_all(mus, ys)
            _close(got["rho"], want["rho"], 1e-12, "rho")
            _close(got["sap_matrix"], want["sap_matrix"], 1e-9, "sap")
