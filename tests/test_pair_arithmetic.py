"""The algebra of the pair kernel's constant-sign loops (csrc/reg_sorted.cu: loop_const), restated in numpy float32 and
held against a float64 evaluation of the reference formula (utils/trainer.py:390-401): CPU tier, no GPU.

A constant-sign tile needs, per row i, only S1 = sum_j r_ij and S2 = sum_j r_ij^2 with r = (1 - tanh(f (x_i - x_j))) / 2
(DESIGN section 2).  The kernel forms them as sums of q = 1 - r = 1 / (1 + E_j F_i) in three ways; this file checks that
each of them IS those sums (to float32 accumulation accuracy), that the shared-reciprocal forms stay finite exactly up to
the range the device flag guards (|u| <= 31), and that the packed Newton reciprocal is accurate to an ulp.
The GPU tier (tests/test_gpu_parity.py) checks the kernels themselves against the oracle and the golden vectors."""
import numpy as np
import pytest

f32 = np.float32
LOG2E = 1.4426950408889634


def fma(a, b, c):
    """IEEE float32 fused multiply-add (exact product and sum in float64 -- 24 x 24 bits fit -- then one rounding)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def rcp_newton(x, ops=5):
    """csrc/reg_sorted.cu: rcp_newton2 -- magic-constant seed, then quadratic + cubic step (5 FFMA) or three quadratic (6)."""
    x = np.asarray(x, f32)
    y = (np.int32(0x7EF311C7) - x.view(np.int32)).view(f32)
    one = np.ones_like(x)
    if ops == 5:
        e = fma(-x, y, one)
        y = fma(y, e, y)
        e = fma(-x, y, one)
        t = fma(e, e, e)
        return fma(y, t, y)
    for _ in range(3):
        e = fma(-x, y, one)
        y = fma(y, e, y)
    return y


def rcp_f32(x):
    return (1.0 / np.asarray(x, np.float64)).astype(f32)  # stands for MUFU.RCP (1 ulp)


def truth(xi, xj, factor):
    """float64: S1 = sum q, S2 = sum q^2 with q = (1 + tanh(f (x_i - x_j))) / 2 = 1 - r."""
    t = np.tanh(factor * (np.float64(xi) - xj.astype(np.float64)))
    q = 0.5 * (1.0 + t)
    return q.sum(), (q * q).sum()


def kernel_operands(xi, xj, factor):
    cabs = f32(abs(2.0 * factor * LOG2E))
    ui, uj = f32(cabs * f32(xi)), (cabs * xj.astype(f32)).astype(f32)
    E = np.exp2(uj.astype(np.float64)).astype(f32)          # Es, built once per element
    F = f32(np.exp2(-np.float64(ui)))                       # per row, as the row loader builds it
    return E, F


def sums_per_pair(E, F, newton_every=0):
    s = fma(E, F, np.ones_like(E))
    q = rcp_f32(s)
    if newton_every:
        q[::newton_every] = rcp_newton(s[::newton_every])
    return q.astype(np.float64).sum(), (q.astype(np.float64) ** 2).sum()


def sums_shared(E, F):
    """one reciprocal per two pairs, both quotients (ARVAE_SHARE_FORM 1)"""
    a, b = fma(E[0::2], F, f32(1)), fma(E[1::2], F, f32(1))
    rp = rcp_f32((a * b).astype(f32))
    qa, qb = (rp * b).astype(f32), (rp * a).astype(f32)
    q = np.concatenate([qa, qb]).astype(np.float64)
    return q.sum(), (q * q).sum()


def sums_staged(E, F, newton_every=0):
    """sums only, from staged column-pair sums T and products P (ARVAE_SHARE_FORM 3, the shipped loop)"""
    T, P = (E[0::2] + E[1::2]).astype(f32), (E[0::2] * E[1::2]).astype(f32)
    F2 = f32(F * F)
    tm1 = fma(T, F, f32(1))
    p = fma(P, F2, tm1)
    rp = rcp_f32(p)
    if newton_every:
        rp[::newton_every] = rcp_newton(p[::newton_every])
    w = fma(rp, tm1, rp)
    S1 = w.astype(np.float64).sum()
    S2 = (w.astype(np.float64) ** 2).sum() - 2.0 * rp.astype(np.float64).sum()
    return S1, S2


@pytest.mark.parametrize("factor", [1.0, 10.0, -1.0, 0.37])
@pytest.mark.parametrize("form", ["per_pair", "per_pair_newton", "shared", "staged", "staged_newton"])
def test_constant_sign_sums_equal_the_reference_formula(form, factor):
    rng = np.random.default_rng(5)
    u_cap = 31.0 / abs(2.0 * factor * LOG2E)                 # the shared forms' guarded range in units of x
    xj = np.clip(rng.normal(0.0, 1.2, 4096), -u_cap, u_cap).astype(f32)
    for xi in (f32(0.0), f32(0.9 * u_cap), f32(-0.9 * u_cap), f32(0.013)):
        xs_i, xs_j = (xi, xj) if factor > 0 else (-xi, -xj)  # the kernels work on sgn(f) x
        E, F = kernel_operands(xs_i, xs_j, abs(factor))
        got = {"per_pair": lambda: sums_per_pair(E, F), "per_pair_newton": lambda: sums_per_pair(E, F, 3),
               "shared": lambda: sums_shared(E, F), "staged": lambda: sums_staged(E, F),
               "staged_newton": lambda: sums_staged(E, F, 8)}[form]()
        want = truth(xi, xj, factor)
        assert abs(got[0] - want[0]) <= 2e-6 * max(want[0], 1.0), (form, xi)
        assert abs(got[1] - want[1]) <= 2e-6 * max(want[0], 1.0), (form, xi)
        # the gradient sum S1 - S2 = sum q (1 - q) is what carries the cancellation: absolute accuracy of the sums
        assert abs((got[0] - got[1]) - (want[0] - want[1])) <= 4e-6 * max(want[0], 1.0), (form, xi)


def test_shared_reciprocal_range_is_exactly_what_the_flag_guards():
    """a b <= (1 + 2^62)^2 is finite in float32 for |u| <= 31 (kSharedMaxAbsU) at both extremes; one step further out it
    overflows and the quotients turn into 0 / NaN -- which is why a device flag sends such calls to the per-pair form."""
    E = np.exp2(np.array([31.0, 31.0, -31.0, -31.0])).astype(f32)
    for ui in (31.0, -31.0):
        S1, S2 = sums_staged(E, f32(np.exp2(-ui)))
        S1p, S2p = sums_per_pair(E, f32(np.exp2(-ui)))
        assert np.isfinite([S1, S2]).all()
        assert abs(S1 - S1p) < 1e-6 and abs(S2 - S2p) < 1e-6
    E = np.exp2(np.array([40.0, 40.0])).astype(f32)
    with np.errstate(over="ignore", invalid="ignore"):
        S1, _ = sums_staged(E, f32(np.exp2(40.0)))            # u_i = -40: 1 + 2^80 each, the product overflows
        S1p, _ = sums_per_pair(E, f32(np.exp2(40.0)))
    assert not np.isfinite(S1) or abs(S1 - S1p) > 1e-30       # the shared form is wrong here ...
    assert np.isfinite(S1p) and S1p < 1e-20                   # ... the per-pair form (valid to |u| <= 62) is not


@pytest.mark.parametrize("ops", [5, 6])
def test_packed_newton_reciprocal_is_accurate_to_an_ulp(ops):
    rng = np.random.default_rng(0)
    x = np.concatenate([np.exp2(rng.uniform(0.0, 124.0, 400_000)), 1.0 + np.exp2(rng.uniform(-30.0, 1.0, 200_000)),
                        [1.0, 2.0, 3.0, 2.0 ** 124]]).astype(f32)
    y = rcp_newton(x, ops).astype(np.float64)
    rel = np.abs(y * x.astype(np.float64) - 1.0)
    assert rel.max() <= (1.3 if ops == 5 else 1.01) * 2.0 ** -24
    assert rcp_newton(np.array([1.0], f32), ops)[0] == f32(1.0)
