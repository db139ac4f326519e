"""GPU parity tests: the CUDA path (through the C ABI) against

* the committed golden vectors (outputs of the unmodified reference, tests/golden/), and
* the CPU oracle (oracle/, float64 mode) on the same seeded inputs, including batch sizes the
  reference cannot allocate, where row samples and size-independent properties are used.

Bars (tests/util.py): loss 1e-5 relative; gradient 1e-5 of the column max; sign matrix bit-exact.
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden, golden_names
from util import assert_grad_close, assert_loss_close

pytestmark = pytest.mark.gpu

ALGOS = [1]  # ARVAE_ALGO_DENSE; the sorted path adds itself below when it exists
try:
    from arvae_b200 import ops as _ops
    if getattr(_ops, "HAVE_SORTED", False):
        ALGOS.append(2)
    if getattr(_ops, "HAVE_TRIANGLE", False):
        ALGOS.append(3)
except Exception:  # pragma: no cover
    pass


@pytest.fixture(scope="module")
def ab():
    import arvae_b200
    from arvae_b200 import _lib
    assert torch.cuda.is_available(), "these tests need the B200"
    _lib.load()  # fail loudly if the extension is missing
    return arvae_b200


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------------------------
# golden vectors from the unmodified reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("reg_c"))
def test_fused_dim_loop_matches_reference(ab, name, algo):
    g = golden(name)
    z = dev(g["z"]).requires_grad_(True)
    labels = dev(g["labels"])
    dims = tuple(int(d) for d in g["reg_dims"])
    loss = ab.reg_loss_fused(z, labels, dims, float(g["gamma"]), float(g["delta"]), algo=algo)
    assert loss.dtype == torch.float32 and loss.dim() == 0
    loss.backward()
    assert_loss_close(loss.item(), g["loss"])
    assert_grad_close(z.grad.cpu().numpy(), g["grad_z"])
    untouched = [d for d in range(z.shape[1]) if d not in dims]
    assert not z.grad[:, untouched].any()


@pytest.mark.parametrize("name", ["reg_c1_mnist_b64", "reg_c2_dsprites_b512"])
def test_per_dim_calls_like_the_trainers(ab, name):
    """imagevae/image_vae_trainer.py:171-180: one call per dim with the strided view labels[:, dim]."""
    g = golden(name)
    z = dev(g["z"]).requires_grad_(True)
    labels = dev(g["labels"])
    gamma, delta = float(g["gamma"]), float(g["delta"])
    reg_loss = 0.0
    for k, dim in enumerate(int(d) for d in g["reg_dims"]):
        view = labels[:, dim]
        assert not view.is_contiguous() or labels.shape[1] == 1
        one = ab.compute_reg_loss(z, view, dim, gamma=gamma, factor=delta)
        assert_loss_close(one.item(), g["per_dim_loss"][k])
        reg_loss += one
    reg_loss.backward()
    assert_loss_close(reg_loss.item(), g["loss"])
    assert_grad_close(z.grad.cpu().numpy(), g["grad_z"])


def test_single_call_negative_dim(ab):
    g = golden("reg_single_negdim")
    z = dev(g["z"]).requires_grad_(True)
    labels = dev(g["labels"])
    loss = ab.compute_reg_loss(z, labels[:, int(g["label_col"])], int(g["reg_dim"]), float(g["gamma"]),
                               factor=float(g["delta"]))
    loss.backward()
    assert_loss_close(loss.item(), g["loss"])
    assert_grad_close(z.grad.cpu().numpy(), g["grad_z"])


def test_tuple_reg_dim_extension_and_tensor_scalars(ab):
    g = golden("reg_c1_mnist_b64")
    z = dev(g["z"])
    labels = dev(g["labels"])
    dims = tuple(int(d) for d in g["reg_dims"])
    loss = ab.compute_reg_loss(z, labels, dims, torch.tensor(float(g["gamma"])), torch.tensor([float(g["delta"])]))
    assert not loss.requires_grad
    assert_loss_close(loss.item(), g["loss"])


# ------------------------------------------------------------------------------------------------
# sign matrix: bit-exact
# ------------------------------------------------------------------------------------------------
def test_sign_matrix_bit_exact_vs_reference(ab, oracle_mod):
    g = golden("sign_matrix_special")
    s = ab.sign_matrix(dev(g["a"])).cpu().numpy()
    assert s.dtype == np.int8
    assert np.array_equal(s, g["sign"])
    assert np.array_equal(s, oracle_mod.sign_matrix(g["a"]))
    x = dev(g["x"]).requires_grad_(True)
    loss = ab.reg_loss_sign(x, dev(g["a"]), factor=float(g["delta"]))
    loss.backward()
    assert_loss_close(loss.item(), g["loss"])
    assert_grad_close(x.grad.cpu().numpy(), g["grad_x"])


def test_sign_matrix_bit_exact_on_dataset_like_labels(ab, oracle_mod):
    from arvae_b200 import synth
    for kind, col in (("dsprites", 1), ("dsprites", 3), ("music", 0), ("morpho", 4)):
        lab = synth.make_labels(kind, 700, 7)[:, col].contiguous()
        got = ab.sign_matrix(lab.cuda()).cpu().numpy()
        assert np.array_equal(got, oracle_mod.sign_matrix(lab.numpy()))


def _sign_row_sums_numpy(a):
    """sum_j sign(a_i - a_j) in O(B log B): (# a_j < a_i) - (# a_j > a_i); NaN compares false both ways."""
    a = np.asarray(a, dtype=np.float32)
    valid = np.sort(a[~np.isnan(a)])
    out = np.zeros(a.shape[0], dtype=np.int64)
    ok = ~np.isnan(a)
    out[ok] = np.searchsorted(valid, a[ok], "left") - (valid.size - np.searchsorted(valid, a[ok], "right"))
    return out


@pytest.mark.parametrize("algo", [1, 2])
def test_hot_path_sign_sums_bit_exact_all_rows(ab, oracle_mod, algo):
    """The integer check of the attribute sign matrix THROUGH the pair kernels: every row's sum_j sign(a_i - a_j),
    accumulated by the dense loop's compares / the sorted kernel's tile classes while they evaluate the loss, equals
    the row sums of the oracle's sign matrix (reference utils/trainer.py:394-395,400) exactly -- on tie-heavy
    dSprites grids, NaN / +-inf / +-0 / subnormal labels, and with an outlier segment in the sorted order."""
    from arvae_b200 import synth
    B = 4096
    lab = synth.make_labels("dsprites", B, 11)[:, 1:6].contiguous()
    g = golden("sign_matrix_special")
    special = torch.from_numpy(np.resize(g["a"], B).astype(np.float32))
    special[torch.randperm(B, generator=torch.Generator().manual_seed(1))[:B // 2]] = 0.25
    lab = torch.cat([lab, special[:, None], synth.make_labels("morpho", B, 5)[:, 4:5]], dim=1).contiguous()
    R = lab.shape[1]
    z = torch.randn(B, R, generator=torch.Generator().manual_seed(2))
    z[::97, 3] = 30.0   # outliers of the factorised tanh: their own segment of the sorted order
    _, grad_cols, _, row_sign = ab.reg_loss_rows(z.cuda(), lab.cuda(), tuple(range(R)), 1.0, 1.0, 0, B, algo=algo,
                                                 want_row_sign=True)
    got = row_sign.cpu().numpy()
    for r in range(R):
        a = lab[:, r].numpy()
        assert np.array_equal(got[:, r], _sign_row_sums_numpy(a)), r
    s_full = oracle_mod.sign_matrix(lab[:, R - 2].numpy()).astype(np.int64).sum(axis=1)  # the special-value column
    assert np.array_equal(got[:, R - 2], s_full)
    # row blocks (the NCCL sharding unit) give the same integers
    _, _, _, part = ab.reg_loss_rows(z.cuda(), lab.cuda(), tuple(range(R)), 1.0, 1.0, 1000, 3000, algo=algo,
                                     want_row_sign=True)
    assert np.array_equal(part.cpu().numpy(), got[1000:3000])


def test_hot_path_sign_sums_bit_exact_full_size(ab):
    """Same integers at the full C4 size (B = 65 536, R = 6), where no dense sign matrix fits: every row."""
    from arvae_b200 import synth
    c = synth.make_case("c4_mnist_b65536")
    lab = c["labels"].clone()
    lab[:, 2] = torch.round(lab[:, 2])          # a tie-heavy column
    lab[123, 3] = float("nan"); lab[60000, 3] = float("nan"); lab[7, 4] = float("inf")
    B = c["B"]
    _, _, _, row_sign = ab.reg_loss_rows(c["z"].cuda(), lab.cuda(), c["reg_dims"], c["gamma"], c["delta"], 0, B,
                                         want_row_sign=True)
    got = row_sign.cpu().numpy()
    for r, dim in enumerate(c["reg_dims"]):
        assert np.array_equal(got[:, r], _sign_row_sums_numpy(lab[:, dim].numpy())), dim


# ------------------------------------------------------------------------------------------------
# edge cases the reference handles
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", golden_names("edge_"))
def test_edge_cases(ab, name, algo):
    g = golden(name)
    x = dev(g["x"], torch.float32).requires_grad_(True)
    a = dev(g["a"])  # int64 / float64 labels stay in their dtype: sign is taken there
    if algo == 1:
        loss = ab.reg_loss_sign(x, a, factor=float(g["delta"]))
    else:
        loss = ab.reg_loss_fused(x.reshape(-1, 1), a.reshape(-1, 1), (0,), 1.0, float(g["delta"]), algo=algo)
    loss.backward()
    assert_loss_close(loss.item(), g["loss"])
    assert_grad_close(x.grad.cpu().numpy(), g["grad_x"])
    if name in ("edge_all_equal_both", "edge_b1"):
        assert loss.item() == 0.0 and not x.grad.any()


def test_empty_batch_gives_nan_like_the_reference(ab):
    z = torch.zeros(0, 4, device="cuda")
    lab = torch.zeros(0, 4, device="cuda")
    assert torch.isnan(ab.reg_loss_fused(z, lab, (1, 2), 1.0, 1.0))
    assert torch.isnan(ab.compute_reg_loss(z, lab[:, 0], 0, 1.0))


def test_length_mismatch_raises_runtime_error(ab):
    z = torch.zeros(6, 4, device="cuda")
    with pytest.raises(RuntimeError, match="must match"):
        ab.compute_reg_loss(z, torch.zeros(5, device="cuda"), 0, 1.0)


# ------------------------------------------------------------------------------------------------
# against the float64 oracle on seeded inputs (sizes the oracle finishes in seconds)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("B,delta,kind", [(1537, 1.0, "morpho"), (2048, 10.0, "music"), (3000, 1.0, "dsprites"),
                                          (257, 0.05, "morpho"), (1024, 25.0, "dsprites")])
def test_against_f64_oracle(ab, oracle_mod, B, delta, kind, algo):
    from arvae_b200 import synth
    g = torch.Generator().manual_seed(B)
    labels = synth.make_labels(kind, B, B + 1)
    Z = labels.shape[1]
    z = torch.randn(B, Z, generator=g)
    dims = tuple(range(1, Z)) if kind != "music" else tuple(range(Z))
    ref_loss, ref_grad = oracle_mod.compute_reg_loss_multi(z.numpy(), labels.numpy(), dims, 2.0, delta, f64=True)
    zc = z.cuda().requires_grad_(True)
    loss = ab.reg_loss_fused(zc, labels.cuda(), dims, 2.0, delta, algo=algo)
    loss.backward()
    assert_loss_close(loss.item(), ref_loss)
    assert_grad_close(zc.grad.cpu().numpy(), ref_grad)


@pytest.mark.parametrize("algo", ALGOS)
def test_random_labels_worst_case_for_cancellation(ab, oracle_mod, algo):
    """Labels independent of z: row gradient sums cancel to O(sqrt(B)), the hardest case for the
    absolute accuracy of the fp32 partial sums."""
    B = 4096
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, 3, generator=g)
    labels = torch.randn(B, 3, generator=g)
    ref_loss, ref_grad = oracle_mod.compute_reg_loss_multi(z.numpy(), labels.numpy(), (0, 1, 2), 1.0, 1.0, f64=True)
    zc = z.cuda().requires_grad_(True)
    loss = ab.reg_loss_fused(zc, labels.cuda(), (0, 1, 2), 1.0, 1.0, algo=algo)
    loss.backward()
    assert_loss_close(loss.item(), ref_loss)
    assert_grad_close(zc.grad.cpu().numpy(), ref_grad)


# ------------------------------------------------------------------------------------------------
# row blocks (the multi-GPU sharding unit), determinism
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
def test_row_blocks_add_up_and_rows_are_bitwise_identical(ab, algo):
    g = golden("reg_c2_dsprites_b4096")
    z = dev(g["z"])
    labels = dev(g["labels"])
    dims = tuple(int(d) for d in g["reg_dims"])
    gamma, delta = float(g["gamma"]), float(g["delta"])
    B = z.shape[0]
    full_loss, full_grad, full_rows = ab.reg_loss_rows(z, labels, dims, gamma, delta, 0, B, algo=algo,
                                                       want_row_loss=True)
    edges = [0, 1, 513, 2048, 4095, 4096]
    parts, grads, rows = [], [], []
    for r0, r1 in zip(edges[:-1], edges[1:]):
        l, gc, rl = ab.reg_loss_rows(z, labels, dims, gamma, delta, r0, r1, algo=algo, want_row_loss=True)
        parts.append(l.item())
        grads.append(gc)
        rows.append(rl)
    assert abs(sum(parts) - full_loss.item()) <= (1e-12 if algo == 1 else 1e-8) * abs(full_loss.item())
    if algo == 1:
        # dense path: each row sweeps the same columns in the same order whatever the row block
        assert torch.equal(torch.cat(grads, 0), full_grad)
        assert torch.equal(torch.cat(rows, 0), full_rows)
    else:
        # sorted path: the grouping of rows into tiles (hence the tile classes) depends on the row
        # block, so equality holds to fp32 rounding of the partial sums, not bitwise
        assert_grad_close(torch.cat(grads, 0).cpu().numpy(), full_grad.cpu().numpy(), 2e-6)
        assert torch.allclose(torch.cat(rows, 0), full_rows, rtol=1e-6, atol=0)
    assert_loss_close(full_loss.item(), g["loss"])
    # per-row sums add up to the loss
    tot = full_rows.sum().item() * gamma / (B * B)
    assert abs(tot - full_loss.item()) <= 1e-9 * abs(tot)


def test_run_to_run_bitwise_reproducible(ab):
    from arvae_b200 import synth
    c = synth.make_case("c4_mnist_b65536", B=8192)
    z, labels = c["z"].cuda(), c["labels"].cuda()
    outs = []
    for _ in range(3):
        zz = z.clone().requires_grad_(True)
        l = ab.reg_loss_fused(zz, labels, c["reg_dims"], c["gamma"], c["delta"])
        l.backward()
        outs.append((l.clone(), zz.grad.clone()))
    for l, gr in outs[1:]:
        assert torch.equal(l, outs[0][0]) and torch.equal(gr, outs[0][1])


# ------------------------------------------------------------------------------------------------
# full-size config (C4: B=65536, R=6): row samples against the oracle + size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
def test_c4_full_size_row_samples_and_properties(ab, oracle_mod, algo):
    from arvae_b200 import synth
    c = synth.make_case("c4_mnist_b65536")
    z, labels = c["z"], c["labels"]
    B = c["B"]
    dims = c["reg_dims"]
    zc, lc = z.cuda(), labels.cuda()
    loss64, grad_cols, row_loss = ab.reg_loss_rows(zc, lc, dims, c["gamma"], c["delta"], 0, B, algo=algo,
                                                   want_row_loss=True)
    torch.cuda.synchronize()
    grad_cols = grad_cols.cpu().numpy()
    row_loss = row_loss.cpu().numpy()
    # (1) sampled rows vs the float64 oracle: per-row loss sums and gradients
    rows = np.r_[0, 1, 511, 512, 8191, 30000, 65534, 65535, np.random.RandomState(0).randint(0, B, 24)]
    scale = np.abs(grad_cols).max(axis=0)
    for r, dim in enumerate(dims):
        x = z[:, dim].numpy().astype(np.float64)
        a = labels[:, dim].numpy().astype(np.float64)
        for i in rows:
            _, rl, g = oracle_mod.reg_rows(x, a, c["delta"], int(i), int(i) + 1, f64=True)
            assert abs(row_loss[i, r] - rl[0]) <= 1e-6 * rl[0], (i, r, row_loss[i, r], rl[0])
            assert abs(grad_cols[i, r] - c["gamma"] * g[0]) <= 1e-5 * scale[r], (i, r)
    # (2) the loss is the sum of the row sums, and lies in [0, 2*gamma*R]
    tot = row_loss.sum() * c["gamma"] / (float(B) * float(B))
    assert abs(tot - loss64.item()) <= 1e-9 * tot  # per-CTA loss partials carry 2^-24 fixed point
    assert 0.0 < loss64.item() < 2.0 * c["gamma"] * len(dims)
    # (3) antisymmetry: gradients of a pairwise-difference loss sum to zero over the batch
    colsum = grad_cols.astype(np.float64).sum(axis=0)
    assert np.all(np.abs(colsum) <= 1e-5 * scale * np.sqrt(B)), colsum
    # (4) invariance: shifting a latent column by a constant changes nothing but rounding
    z2 = zc.clone()
    z2[:, dims[0]] += 0.5
    loss_shift, _, _ = ab.reg_loss_rows(z2, lc, dims, c["gamma"], c["delta"], 0, B, want_grad=False, algo=algo)
    assert_loss_close(loss_shift.item(), loss64.item(), 2e-6)
    # (5) a monotone transform of the labels leaves the sign matrix, hence everything, unchanged
    l2 = lc.clone()
    l2[:, dims[1]] = l2[:, dims[1]] * 3.0 + 1.0
    loss_mono, g_mono, _ = ab.reg_loss_rows(zc, l2, dims, c["gamma"], c["delta"], 0, B, algo=algo)
    assert torch.equal(loss_mono, loss64)
    assert np.array_equal(g_mono.cpu().numpy(), grad_cols)


def test_bool_and_uint8_labels_are_refused(ab):
    """The reference raises on bool labels (`a - a.T`, utils/trainer.py:395) and, for uint8, takes the sign of a
    WRAPPED difference; neither is computed silently here."""
    z = torch.randn(8, 2, device="cuda")
    with pytest.raises(RuntimeError, match="bool tensors is not supported"):
        ab.compute_reg_loss(z, torch.zeros(8, dtype=torch.bool, device="cuda"), 0, 1.0)
    with pytest.raises(RuntimeError, match="uint8"):
        ab.compute_reg_loss(z, torch.zeros(8, dtype=torch.uint8, device="cuda"), 0, 1.0)
    ab.compute_reg_loss(z, torch.zeros(8, dtype=torch.int8, device="cuda"), 0, 1.0)  # signed narrow ints are exact in f32


def test_c5_largest_batch_properties(ab, oracle_mod):
    """BASELINE.json C5's largest batch (B = 262 144, R = 6: 4.1e11 ordered pairs, a 256 GiB temporary in the reference):
    sampled rows against the float64 oracle, antisymmetry, loss = sum of row sums, bitwise reproducibility."""
    from arvae_b200 import synth
    c = synth.make_case("c4_mnist_b65536", B=262144)
    z, labels, B, dims = c["z"], c["labels"], c["B"], c["reg_dims"]
    zc, lc = z.cuda(), labels.cuda()
    loss64, grad_cols, row_loss = ab.reg_loss_rows(zc, lc, dims, c["gamma"], c["delta"], 0, B, want_row_loss=True)
    loss_b, grad_b, _ = ab.reg_loss_rows(zc, lc, dims, c["gamma"], c["delta"], 0, B)
    assert torch.equal(loss64, loss_b) and torch.equal(grad_cols, grad_b)
    g = grad_cols.cpu().numpy()
    rl = row_loss.cpu().numpy()
    scale = np.abs(g).max(axis=0)
    rows = np.r_[0, 8191, 8192, 131071, B - 1, np.random.RandomState(1).randint(0, B, 6)]
    for r, dim in enumerate(dims[:3]):
        x = z[:, dim].numpy().astype(np.float64)
        a = labels[:, dim].numpy().astype(np.float64)
        for i in rows:
            _, ref_rl, ref_g = oracle_mod.reg_rows(x, a, c["delta"], int(i), int(i) + 1, f64=True)
            assert abs(rl[i, r] - ref_rl[0]) <= 1e-6 * ref_rl[0], (i, r)
            assert abs(g[i, r] - c["gamma"] * ref_g[0]) <= 1e-5 * scale[r], (i, r)
    tot = rl.sum() * c["gamma"] / (float(B) * float(B))
    assert abs(tot - loss64.item()) <= 1e-9 * tot
    assert np.all(np.abs(g.astype(np.float64).sum(axis=0)) <= 1e-5 * scale * np.sqrt(B))


# ------------------------------------------------------------------------------------------------
# latent head: reparametrize + KLD (+ reg) fused
# ------------------------------------------------------------------------------------------------
def test_fused_head_matches_reference(ab):
    g = golden("head_c3_measure_b2048")
    loc = dev(g["loc"]).requires_grad_(True)
    # the encoder's exp() stays stock PyTorch (measurevae/encoder.py:120-123); take it on the CPU like the
    # golden run did so that scale has the same bits (CUDA expf differs from the CPU's in the last ulp)
    scale = torch.exp(torch.from_numpy(g["log_std"])).cuda().requires_grad_(True)
    dims = tuple(int(d) for d in g["reg_dims"])
    z, kld, reg = ab.reparam_kld_reg(loc, scale, dev(g["eps"]), dev(g["labels"]), dims, float(g["beta"]),
                                     float(g["capacity"]), float(g["gamma"]), float(g["delta"]))
    assert np.array_equal(z.detach().cpu().numpy(), g["z_tilde"])  # bit-identical rsample
    assert_loss_close(kld.item(), g["kld_loss"])
    assert_loss_close(reg.item(), g["reg_loss"])
    (kld + reg).backward()
    assert_grad_close(loc.grad.cpu().numpy(), g["grad_loc"])
    assert_grad_close(scale.grad.cpu().numpy(), g["grad_scale"])
    assert_grad_close((scale.grad * scale.detach()).cpu().numpy(), g["grad_log_std"])  # exp backward


def test_one_launch_head_from_log_std_matches_reference(ab):
    """exp(log_std) + rsample + KLD + reg in ONE launch, ONE more for the backward (csrc/head_fused.cu), against the
    unmodified reference's outputs (goldens).  CUDA expf and the CPU's exp differ in the last ulp, so z_tilde is
    compared to 2 ulp of its terms here (loc + eps * scale cancels near zero); the bit-identical check is
    test_fused_head_matches_reference (scale given)."""
    from arvae_b200 import _lib
    lib = _lib.load()
    g = golden("head_c3_measure_b2048")
    dims = tuple(int(d) for d in g["reg_dims"])
    for rep in range(3):  # the same zero-once workspace serves every call
        loc = dev(g["loc"]).requires_grad_(True)
        log_std = dev(g["log_std"]).requires_grad_(True)
        lib.arvae_launch_count(1)
        z, scale, kld, reg = ab.latent_loss_head(loc, log_std, dev(g["eps"]), dev(g["labels"]), dims, float(g["beta"]),
                                                 float(g["capacity"]), float(g["gamma"]), float(g["delta"]))
        assert lib.arvae_launch_count(1) == 1
        np.testing.assert_allclose(z.detach().cpu().numpy(), g["z_tilde"], rtol=3e-7, atol=1e-6)
        np.testing.assert_allclose(scale.detach().cpu().numpy(), np.exp(g["log_std"]), rtol=3e-7)
        assert_loss_close(kld.item(), g["kld_loss"])
        assert_loss_close(reg.item(), g["reg_loss"])
        (kld + reg).backward()
        assert lib.arvae_launch_count(1) == 1
        assert_grad_close(loc.grad.cpu().numpy(), g["grad_loc"])
        assert_grad_close(log_std.grad.cpu().numpy(), g["grad_log_std"])


@pytest.mark.parametrize("name", ["reg_c1_mnist_b64", "reg_c2_dsprites_b512", "reg_c2_dsprites_b4096", "edge_rand_b129",
                                  "edge_nan_inf_labels", "edge_b1", "edge_negative_factor"])
def test_one_launch_head_reg_part_matches_reg_goldens(ab, name):
    """The pair sweep inside the one-launch head against the reference's reg-loss goldens: loc = z, eps = 0."""
    g = golden(name)
    if "z" in g:   # trainer-loop fixtures: z [B, Z], labels [B, A], reg_dims
        z0, labels = dev(g["z"]), dev(g["labels"])
        dims = tuple(int(d) for d in g["reg_dims"])
        gamma, ref = float(g["gamma"]), g["grad_z"]
    else:          # reg_loss_sign fixtures: x [B], a [B]
        z0, labels = dev(g["x"]).reshape(-1, 1), dev(g["a"]).reshape(-1, 1)
        dims, gamma, ref = (0,), 1.0, g["grad_x"].reshape(-1, 1)
    loc = z0.clone().requires_grad_(True)
    z, kld, reg = ab.reparam_kld_reg(loc, torch.ones_like(loc), torch.zeros_like(loc), labels, dims, 0.0, 0.0, gamma,
                                     float(g["delta"]))
    assert torch.equal(z.detach(), z0)
    assert_loss_close(reg.item(), float(g["loss"]))
    reg.backward()
    assert_grad_close(loc.grad.cpu().numpy(), ref)


def test_fused_head_with_decoder_gradient(ab):
    """z_tilde also feeds the decoder: its upstream gradient must flow through the same backward."""
    g = golden("head_c3_measure_b2048")
    w = torch.linspace(-1, 1, g["loc"].size).reshape(g["loc"].shape).cuda()
    outs = []
    for fused in (True, False):
        loc = dev(g["loc"]).requires_grad_(True)
        scale = torch.exp(dev(g["log_std"])).requires_grad_(True)
        dims = tuple(int(d) for d in g["reg_dims"])
        if fused:
            z, kld, reg = ab.reparam_kld_reg(loc, scale, dev(g["eps"]), dev(g["labels"]), dims, 4.0, 50.0, 1.0, 10.0)
        else:  # composable pieces: head node + reg node, glued by autograd
            z, kld_mean = ab.latent_head(loc, scale, dev(g["eps"]))
            kld = 4.0 * (kld_mean - 50.0).abs()
            reg = ab.reg_loss_fused(z, dev(g["labels"]), dims, 1.0, 10.0)
        total = kld + reg + (z * w).sum()
        total.backward()
        outs.append((total.item(), loc.grad.clone(), scale.grad.clone()))
    assert_loss_close(outs[0][0], outs[1][0], 1e-6)
    assert_grad_close(outs[0][1].cpu().numpy(), outs[1][1].cpu().numpy(), 1e-6)
    assert_grad_close(outs[0][2].cpu().numpy(), outs[1][2].cpu().numpy(), 1e-6)
    # and against plain torch autograd for the non-reg part
    loc = dev(g["loc"]).requires_grad_(True)
    scale = torch.exp(dev(g["log_std"])).requires_grad_(True)
    d = torch.distributions.Normal(loc, scale)
    zt = loc + dev(g["eps"]) * scale
    kl = torch.distributions.kl.kl_divergence(d, torch.distributions.Normal(torch.zeros_like(loc), torch.ones_like(scale)))
    ref = 4.0 * (kl.sum(1).mean() - 50.0).abs() + (zt * w).sum()
    ref.backward()
    loc2 = dev(g["loc"]).requires_grad_(True)
    scale2 = torch.exp(dev(g["log_std"])).requires_grad_(True)
    z, kld_mean = ab.latent_head(loc2, scale2, dev(g["eps"]))
    (4.0 * (kld_mean - 50.0).abs() + (z * w).sum()).backward()
    assert_grad_close(loc2.grad.cpu().numpy(), loc.grad.cpu().numpy())
    assert_grad_close(scale2.grad.cpu().numpy(), scale.grad.cpu().numpy())


def test_kld_drop_in_with_capacity_tensor(ab):
    g = golden("kld_c1_capacity_tensor")
    loc = dev(g["loc"]).requires_grad_(True)
    scale = dev(g["scale"]).requires_grad_(True)
    z_dist = torch.distributions.Normal(loc=loc, scale=scale)
    torch.manual_seed(0)
    z_tilde, z_prior, prior = ab.reparametrize(z_dist)
    cap = torch.FloatTensor([float(g["capacity"][0])]).cuda()  # image_vae_trainer.py:94 passes a [1] tensor
    k = ab.compute_kld_loss(z_dist, prior, beta=float(g["beta"]), c=cap)
    assert tuple(k.shape) == (1,)
    assert_loss_close(k.item(), g["kld_loss"][0])
    k.sum().backward()
    assert_grad_close(loc.grad.cpu().numpy(), g["grad_loc"])
    assert_grad_close(scale.grad.cpu().numpy(), g["grad_scale"])
    # RNG order of the reference: eps first, then the unused prior sample
    torch.manual_seed(0)
    eps = torch.distributions.utils._standard_normal(loc.shape, dtype=loc.dtype, device=loc.device)
    zp = torch.normal(torch.zeros_like(loc), torch.ones_like(loc))
    assert torch.equal(z_tilde.detach(), (loc + eps * scale).detach())
    assert torch.equal(z_prior, zp)


# ------------------------------------------------------------------------------------------------
# the host-buffer C-ABI entry (what a non-torch caller binds)
# ------------------------------------------------------------------------------------------------
def test_host_buffer_entry_point(ab):
    from arvae_b200 import _lib
    lib = _lib.load()
    g = golden("reg_c2_dsprites_b4096")
    z = np.ascontiguousarray(g["z"])
    labels = np.ascontiguousarray(g["labels"])
    dims = [int(d) for d in g["reg_dims"]]
    B, Z = z.shape
    loss = ctypes.c_float()
    grad = np.empty_like(z)
    rc = lib.arvae_reg_loss_host_f32(z.ctypes.data_as(ctypes.c_void_p), B, Z,
                                     labels.ctypes.data_as(ctypes.c_void_p), labels.shape[1],
                                     _lib.i32_array(dims), _lib.i32_array(dims), len(dims), float(g["gamma"]),
                                     float(g["delta"]), 0, ctypes.byref(loss),
                                     grad.ctypes.data_as(ctypes.c_void_p), None)
    assert rc == 0, _lib.last_error()
    assert_loss_close(loss.value, g["loss"])
    assert_grad_close(grad, g["grad_z"])
    lib.arvae_host_release()


# ------------------------------------------------------------------------------------------------
# building blocks of the attribute-sorted path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 2, 255, 256, 257, 5000, 8192, 8193, 40000, 65536, 150000])
def test_attr_argsort_matches_stable_sort(ab, B):
    """Sortedness + permutation property at every size class of the bitonic network (single chunk,
    1/2/3 fused global steps), with ties, NaN, +-inf and +-0 in the keys."""
    g = torch.Generator().manual_seed(B)
    a = torch.randint(-50, 50, (B,), generator=g).float() / 7.0 if B % 2 else torch.randn(B, generator=g)
    if B > 20:
        a[3] = float("nan"); a[7] = float("inf"); a[11] = float("-inf"); a[13] = 0.0; a[17] = -0.0; a[19] = float("nan")
    perm = ab.attr_argsort(a.cuda()).cpu().long()
    assert sorted(perm.tolist()) == list(range(B))            # a permutation
    s = a[perm]
    nan = torch.isnan(s)
    n_nan = int(nan.sum())
    assert not nan[:B - n_nan].any() and nan[B - n_nan:].all()  # NaN last
    v = s[:B - n_nan]
    assert bool((v[1:] >= v[:-1]).all())                       # ascending
    # ties (incl. -0 / +0 which compare equal but have distinct keys) keep a deterministic order:
    # equal BIT patterns come out by increasing original index
    bits = a.view(torch.int32)[perm]
    same = bits[1:] == bits[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())


def test_pack_columns(ab):
    z = torch.randn(300, 9, device="cuda")
    lab = torch.randn(300, 5, device="cuda")
    p = ab.pack_columns(z, lab, (1, 8, 3), (0, 4, 2))
    ref = torch.cat([z[:, [1, 8, 3]], lab[:, [0, 4, 2]]], dim=1)
    assert torch.equal(p, ref)
    zt = torch.randn(9, 300, device="cuda").t()  # non-contiguous rows
    assert torch.equal(ab.pack_columns(zt, lab, (2,), (1,)), torch.cat([zt[:, [2]], lab[:, [1]]], dim=1))


def test_outliers_of_the_one_mufu_form_get_their_own_segment(ab, oracle_mod):
    """|2 f log2(e) z| > 62 for a sample -> it sorts into the dim's outlier segment and only ITS pairs use EX2+RCP on
    the latent difference; a single outlier no longer sends the whole dim to two MUFU.  Results match the oracle."""
    B = 9000
    g = torch.Generator().manual_seed(9)
    z = torch.randn(B, 3, generator=g)
    z[5, 1] = 30.0           # 2.885 * 30 > 62
    labels = torch.randn(B, 3, generator=g)
    per_dim = ab.mufu_per_pair(z.cuda(), labels.cuda(), (0, 1, 2), 1.0, 1.0)
    assert per_dim[0] == 1.0 and per_dim[2] == 1.0
    assert abs(per_dim[1] - (2.0 - ((B - 1) / B) ** 2)) < 1e-12
    ref_loss, ref_grad = oracle_mod.compute_reg_loss_multi(z.numpy(), labels.numpy(), (0, 1, 2), 1.0, 1.0, f64=True)
    zc = z.cuda().requires_grad_(True)
    loss = ab.reg_loss_fused(zc, labels.cuda(), (0, 1, 2), 1.0, 1.0)
    loss.backward()
    assert_loss_close(loss.item(), ref_loss)
    assert_grad_close(zc.grad.cpu().numpy(), ref_grad)
    # delta = 10 (the MeasureVAE setting, train_measure_vae.py:46-49): |z| > 2.15 is an outlier, ~3 % of N(0,1)
    # samples; ~94 % of the pairs stay on one MUFU
    for m in ab.mufu_per_pair(z.cuda(), labels.cuda(), (0, 2), 1.0, 10.0):
        assert 1.03 < m < 1.10, m
    # the triangle variant needs one attribute order per dim: there the whole dim falls back
    assert set(ab.mufu_per_pair(z.cuda(), labels.cuda(), (0, 2), 1.0, 10.0, algo=3)) == {2.0}


@pytest.mark.parametrize("zmax", [10.7, 10.8, 21.0])
def test_shared_reciprocal_build_and_its_range_boundary(ab, oracle_mod, zmax):
    """The pair kernel's build for the common case shares one reciprocal between two pairs, 1 / (a b) with
    a, b <= 1 + 2^62, which is valid while |2 f log2(e) z| <= 31 for EVERY element (|z| <= 10.74 at delta = 1); one element
    beyond that (still an inlier of the one-MUFU form, |u| <= 62) sends the call to the complete build.  Both sides of
    the boundary, and widely spread latents inside it (products up to 2^124), match the float64 oracle."""
    B = 8192 + 512
    g = torch.Generator().manual_seed(33)
    z = torch.randn(B, 2, generator=g) * 3.0
    z.clamp_(-10.7, 10.7)
    z[7, 0] = zmax
    z[11, 1] = -zmax
    z[13, 0] = -10.7        # with z[7, 0]: 1 + 2^(u_j - u_i) up to 1 + 2^61.7 inside the shared form
    labels = torch.stack([torch.randn(B, generator=g), torch.randint(0, 5, (B,), generator=g).float()], dim=1)
    assert ab.mufu_per_pair(z.cuda(), labels.cuda(), (0, 1), 1.0, 1.0) == (1.0, 1.0)  # nobody is an outlier
    ref_loss, ref_grad = oracle_mod.compute_reg_loss_multi(z.numpy(), labels.numpy(), (0, 1), 1.0, 1.0, f64=True)
    zc = z.cuda().requires_grad_(True)
    loss = ab.reg_loss_fused(zc, labels.cuda(), (0, 1), 1.0, 1.0, algo=2)
    loss.backward()
    assert_loss_close(loss.item(), ref_loss)
    assert_grad_close(zc.grad.cpu().numpy(), ref_grad)


@pytest.mark.parametrize("algo", [2, 3])
@pytest.mark.parametrize("case", ["delta10", "all_outliers", "nan_latent", "boundary_in_tile"])
def test_outlier_segment_against_f64_oracle(ab, oracle_mod, algo, case):
    B = 8192 + 300
    g = torch.Generator().manual_seed(21)
    z = torch.randn(B, 2, generator=g)
    labels = torch.stack([torch.randn(B, generator=g), torch.randint(0, 7, (B,), generator=g).float()], dim=1)
    delta = 10.0
    if case == "all_outliers":
        z = z * 0.01 + 5.0            # every |u| > 62: no inlier at all (and a shift the loss is invariant to)
    elif case == "nan_latent":
        delta = 1.0
        z[17, 0] = float("inf")       # inf - inf = NaN in the reference too
    elif case == "boundary_in_tile":
        delta = 1.0
        z[:100, 1] = 25.0             # exactly 100 outliers: the segment boundary falls inside a tile
    ref_loss, ref_grad = oracle_mod.compute_reg_loss_multi(z.numpy(), labels.numpy(), (0, 1), 1.0, delta, f64=True)
    zc = z.cuda().requires_grad_(True)
    loss = ab.reg_loss_fused(zc, labels.cuda(), (0, 1), 1.0, delta, algo=algo)
    loss.backward()
    assert_loss_close(loss.item(), ref_loss)
    if case == "nan_latent":
        assert np.isnan(ref_loss) and torch.isnan(zc.grad[:, 0]).any()
        assert_grad_close(zc.grad[:, 1].cpu().numpy(), ref_grad[:, 1])
    else:
        assert_grad_close(zc.grad.cpu().numpy(), ref_grad)


# ------------------------------------------------------------------------------------------------
# CUDA-graph capture (small, launch-latency-bound batches)
# ------------------------------------------------------------------------------------------------
def test_graph_captured_reg_loss_matches_eager(ab):
    from arvae_b200 import graphs
    g = golden("reg_c1_mnist_b64")
    dims = tuple(int(d) for d in g["reg_dims"])
    B, Z = g["z"].shape
    step = graphs.graphed_reg_loss(B, Z, g["labels"].shape[1], dims, float(g["gamma"]), float(g["delta"]))
    for trial in range(3):  # replays must pick up new input contents
        gen = torch.Generator().manual_seed(trial)
        z = (dev(g["z"]) if trial == 0 else torch.randn(B, Z, generator=gen).cuda()).requires_grad_(True)
        labels = dev(g["labels"]) if trial == 0 else torch.randn(B, g["labels"].shape[1], generator=gen).cuda()
        loss = step(z, labels)
        loss.backward()
        z2 = z.detach().clone().requires_grad_(True)
        ref = ab.reg_loss_fused(z2, labels, dims, float(g["gamma"]), float(g["delta"]))
        ref.backward()
        assert torch.equal(loss.detach(), ref.detach())
        assert torch.equal(z.grad, z2.grad)
        if trial == 0:
            assert_loss_close(loss.item(), g["loss"])
            assert_grad_close(z.grad.cpu().numpy(), g["grad_z"])


def test_graph_captured_latent_head_matches_eager(ab):
    from arvae_b200 import graphs
    g = golden("head_c3_measure_b2048")
    dims = tuple(int(d) for d in g["reg_dims"])
    B, Z = g["loc"].shape
    head = graphs.graphed_latent_head(B, Z, g["labels"].shape[1], dims, float(g["beta"]), float(g["capacity"]),
                                      float(g["gamma"]), float(g["delta"]))
    loc = dev(g["loc"]).requires_grad_(True)
    scale = torch.exp(torch.from_numpy(g["log_std"])).cuda().requires_grad_(True)
    z, kld, reg = head(loc, scale, dev(g["eps"]), dev(g["labels"]))
    (kld + reg).backward()
    assert np.array_equal(z.detach().cpu().numpy(), g["z_tilde"])
    assert_loss_close(kld.item(), g["kld_loss"])
    assert_loss_close(reg.item(), g["reg_loss"])
    assert_grad_close(loc.grad.cpu().numpy(), g["grad_loc"])
    assert_grad_close(scale.grad.cpu().numpy(), g["grad_scale"])


def test_half_precision_latents_are_upcast_and_float64_is_refused(ab):
    g = golden("reg_c1_mnist_b64")
    labels = dev(g["labels"])
    dims = tuple(int(d) for d in g["reg_dims"])
    for dt in (torch.bfloat16, torch.float16):
        z = dev(g["z"]).to(dt).requires_grad_(True)
        loss = ab.reg_loss_fused(z, labels, dims, float(g["gamma"]), float(g["delta"]))
        assert loss.dtype == dt
        loss.backward()
        assert z.grad.dtype == dt and torch.isfinite(z.grad.float()).all()
        # equals the float32 op on the rounded latents, rounded back
        z32 = z.detach().float().requires_grad_(True)
        ref = ab.reg_loss_fused(z32, labels, dims, float(g["gamma"]), float(g["delta"]))
        ref.backward()
        assert torch.equal(loss.detach(), ref.detach().to(dt))
        assert torch.equal(z.grad, z32.grad.to(dt))
        one = ab.compute_reg_loss(z.detach(), labels[:, 1], 1, 10.0)
        assert one.dtype == dt
    with pytest.raises(RuntimeError, match="float32"):
        ab.reg_loss_fused(dev(g["z"]).double(), labels, dims, 1.0, 1.0)
    # opt-in: float64 latents rounded to float32, results returned as float64 like the reference's (utils/trainer.py:374-376)
    from arvae_b200 import ops
    ops.FLOAT64_POLICY = "compute_in_float32"
    try:
        z64 = dev(g["z"]).double().requires_grad_(True)
        loss64 = ab.reg_loss_fused(z64, labels, dims, float(g["gamma"]), float(g["delta"]))
        loss64.backward()
        assert loss64.dtype == torch.float64 and z64.grad.dtype == torch.float64
        assert_loss_close(loss64.item(), g["loss"])
        assert_grad_close(z64.grad.cpu().numpy(), g["grad_z"])
    finally:
        ops.FLOAT64_POLICY = "raise"
