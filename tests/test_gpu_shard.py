"""The NVLink-sharded step (csrc/reg_shard.cuh) on ONE GPU: G virtual ranks of one process, driven phase by phase
(arvae_b200.distributed.LocalShardGroup).  Every kernel of the multi-GPU path runs -- per-rank sort, publish into every
rank's buffer, merge by binary search, the pair kernel on each rank's CTA range, the pull of row sums and loss
partials -- only the transport is same-device memory instead of NVLink.  The bar: loss and every
gradient element BITWISE equal to the single-GPU op on the concatenated batch, for every world size.
tests/test_gpu_multi.py repeats the check with one process per GPU over real peer memory (needs >= 2 devices)."""
import numpy as np
import pytest
import torch

from util import assert_grad_close, assert_loss_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    import arvae_b200
    from arvae_b200 import _lib
    assert torch.cuda.is_available(), "these tests need the B200"
    _lib.load()
    return arvae_b200


def _single(ab, z, labels, dims, gamma, delta):
    loss64, grad_cols, _ = ab.reg_loss_rows(z, labels, dims, gamma, delta, 0, z.shape[0], algo=ab.ALGO_SORTED)
    return loss64, grad_cols


def _sharded(z, labels, dims, gamma, delta, world, sizes=None):
    from arvae_b200 import distributed as adist
    B = z.shape[0]
    if sizes is None:
        assert B % world == 0
        sizes = [B // world] * world
    offs = np.concatenate([[0], np.cumsum(sizes)])
    grp = adist.LocalShardGroup(world, max(sizes), len(dims))
    try:
        zp = [z[offs[g]:offs[g + 1]] for g in range(world)]
        lp = [labels[offs[g]:offs[g + 1]] for g in range(world)]
        outs = grp.step(zp, lp, dims, dims, gamma, delta)
        torch.cuda.synchronize()
        for h in grp.ranks:
            status, epoch = h.status()
            assert status == 0 and epoch == 1
        # a second step through the same communicator (epochs, accumulator clearing)
        outs2 = grp.step(zp, lp, dims, dims, gamma, delta)
        torch.cuda.synchronize()
        for (l1, _, g1), (l2, _, g2) in zip(outs, outs2):
            assert torch.equal(l1, l2) and torch.equal(g1, g2)
        return [o[0].clone() for o in outs2], torch.cat([o[2] for o in outs2], dim=0)
    finally:
        grp.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("B,kind,delta", [(16384, "morpho", 1.0), (4096, "dsprites", 1.0), (8192, "music", 10.0)])
def test_virtual_ranks_bitwise_equal_single_gpu(ab, world, B, kind, delta):
    from arvae_b200 import synth
    labels = synth.make_labels(kind, B, 3 + B)
    Z = labels.shape[1]
    z = torch.randn(B, Z, generator=torch.Generator().manual_seed(B + world))
    dims = tuple(range(1, Z)) if kind != "music" else tuple(range(Z))
    zc, lc = z.cuda(), labels.cuda()
    ref_loss, ref_grad = _single(ab, zc, lc, dims, 10.0, delta)
    losses, grad = _sharded(zc, lc, dims, 10.0, delta, world)
    for l in losses:
        assert torch.equal(l, ref_loss), (l.item(), ref_loss.item())
    assert torch.equal(grad, ref_grad)


def test_unequal_row_counts_and_ragged_batch(ab, oracle_mod):
    """Ranks with different numbers of rows, B not a multiple of the tile sizes, NaN / inf attributes and an
    out-of-range latent (outlier segment)."""
    B = 5000
    g = torch.Generator().manual_seed(3)
    z = torch.randn(B, 3, generator=g)
    labels = torch.randint(-20, 20, (B, 3), generator=g).float() / 4.0
    labels[5, 0] = float("nan"); labels[4000, 0] = float("nan"); labels[77, 1] = float("inf"); labels[4999, 1] = float("-inf")
    z[100, 2] = 40.0
    z[4500, 2] = -33.0
    zc, lc = z.cuda(), labels.cuda()
    ref_loss, ref_grad = _single(ab, zc, lc, (0, 1, 2), 2.0, 1.0)
    losses, grad = _sharded(zc, lc, (0, 1, 2), 2.0, 1.0, 3, sizes=[1000, 2500, 1500])
    assert all(torch.equal(l, ref_loss) for l in losses)
    assert torch.equal(grad, ref_grad)
    o_loss, o_grad = oracle_mod.compute_reg_loss_multi(z.numpy(), labels.numpy(), (0, 1, 2), 2.0, 1.0, f64=True)
    assert_loss_close(ref_loss.item(), o_loss)
    assert_grad_close(grad.cpu().numpy(), o_grad[:, :3])


def test_sharded_step_matches_reference_golden(ab):
    from conftest import golden
    g = golden("reg_c2_dsprites_b4096")
    z = torch.from_numpy(g["z"]).cuda()
    labels = torch.from_numpy(g["labels"]).cuda()
    dims = tuple(int(d) for d in g["reg_dims"])
    losses, grad = _sharded(z, labels, dims, float(g["gamma"]), float(g["delta"]), 4)
    assert_loss_close(losses[0].item(), g["loss"])
    assert_grad_close(grad.cpu().numpy(), g["grad_z"][:, list(dims)])


@pytest.mark.parametrize("pinned", [True, False])
def test_host_buffer_entry_of_the_sharded_step(ab, pinned):
    """arvae_shard_reg_loss_host_f32 (what bench.py's e2e leg calls on every rank) with a one-rank communicator: host
    buffers in, host loss and dL/dz out.  With a pinned result buffer the finalize kernel writes the gradient straight
    into host memory; with pageable memory it goes through a device buffer and a copy.  Both equal the single-GPU op."""
    import ctypes
    from arvae_b200 import _lib, distributed as adist, synth
    lib = _lib.load()
    c = synth.make_case("c4_mnist_b65536", B=9000)  # not a multiple of any tile size; two sorted runs
    dims = tuple(c["reg_dims"])
    z, lab = c["z"].contiguous(), c["labels"].contiguous()
    B, Z = z.shape
    grad = torch.full((B, Z), float("nan"))
    if pinned:
        z, lab, grad = z.pin_memory(), lab.pin_memory(), grad.pin_memory()
    h = adist._ShardHandle(0, 1, B, len(dims), torch.device("cuda", torch.cuda.current_device()))
    try:
        for step in range(2):
            loss = ctypes.c_float()
            rc = lib.arvae_shard_reg_loss_host_f32(h.ctx, ctypes.c_void_p(z.data_ptr()), Z, ctypes.c_void_p(lab.data_ptr()),
                                                   lab.shape[1], _lib.i32_array(dims), _lib.i32_array(dims), len(dims),
                                                   _lib.i64_array([B]), c["gamma"], c["delta"], ctypes.byref(loss),
                                                   ctypes.c_void_p(grad.data_ptr()), None)
            _lib.check(rc, "arvae_shard_reg_loss_host_f32")
        assert h.status() == (0, 2)
    finally:
        h.close()
    zc = z.cuda().requires_grad_(True)
    ref = ab.reg_loss_fused(zc, lab.cuda(), dims, c["gamma"], c["delta"], algo=ab.ALGO_SORTED)
    ref.backward()
    assert loss.value == ref.item()
    assert torch.equal(grad, zc.grad.cpu())


def test_full_c4_batch_eight_virtual_ranks_bitwise(ab):
    """The benchmark configuration itself (B = 65 536, R = 6) cut over eight ranks: bitwise the single-GPU result."""
    from arvae_b200 import synth
    c = synth.make_case("c4_mnist_b65536")
    zc, lc = c["z"].cuda(), c["labels"].cuda()
    ref_loss, ref_grad = _single(ab, zc, lc, c["reg_dims"], c["gamma"], c["delta"])
    losses, grad = _sharded(zc, lc, c["reg_dims"], c["gamma"], c["delta"], 8)
    assert all(torch.equal(l, ref_loss) for l in losses)
    assert torch.equal(grad, ref_grad)


def test_a_peer_that_never_shows_up_gives_nan_after_a_bounded_wait(ab):
    """Failure detection of the sharded step: rank 0 runs its whole step while rank 1 never publishes.  Its kernels give
    up after the wait bound (ARVAE_SHARD_WAIT_MS), the communicator reports status 1, loss and gradient are NaN -- the
    GPU is not hung."""
    import os
    from arvae_b200 import distributed as adist, synth
    c = synth.make_case("c4_mnist_b65536", B=2048)
    dims = tuple(c["reg_dims"])
    z, lab = c["z"].cuda(), c["labels"].cuda()
    os.environ["ARVAE_SHARD_WAIT_MS"] = "100"
    try:
        grp = adist.LocalShardGroup(2, 1024, len(dims))
        try:
            h = grp.ranks[0]
            loss64, loss32, grad = h.step(z[:1024], lab[:1024], dims, dims, [1024, 1024], c["gamma"], c["delta"], True, 0)
            torch.cuda.synchronize()
            status, epoch = h.status()
            assert status == 1
            assert torch.isnan(loss64) and torch.isnan(grad).all()
        finally:
            grp.close()
    finally:
        os.environ["ARVAE_SHARD_WAIT_MS"] = "30000"
        adist.LocalShardGroup(1, 16, 1).close()  # restores the default bound for the rest of the process
        del os.environ["ARVAE_SHARD_WAIT_MS"]
