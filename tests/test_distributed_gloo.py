"""Host-side logic of the row-block sharded loss (arvae_b200/distributed.py) on CPU.

world_size-2 ``gloo`` processes run the real packing / all-gather / all-reduce / autograd wiring;
only the device call behind ``_rows_backend`` (the C ABI, CUDA-only) is replaced by the CPU oracle,
because this container has no GPU.  The GPU version of the same check is
tests/test_gpu_multi.py (``-m gpu``, needs >= 2 devices).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_rows_backend(packed, R, gamma, factor, row_begin, row_end, want_grad, algo):
    import oracle
    p = packed.detach().cpu().numpy()
    loss, grad = oracle.compute_reg_loss_multi(p[:, :R], p[:, R:], list(range(R)), gamma, factor, f64=True,
                                               row_begin=row_begin, row_end=row_end)
    loss64 = torch.tensor(loss, dtype=torch.float64)
    return loss64, (torch.from_numpy(grad[:, :R].astype(np.float32)) if want_grad else None)


def _oracle_scatter(grad_cols, grad_out, reg_dims, n, Z):
    out = torch.zeros(n, Z, dtype=torch.float32)
    for r, d in enumerate(reg_dims):
        out[:, d] += grad_cols[:, r]
    return out * (1.0 if grad_out is None else float(grad_out))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arvae_b200 import distributed as adist, synth
        adist._rows_backend = _oracle_rows_backend   # test-only: CPU stand-in for the CUDA call
        adist._scatter = _oracle_scatter
        torch.set_num_threads(2)
        c = synth.make_case("c2_dsprites_b4096", B=384)
        n = c["B"] // world
        z_local = c["z"][rank * n:(rank + 1) * n].clone().requires_grad_(True)
        lab_local = c["labels"][rank * n:(rank + 1) * n].clone()
        # bypass the CUDA-only dtype/device guard of the public wrapper: call the autograd node directly
        dims = tuple(c["reg_dims"])
        loss = adist._ShardedRegLossFn.apply(z_local, lab_local, dims, dims, c["gamma"], c["delta"], None, 0, None, 1.0)
        (loss * 3.0).backward()
        # the same step under DistributedDataParallel with ddp_average (grad_scale = world): the rank-averaged
        # parameter gradient must be the single-process global-batch one
        torch.manual_seed(7)
        enc = torch.nn.Linear(6, c["z"].shape[1], bias=False)
        ddp = torch.nn.parallel.DistributedDataParallel(enc)
        x_all = torch.randn(c["B"], 6, generator=torch.Generator().manual_seed(11))
        z2 = ddp(x_all[rank * n:(rank + 1) * n])
        loss2 = adist._ShardedRegLossFn.apply(z2, lab_local, dims, dims, c["gamma"], c["delta"], None, 0, None, float(world))
        loss2.backward()
        q.put((rank, float(loss), z_local.grad.numpy(), float(loss2), enc.weight.grad.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_reproduces_single_rank_result(oracle_mod):
    from arvae_b200 import synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=150) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    c = synth.make_case("c2_dsprites_b4096", B=384)
    ref_loss, ref_grad = oracle_mod.compute_reg_loss_multi(c["z"].numpy(), c["labels"].numpy(), c["reg_dims"],
                                                           c["gamma"], c["delta"], f64=True)
    # every rank holds the same global loss
    assert results[0][1] == results[1][1]
    assert abs(results[0][1] - ref_loss) <= 1e-6 * abs(ref_loss)
    # each rank's gradient rows are its block of the single-rank gradient (times the upstream 3.0)
    got = np.concatenate([r[2] for r in results], axis=0)
    assert np.allclose(got, 3.0 * ref_grad.astype(np.float32), rtol=1e-6, atol=1e-9)
    # DDP + ddp_average: W.grad (averaged over ranks by DDP) == x^T dL/dz of the global batch on one process
    torch.manual_seed(7)
    enc = torch.nn.Linear(6, c["z"].shape[1], bias=False)
    x_all = torch.randn(c["B"], 6, generator=torch.Generator().manual_seed(11))
    z_all = enc(x_all).detach()
    l2, g2 = oracle_mod.compute_reg_loss_multi(z_all.numpy(), c["labels"].numpy(), c["reg_dims"], c["gamma"], c["delta"],
                                               f64=True)
    w_ref = g2.T @ x_all.numpy().astype(np.float64)
    for r in results:
        assert abs(r[3] - l2) <= 1e-6 * abs(l2)
        assert np.allclose(r[4], w_ref, rtol=2e-5, atol=1e-7 * np.abs(w_ref).max())
    assert np.array_equal(results[0][4], results[1][4])


def test_pack_columns_layout():
    from arvae_b200 import distributed as adist
    z = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    lab = 100 + torch.arange(15, dtype=torch.float32).reshape(3, 5)
    p = adist.pack_columns(z, lab, (1, 3), (2, 4))
    assert p.shape == (3, 4) and p.is_contiguous()
    assert torch.equal(p[:, 0], z[:, 1]) and torch.equal(p[:, 1], z[:, 3])
    assert torch.equal(p[:, 2], lab[:, 2]) and torch.equal(p[:, 3], lab[:, 4])
