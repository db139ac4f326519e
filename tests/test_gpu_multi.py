"""Row-block sharding on real GPUs (needs >= 2 devices; skipped otherwise): every world size gives
the single-GPU loss and gradient.  On the dense path (B < 8192) each rank's gradient rows are BITWISE
identical to the same rows of the single-GPU run (each row is owned by one rank and sweeps the same
columns in the same order); on the attribute-sorted path the row tiles differ per partition, so
equality holds to fp32 rounding (2e-6 of the column max)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from arvae_b200 import distributed as adist, synth
        c = synth.make_case("c4_mnist_b65536", B=B)
        n = B // world
        z_local = c["z"][rank * n:(rank + 1) * n].cuda().requires_grad_(True)
        lab_local = c["labels"][rank * n:(rank + 1) * n].cuda()
        loss = adist.reg_loss_sharded(z_local, lab_local, c["reg_dims"], c["gamma"], c["delta"])
        loss.backward()
        torch.cuda.synchronize()
        q.put((rank, float(loss), z_local.grad.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("B", [4096, 16384])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world, B):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import arvae_b200
    from arvae_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + world + (B // 4096)
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = synth.make_case("c4_mnist_b65536", B=B)
    z = c["z"].cuda().requires_grad_(True)
    loss = arvae_b200.reg_loss_fused(z, c["labels"].cuda(), c["reg_dims"], c["gamma"], c["delta"])
    loss.backward()
    for r in results:
        assert r[1] == results[0][1]
        assert abs(r[1] - loss.item()) <= 1e-6 * abs(loss.item())
    got = np.concatenate([r[2] for r in results], axis=0)
    ref = z.grad.cpu().numpy()
    if B < 8192:
        assert np.array_equal(got, ref)
    else:
        scale = np.abs(ref).max(axis=0) + 1e-30
        assert np.all(np.abs(got - ref).max(axis=0) <= 2e-6 * scale)
