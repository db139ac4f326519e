#!/usr/bin/env python
"""Forward+backward time of the reg loss on every BASELINE.json config shape (and a tie-heavy large batch),
next to the reference's own op chain executed by stock PyTorch on the SAME GPU where it fits in memory.
Prints one JSON object per line; used for profiles/ (not the headline bench)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import arvae_b200  # noqa: E402
from arvae_b200 import synth  # noqa: E402
from oracle import torch_port  # noqa: E402  (bench/measurement use of the reference port)


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


cases = [("c1_mnist_b64", None, "morpho"), ("c2_dsprites_b4096", None, "dsprites"), ("c3_measure_b2048", None, "music"),
         ("c4_mnist_b65536", 8192, "morpho"), ("c4_mnist_b65536", None, "morpho"), ("c2_dsprites_b4096", 65536, "dsprites"),
         ("c3_measure_b2048", 65536, "music"), ("c4_mnist_b65536", 262144, "morpho")]
for name, B, kind in cases:
    c = synth.make_case(name, B)
    z, lab = c["z"].cuda(), c["labels"].cuda()
    dims, gamma, delta = c["reg_dims"], c["gamma"], c["delta"]
    pairs = float(c["B"]) ** 2 * len(dims)

    def ours():
        zz = z.detach().requires_grad_(True)
        arvae_b200.reg_loss_fused(zz, lab, dims, gamma, delta).backward()

    def per_dim_calls():  # exactly what the trainers do: one call per dim
        zz = z.detach().requires_grad_(True)
        tot = 0.0
        for d in dims:
            tot = tot + arvae_b200.compute_reg_loss(zz, lab[:, d], d, gamma, delta)
        tot.backward()

    def stock():
        zz = z.detach().requires_grad_(True)
        torch_port.reg_loss_dims(zz, lab, dims, gamma, delta).backward()

    iters = 50 if c["B"] <= 8192 else (10 if c["B"] <= 65536 else 3)
    row = {"config": name, "B": c["B"], "R": len(dims), "labels": kind, "delta": delta,
           "ours_fused_ms": timeit(ours, iters), "ours_per_dim_calls_ms": timeit(per_dim_calls, iters)}
    if c["B"] <= 8192:  # launch-latency-bound sizes: CUDA-graph replay of the same forward + backward
        from arvae_b200 import graphs
        step = graphs.graphed_reg_loss(c["B"], z.shape[1], lab.shape[1], dims, gamma, delta)

        def graphed():
            zz = z.detach().requires_grad_(True)
            step(zz, lab).backward()
        row["ours_graphed_ms"] = timeit(graphed, iters)
    if c["B"] <= 8192:
        try:
            row["stock_torch_cuda_ms"] = timeit(stock, 5)
        except RuntimeError as e:  # OOM
            row["stock_torch_cuda_ms"] = None
            torch.cuda.empty_cache()
    else:
        row["stock_torch_cuda_ms"] = None  # B^2 temporaries do not fit (16 GiB each at B=65536)
    row["ours_gpairs_s"] = pairs / (row["ours_fused_ms"] * 1e-3) / 1e9
    row["mufu_per_pair_by_dim"] = list(arvae_b200.mufu_per_pair(z, lab, dims, gamma, delta))
    print(json.dumps(row), flush=True)
