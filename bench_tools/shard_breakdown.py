#!/usr/bin/env python
"""Times one rank's share of the row-block sharded step on ONE GPU (no NCCL): rows [0, B/G) against all
B columns, with CUDA events around each phase.  Used to see what bounds multi-GPU scaling besides the
pair kernel (sort, gather, row selection, epilogue, scatter).  Run under ncu for a per-kernel list."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import arvae_b200  # noqa: E402
from arvae_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shards", type=int, default=8)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()

c = synth.make_case("c4_mnist_b65536", B=args.batch)
B, G = c["B"], args.shards
n = B // G
z, lab = c["z"].cuda(), c["labels"].cuda()
dims = c["reg_dims"]
R = len(dims)
packed = ops.pack_columns(z, lab, dims, dims)


def ev():
    return torch.cuda.Event(enable_timing=True)


tot = {"pack": 0.0, "fwdbwd": 0.0, "scatter": 0.0}
for it in range(args.iters + 3):
    e = [ev() for _ in range(4)]
    e[0].record()
    p = ops.pack_columns(z[:n], lab[:n], dims, dims)
    e[1].record()
    loss, gc, _ = ops.reg_loss_rows(packed[:, :R], packed[:, R:], tuple(range(R)), c["gamma"], c["delta"], 0, n)
    e[2].record()
    gz = ops._scatter_bwd(gc, None, dims, n, z.shape[1])
    e[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        tot["pack"] += e[0].elapsed_time(e[1])
        tot["fwdbwd"] += e[1].elapsed_time(e[2])
        tot["scatter"] += e[2].elapsed_time(e[3])
print({k: round(v / args.iters * 1e3, 1) for k, v in tot.items()}, "us per step; shard rows", n, "of", B)
