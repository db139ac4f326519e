#!/usr/bin/env python
"""A few emulated sharded steps (G virtual ranks on one GPU) for profilers: python bench_tools/shard_one.py [world] [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arvae_b200 import synth, distributed as adist
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = synth.make_case("c4_mnist_b65536", B)
dims = tuple(c["reg_dims"]); n = B // G
z, lab = c["z"].cuda(), c["labels"].cuda()
grp = adist.LocalShardGroup(G, n, len(dims))
for it in range(iters):
    outs = grp.step([z[g*n:(g+1)*n] for g in range(G)], [lab[g*n:(g+1)*n] for g in range(G)], dims, dims, c["gamma"], c["delta"])
torch.cuda.synchronize()
print("loss", outs[0][0].item())
grp.close()
