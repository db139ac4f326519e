#!/usr/bin/env python
"""A few single-GPU forward calls for profilers: python bench_tools/single_one.py [B] [iters] [delta]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arvae_b200
from arvae_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = synth.make_case("c4_mnist_b65536", B)
delta = float(sys.argv[3]) if len(sys.argv) > 3 else c["delta"]
z, lab = c["z"].cuda(), c["labels"].cuda()
for it in range(iters):
    loss, gc, _ = arvae_b200.reg_loss_rows(z, lab, c["reg_dims"], c["gamma"], delta, 0, B, algo=2)
torch.cuda.synchronize()
print("loss", loss.item())
