#!/usr/bin/env python
"""The one-launch latent-loss head (csrc/head_fused.cu) on C1 / C2 / C3: a few forward+backward steps for profilers
(ncu --metrics gpu__time_duration.sum gives the kernels' own durations):  python bench_tools/head_small_times.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import arvae_b200
from arvae_b200 import synth
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for name, beta in (("c1_mnist_b64", 4.0), ("c2_dsprites_b4096", 4.0), ("c3_measure_b2048", 0.001)):
    c = synth.make_case(name)
    loc0, log_std0, eps0 = synth.make_latent_head(c["B"], c["Z"], 77)
    loc = loc0.cuda().requires_grad_(True)
    scale = torch.exp(log_std0).cuda().requires_grad_(True)
    eps, lab = eps0.cuda(), c["labels"].cuda()
    for it in range(iters):
        loc.grad = scale.grad = None
        z, kld, reg = arvae_b200.reparam_kld_reg(loc, scale, eps, lab, c["reg_dims"], beta, 0.0, c["gamma"], c["delta"])
        (kld + reg).backward()
    torch.cuda.synchronize()
    print(name, float(kld), float(reg))
