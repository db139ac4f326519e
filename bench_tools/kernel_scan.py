#!/usr/bin/env python
"""Pair-kernel time vs number of rows (row-block shards 1/1 .. 1/32 of C4) using the library's own
event hooks: separates per-launch fixed cost from per-pair cost."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arvae_b200 import _lib, ops, synth

lib = _lib.load()
c = synth.make_case("c4_mnist_b65536")
z, lab = c["z"].cuda(), c["labels"].cuda()
dims = c["reg_dims"]; R = len(dims); B = c["B"]
packed = ops.pack_columns(z, lab, dims, dims)
for shards in (1, 8):
    n = B // shards
    for off in (0, B - n):
        lib.arvae_profile_enable(1)
        for _ in range(8):
            ops.reg_loss_rows(packed[:, :R], packed[:, R:], tuple(range(R)), c["gamma"], c["delta"], off, off + n)
        torch.cuda.synchronize()
        s, k = ctypes.c_float(), ctypes.c_int()
        lib.arvae_profile_pair_kernel_ms(ctypes.byref(s), ctypes.byref(k))
        lib.arvae_profile_enable(0)
        us = s.value / k.value * 1e3
        print(f"rows {n:6d} (offset {off:6d}): pair kernel {us:8.1f} us  -> {us / n * 1e3:7.2f} ns/row  ideal-share ratio {us / (6310.0 / shards):.3f}")
