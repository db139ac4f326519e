"""CUDA-graph capture of the hot path for small, launch-latency-bound batches.

At the reference's real training sizes (B = 64 ... 256, `train_image_vae.py`, `train_measure_vae.py`) the
pair math is microseconds; what costs is the ~10-20 kernel launches and the Python / autograd dispatch per
step.  Every entry point of the C ABI is stream-ordered, allocation-free and sync-free, so a whole
forward + backward can be captured once and replayed as one graph launch:

    step = arvae_b200.graphs.graphed_reg_loss(B, Z, A, reg_dims, gamma, factor)
    loss = step(z, labels)          # z requires grad; loss.backward() replays the captured backward

Only single-GPU calls are offered: capturing the NCCL collectives of ``distributed.reg_loss_sharded`` through
``make_graphed_callables`` deadlocked on the 2-GPU box, and at large batches replay gains nothing anyway.

Shapes, reg dims, gamma and factor are baked into the graph (they are constants of a training run);
``z`` and ``labels`` contents are free.  Plumbing only: the capture itself is
``torch.cuda.make_graphed_callables``.
"""
from __future__ import annotations

import contextlib
import gc
from typing import Callable, Sequence

import torch

from . import ops


@contextlib.contextmanager
def quiet_gc():
    """Collect now and keep the cyclic garbage collector off while a stream is capturing: if it ran during the capture
    and freed an older CUDA graph / event (graphed callables sit in reference cycles), that free is an operation 'not
    permitted when stream is capturing' and invalidates the capture (seen as an order-dependent test failure)."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def graphed_reg_loss(B: int, Z: int, A: int, reg_dims: Sequence[int], gamma: float, factor: float = 1.0,
                     device="cuda", algo: int = ops.ALGO_AUTO) -> Callable[[torch.Tensor, torch.Tensor], torch.Tensor]:
    """Graph-captured ``reg_loss_fused(z, labels, reg_dims, gamma, factor)`` for fixed shapes [B,Z] / [B,A]."""
    dims = tuple(int(d) for d in reg_dims)

    def fn(z, labels):
        return ops.reg_loss_fused(z, labels, dims, gamma, factor, algo=algo)

    z0 = torch.randn(B, Z, device=device, requires_grad=True)
    l0 = torch.randn(B, A, device=device)  # labels never get a gradient
    with quiet_gc():
        return torch.cuda.make_graphed_callables(fn, (z0, l0), allow_unused_input=True)


def graphed_latent_head(B: int, Z: int, A: int, reg_dims: Sequence[int], beta: float, capacity: float, gamma: float,
                        factor: float = 1.0, device="cuda", algo: int = ops.ALGO_AUTO):
    """Graph-captured ``reparam_kld_reg(loc, scale, eps, labels, ...)`` -> (z_tilde, kld_loss, reg_loss).

    The whole latent-loss head of a train step (imagevae/image_vae_trainer.py:157-180: reparametrize, KLD,
    regularization and their backward) in two graph launches (forward, backward)."""
    dims = tuple(int(d) for d in reg_dims)

    def fn(loc, scale, eps, labels):
        return ops.reparam_kld_reg(loc, scale, eps, labels, dims, beta, capacity, gamma, factor, algo=algo)

    loc = torch.randn(B, Z, device=device, requires_grad=True)
    scale = (torch.rand(B, Z, device=device) + 0.5).requires_grad_(True)
    eps = torch.randn(B, Z, device=device)
    lab = torch.randn(B, A, device=device)
    with quiet_gc():
        return torch.cuda.make_graphed_callables(fn, (loc, scale, eps, lab), allow_unused_input=True)

