"""Torch-CPU port of the reference's op chain for the hot path.

TEST INFRASTRUCTURE ONLY (see ``arvae_oracle.c``).  This is what
``bench.py`` times as ``cpu_baseline`` / ``--impl reference``: the same ATen
ops the reference dispatches, in the same order, materialising the same
pair-sized temporaries, differentiated by stock autograd -- but able to
take a row block of the pair matrix so a bounded sample of a batch the
reference cannot allocate (B=65536 needs 16 GiB per temporary) can be timed.

Restated from /root/reference:
  utils/trainer.py:369-376  compute_reg_loss       -> :func:`compute_reg_loss`
  utils/trainer.py:378-403  reg_loss_sign          -> :func:`reg_loss_sign`
  utils/trainer.py:354-367  compute_kld_loss       -> :func:`compute_kld_loss`
  imagevae/mnist_vae.py:74-87 reparametrize        -> :func:`reparametrize`
  imagevae/image_vae_trainer.py:171-180 (dim loop) -> :func:`reg_loss_dims`
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


def reg_loss_sign(latent_code: torch.Tensor, attribute: torch.Tensor, factor: float = 1.0,
                  rows: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """Mean over pairs of ``|tanh(factor*(x_i-x_j)) - sign(a_i-a_j)|``.

    Full matrix (``rows=None``): the reference's exact sequence -- column
    vector tiled to [B,B], minus its transpose, flattened; tanh of the scaled
    latent distances; sign of the attribute distances cast to float; L1 mean.

    With ``rows=(r0, r1)``: only rows r0..r1 of both distance matrices are
    built ([r1-r0, B] temporaries) and the sum is divided by B*B, so disjoint
    row blocks add up to the full value.  Gradient w.r.t. the block's own
    ``x_i`` then has to be doubled by the caller to recover the full-matrix
    gradient (the transposed operand's share) -- done in
    :func:`reg_loss_dims_rows_backward`.
    """
    n = latent_code.shape[0]
    if rows is None:
        tiled_x = latent_code.view(-1, 1).repeat(1, n)
        dx = (tiled_x - tiled_x.transpose(1, 0)).view(-1, 1)
        tiled_a = attribute.view(-1, 1).repeat(1, n)
        da = (tiled_a - tiled_a.transpose(1, 0)).view(-1, 1)
        t = torch.tanh(dx * factor)
        s = torch.sign(da).float()
        return torch.nn.L1Loss()(t, s)
    r0, r1 = rows
    m = r1 - r0
    tiled_x = latent_code[r0:r1].view(-1, 1).repeat(1, n)
    dx = (tiled_x - latent_code.detach().view(1, -1).repeat(m, 1)).view(-1, 1)
    tiled_a = attribute[r0:r1].view(-1, 1).repeat(1, n)
    da = (tiled_a - attribute.view(1, -1).repeat(m, 1)).view(-1, 1)
    t = torch.tanh(dx * factor)
    s = torch.sign(da).float()
    return torch.nn.L1Loss(reduction="sum")(t, s) / (float(n) * float(n))


def compute_reg_loss(z: torch.Tensor, labels: torch.Tensor, reg_dim: int, gamma: float,
                     factor: float = 1.0) -> torch.Tensor:
    """gamma * reg_loss_sign(z[:, reg_dim], labels, factor)."""
    return gamma * reg_loss_sign(z[:, reg_dim], labels, factor=factor)


def reg_loss_dims(z: torch.Tensor, labels: torch.Tensor, reg_dims: Sequence[int], gamma: float,
                  factor: float = 1.0) -> torch.Tensor:
    """The trainers' loop: one compute_reg_loss per regularised dim, label
    column ``dim`` paired with latent ``dim``."""
    total = 0.0
    for dim in reg_dims:
        total = total + compute_reg_loss(z, labels[:, dim], dim, gamma=gamma, factor=factor)
    return total


def reg_loss_dims_rows_fwdbwd(z: torch.Tensor, labels: torch.Tensor, reg_dims: Sequence[int],
                              gamma: float, factor: float, rows: Tuple[int, int]
                              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Forward + backward of the dim loop restricted to a row block: returns
    the block's share of the loss and d(full loss)/dz for the block's rows."""
    zr = z.detach().clone().requires_grad_(True)
    total = 0.0
    for dim in reg_dims:
        total = total + gamma * reg_loss_sign(zr[:, dim], labels[:, dim], factor, rows=rows)
    total.backward()
    r0, r1 = rows
    return total.detach(), 2.0 * zr.grad[r0:r1]


def reparametrize(loc: torch.Tensor, scale: torch.Tensor, eps: Optional[torch.Tensor] = None):
    """Normal(loc, scale).rsample() then the (unused) prior draw, in the
    reference's RNG order: eps first, z_prior second."""
    dist = torch.distributions.Normal(loc=loc, scale=scale)
    if eps is None:
        z_tilde = dist.rsample()
    else:
        z_tilde = loc + eps * scale
    prior = torch.distributions.Normal(loc=torch.zeros_like(loc), scale=torch.ones_like(scale))
    z_prior = prior.sample()
    return z_tilde, z_prior, dist, prior


def compute_kld_loss(z_dist, prior_dist, beta: float, c=0.0) -> torch.Tensor:
    """beta * |mean_b sum_d KL(z_dist || prior) - c|."""
    kl = torch.distributions.kl.kl_divergence(z_dist, prior_dist)
    return beta * (kl.sum(1).mean() - c).abs()
