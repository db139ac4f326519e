"""Row-block sharding of the pairwise loss across the GPUs of one NVSwitch box.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Rank g holds the
samples its own encoder produced -- ``z_local`` [B/G, Z] and ``labels_local``
[B/G, A] -- and needs every sample's regularised columns as the "columns" of the
pair matrix:

1. all-gather of the packed local slice [B/G, 2R] (latent columns ‖ attribute
   columns; 384 KiB per rank at B=65536, R=6, G=8 -- latency-bound on NVSwitch),
2. the pair kernel over this rank's rows x all B columns x R dims,
3. all-reduce(sum) of one float64 loss partial.

There is no gradient exchange: by antisymmetry the full
gradient of row i is a row sum over all columns, which rank g already has for
its own rows.  The gathered remote columns are therefore constants in autograd.
The returned loss / gradient are those of the GLOBAL-batch loss; DDP's later 1/G
averaging of parameter gradients applies on top, as for any global-batch loss.

The reference has no distributed code; this follows BASELINE.json's
north_star and SURVEY section 8(e).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops


def pack_columns(z_local: torch.Tensor, labels_local: torch.Tensor, reg_dims: Sequence[int],
                 label_cols: Sequence[int]) -> torch.Tensor:
    """[B_local, 2R] float32: the R regularised latent columns, then the R attribute columns."""
    if z_local.is_cuda and labels_local.dtype == torch.float32:
        return ops.pack_columns(z_local.detach(), labels_local.detach(), reg_dims, label_cols)  # one launch
    zc = z_local.detach()[:, list(reg_dims)]  # host-logic tests (gloo, CPU tensors) and exact-cast label dtypes
    lc = labels_local.detach()[:, list(label_cols)].to(torch.float32)
    return torch.cat([zc, lc], dim=1).contiguous()


def gather_columns(packed_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """All-gather equal-sized row blocks; returns ([B, 2R], row offset of this rank)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = packed_local.shape[0]
    out = torch.empty((world * n_local, packed_local.shape[1]), dtype=packed_local.dtype,
                      device=packed_local.device)
    dist.all_gather_into_tensor(out, packed_local, group=group)
    return out, rank * n_local


def _rows_backend(packed: torch.Tensor, R: int, gamma: float, factor: float, row_begin: int, row_end: int,
                  want_grad: bool, algo: int):
    """This rank's rows against all gathered columns, through the C ABI (CUDA only).
    Returns (loss share [()] float64, grad_cols [rows, R] float32 or None)."""
    z_view = packed[:, :R]
    lab_view = packed[:, R:]
    loss64, grad_cols, _ = ops.reg_loss_rows(z_view, lab_view, tuple(range(R)), gamma, factor, row_begin, row_end,
                                             want_grad=want_grad, algo=algo)
    return loss64, grad_cols


class _ShardedRegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_local, labels_local, reg_dims, label_cols, gamma, factor, group, algo):
        R = len(reg_dims)
        packed_local = pack_columns(z_local, labels_local, reg_dims, label_cols)
        packed, row0 = gather_columns(packed_local, group)
        want_grad = bool(ctx.needs_input_grad[0])
        loss64, grad_cols = _rows_backend(packed, R, gamma, factor, row0, row0 + z_local.shape[0], want_grad, algo)
        dist.all_reduce(loss64, op=dist.ReduceOp.SUM, group=group)
        ctx.reg_dims = tuple(reg_dims)
        ctx.shape = tuple(z_local.shape)
        if want_grad:
            ctx.save_for_backward(grad_cols)
        return loss64.to(torch.float32)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (grad_cols,) = ctx.saved_tensors
        n, Z = ctx.shape
        grad_z = _scatter(grad_cols, grad_out, ctx.reg_dims, n, Z)
        return grad_z, None, None, None, None, None, None, None


def _scatter(grad_cols, grad_out, reg_dims, n, Z):
    return ops._scatter_bwd(grad_cols, grad_out, reg_dims, n, Z)


def reg_loss_sharded(z_local: torch.Tensor, labels_local: torch.Tensor, reg_dims: Sequence[int], gamma,
                     factor=1.0, label_cols: Optional[Sequence[int]] = None, group=None,
                     algo: int = ops.ALGO_AUTO) -> torch.Tensor:
    """Global-batch attribute-regularization loss from per-rank row blocks.

    Every rank passes its own ``[B/G, Z]`` latents and ``[B/G, A]`` labels (equal sizes on all
    ranks) and gets the same scalar: the loss the reference would compute on the concatenated
    batch.  ``z_local.grad`` receives d(global loss)/d(z_local).
    """
    if not dist.is_initialized():
        raise RuntimeError("arvae_b200.distributed: torch.distributed is not initialised")
    Z = z_local.shape[1]
    dims = ops._normalize_dims(reg_dims, Z)
    A = labels_local.shape[1]
    lcols = tuple(int(c) % A for c in (dims if label_cols is None else label_cols))
    if labels_local.dtype not in ops._EXACT_IN_F32:
        raise RuntimeError("arvae_b200.distributed: labels must be exactly representable in float32 "
                           "(rank conversion of int64/float64 labels needs the global batch)")
    return _ShardedRegLossFn.apply(z_local, labels_local, dims, lcols, ops._scalar(gamma), ops._scalar(factor),
                                   group, int(algo))
