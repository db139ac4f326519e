"""Row-block sharding of the pairwise loss across the GPUs of one NVSwitch box.

One process per GPU (``torch.distributed`` for the plumbing).  Rank g holds the samples its own encoder produced --
``z_local`` [B/G, Z] and ``labels_local`` [B/G, A].  Two transports:

``comm=ShardComm(...)``  (the fast path, csrc/reg_shard.cuh)
    The exchange is done by the kernels themselves over NVLink peer memory, no NCCL call in the step:

    1. every rank argsorts ITS OWN rows per dim and stores the sorted run (key + latent, 12 bytes per element)
       straight into every peer's communication buffer -- the column all-gather;
    2. every rank merges the G runs into the single-GPU sorted order and sweeps its 1/G share of that order's row
       blocks (the plan, tile classes and cost model are the single-GPU ones);
    3. every rank pulls the row sums of its own samples, and all G loss partials, from its peers -- gradient return
       and all-reduce.

    Row sums are fixed-point integers, so the loss and every gradient element are BITWISE those of the single-GPU op
    on the concatenated batch, for any G.

``comm=None``  (NCCL; SURVEY section 8e's plain form)
    all-gather of the packed local slice [B/G, 2R], the pair kernel over this rank's OWN rows x all B columns, and an
    all-reduce(sum) of one float64 loss partial.  No gradient exchange: by antisymmetry the full
    gradient of row i is a row sum over all columns.

Either way the gathered remote columns are constants in autograd, and the returned loss / gradient are those of the
GLOBAL-batch loss.  **Under DistributedDataParallel** the parameter gradients are averaged over ranks afterwards: the
reconstruction term is a local-batch mean, for which that average is the global-batch gradient, but a global-batch
loss would come out 1/G too small.  Pass ``ddp_average=True`` to have the backward multiply by the world size, so
that after DDP's averaging the regularisation gradient is exactly the single-GPU global-batch one (the returned loss
VALUE is unchanged).

The reference has no distributed code; this follows BASELINE.json's north_star and SURVEY section 8(e).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops


# --------------------------------------------------------------------------------------------
# NVLink peer-memory transport
# --------------------------------------------------------------------------------------------
class _ShardHandle:
    """One rank's state of the sharded step (C side: ShardCtx)."""

    def __init__(self, rank: int, world: int, n_cap: int, R_cap: int, device):
        self.rank, self.world, self.n_cap, self.R_cap = int(rank), int(world), int(n_cap), int(R_cap)
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.arvae_shard_create(self.rank, self.world, self.n_cap, self.R_cap, ctypes.byref(self.ctx)),
                       "arvae_shard_create")

    def ipc_handle(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.arvae_shard_ipc_handle(self.ctx, buf), "arvae_shard_ipc_handle")
        return buf.raw

    def comm_ptr(self) -> int:
        return int(self.lib.arvae_shard_comm_ptr(self.ctx))

    def step(self, z_local, labels_local, reg_dims, label_cols, n_all, gamma, factor, want_grad=True, phases=0):
        """Enqueue (phases of) one sharded step; returns (loss64 [()], loss32 [()], grad_cols [n_local, R] | None)."""
        dev = self.device
        R = len(reg_dims)
        n_local = int(n_all[self.rank])
        with torch.cuda.device(dev):
            if phases in (0, 1) or getattr(self, "_out", None) is None:  # later phases of a step reuse its outputs
                self._out = (torch.empty((), dtype=torch.float64, device=dev),
                             torch.empty((), dtype=torch.float32, device=dev),
                             torch.empty((n_local, R), dtype=torch.float32, device=dev) if want_grad else None)
            loss64, loss32, grad_cols = self._out
            rc = self.lib.arvae_shard_reg_loss_f32(
                self.ctx, ops._ptr(z_local), z_local.stride(0), z_local.stride(1), ops._ptr(labels_local),
                labels_local.stride(0), labels_local.stride(1), _lib.i32_array(reg_dims), _lib.i32_array(label_cols), R,
                _lib.i64_array(n_all), float(gamma), float(factor), ops._ptr(loss64), ops._ptr(loss32),
                ops._ptr(grad_cols), int(phases), ops._stream(dev))
            _lib.check(rc, "arvae_shard_reg_loss_f32")
        return loss64, loss32, grad_cols

    def status(self) -> Tuple[int, int]:
        st, ep = ctypes.c_int32(), ctypes.c_uint64()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.arvae_shard_status(self.ctx, ctypes.byref(st), ctypes.byref(ep), ops._stream(self.device)),
                       "arvae_shard_status")
        return st.value, ep.value

    def close(self):
        if self.ctx:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                self.lib.arvae_shard_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()


class ShardComm:
    """NVLink peer-memory communicator of the sharded step for one process per GPU.

    ``n_cap`` = the most rows any rank will ever pass, ``R_cap`` the most regularised dims.  Construction is
    collective over ``group`` (one all-gather of 64-byte CUDA IPC handles); afterwards a step never calls NCCL."""

    def __init__(self, n_cap: int, R_cap: int, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("arvae_b200.distributed.ShardComm: torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("ShardComm: at most 16 ranks (one NVSwitch box)")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.h = _ShardHandle(self.rank, self.world, n_cap, R_cap, device)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, self.h.ipc_handle(), group=group)
        with torch.cuda.device(device):
            _lib.check(self.h.lib.arvae_shard_open_peers(self.h.ctx, b"".join(handles)), "arvae_shard_open_peers")
        dist.barrier(group=group)  # every rank has mapped every buffer before the first step stores into them

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody unmaps while a peer may still pull from it
        self.h.close()


class LocalShardGroup:
    """G virtual ranks of the sharded step inside ONE process on ONE device (tests, and timing one rank's share of a
    G-GPU step on a single GPU).  The ranks' communication buffers are plain device memory of the same process; since
    a rank's wait could never be satisfied by a kernel queued behind it on the same stream, the step is driven
    phase by phase: sort + publish for every rank, then ranking of the own runs for every rank, then apply + plan +
    pairs for every rank, then finalize for every rank."""

    def __init__(self, world: int, n_cap: int, R_cap: int, device="cuda"):
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.world = world
        self.ranks = [_ShardHandle(g, world, n_cap, R_cap, device) for g in range(world)]
        for a in self.ranks:
            for b in self.ranks:
                if a is not b:
                    _lib.check(a.lib.arvae_shard_set_peer(a.ctx, b.rank, ctypes.c_void_p(b.comm_ptr())), "arvae_shard_set_peer")

    def step(self, z_parts, label_parts, reg_dims, label_cols, gamma, factor, want_grad=True, only_rank=None):
        """z_parts[g] / label_parts[g]: rank g's rows.  Returns [(loss64, loss32, grad_cols)] per rank."""
        n_all = [int(z.shape[0]) for z in z_parts]
        outs = [None] * self.world
        for phase in (1, 2, 4, 8):
            for g, h in enumerate(self.ranks):
                outs[g] = h.step(z_parts[g], label_parts[g], reg_dims, label_cols, n_all, gamma, factor, want_grad, phase)
        return outs

    def close(self):
        for h in self.ranks:
            h.close()


# --------------------------------------------------------------------------------------------
# NCCL transport (own rows x all columns)
# --------------------------------------------------------------------------------------------
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
    def forward(ctx, z_local, labels_local, reg_dims, label_cols, gamma, factor, group, algo, comm, grad_scale):
        R = len(reg_dims)
        want_grad = bool(ctx.needs_input_grad[0])
        if comm is not None:
            n_all = [int(z_local.shape[0])] * comm.world
            _, loss32, grad_cols = comm.h.step(z_local.detach(), labels_local.detach(), reg_dims, label_cols, n_all,
                                               gamma, factor, want_grad)
        else:
            packed_local = pack_columns(z_local, labels_local, reg_dims, label_cols)
            packed, row0 = gather_columns(packed_local, group)
            loss64, grad_cols = _rows_backend(packed, R, gamma, factor, row0, row0 + z_local.shape[0], want_grad, algo)
            dist.all_reduce(loss64, op=dist.ReduceOp.SUM, group=group)
            loss32 = loss64.to(torch.float32)
        ctx.reg_dims = tuple(reg_dims)
        ctx.shape = tuple(z_local.shape)
        ctx.grad_scale = float(grad_scale)
        if want_grad:
            ctx.save_for_backward(grad_cols)
        return loss32  # the same rounding of the float64 total, done by the finalize kernel itself

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (grad_cols,) = ctx.saved_tensors
        n, Z = ctx.shape
        if ctx.grad_scale != 1.0:
            grad_out = grad_out * ctx.grad_scale
        grad_z = _scatter(grad_cols, grad_out, ctx.reg_dims, n, Z)
        return (grad_z,) + (None,) * 9


def _scatter(grad_cols, grad_out, reg_dims, n, Z):
    return ops._scatter_bwd(grad_cols, grad_out, reg_dims, n, Z)


def reg_loss_sharded(z_local: torch.Tensor, labels_local: torch.Tensor, reg_dims: Sequence[int], gamma,
                     factor=1.0, label_cols: Optional[Sequence[int]] = None, group=None,
                     algo: int = ops.ALGO_AUTO, comm: Optional[ShardComm] = None,
                     ddp_average: bool = False) -> torch.Tensor:
    """Global-batch attribute-regularization loss from per-rank row blocks.

    Every rank passes its own ``[B/G, Z]`` latents and ``[B/G, A]`` labels (equal sizes on all
    ranks) and gets the same scalar: the loss the reference would compute on the concatenated
    batch.  ``z_local.grad`` receives d(global loss)/d(z_local) -- times the world size when
    ``ddp_average`` is set, which makes DistributedDataParallel's later 1/G averaging of the parameter
    gradients reproduce the single-GPU global-batch gradient (see the module docstring).

    ``comm``: a :class:`ShardComm` selects the NVLink peer-memory transport (no NCCL in the step, results
    bitwise equal to one GPU); ``None`` uses NCCL all-gather / all-reduce around the row-block kernel.
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
    if comm is not None:
        ops._require_cuda_f32(z_local, "z_local")
        if labels_local.dtype != torch.float32:
            labels_local = labels_local.to(torch.float32)
        if z_local.shape[0] > comm.h.n_cap or len(dims) > comm.h.R_cap:
            raise RuntimeError(f"arvae_b200.distributed: ShardComm sized for {comm.h.n_cap} rows x {comm.h.R_cap} dims, "
                               f"got {z_local.shape[0]} x {len(dims)}")
    world = dist.get_world_size(group)
    return _ShardedRegLossFn.apply(z_local, labels_local, dims, lcols, ops._scalar(gamma), ops._scalar(factor),
                                   group, int(algo), comm, float(world) if ddp_average else 1.0)
