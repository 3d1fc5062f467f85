"""Data-parallel ELBO+gradient over the GPUs of one box: the minibatch's points are split contiguously over
the ranks (each rank keeps all S samples of its points; Philox counters use the global point index so the draws do not
depend on the placement), parameters — hence Kuu, its Cholesky and the KL term — are replicated, and ONE sum-allreduce of
the flat [data term, KL, gradients] buffer per step restores the full-batch result (KL pre-weighted by 1/world)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(N, rank, world):
    """Contiguous split of N points: rank r owns [lo, hi)."""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedELBO:
    def __init__(self, model, group=None, library_comm=True):
        """library_comm: the allreduce runs inside libdgp_b200 (dgp_comm_init / dgp_allreduce_grads: the library's own NCCL
        communicator on the ctx's stream) -- torch.distributed only carries the 128-byte ncclUniqueId to the ranks.
        False, or a CPU model (gloo tests of the host logic): torch.distributed.all_reduce."""
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._pinned_out = None
        self._ctx = None
        if library_comm and self.world > 1 and torch.cuda.is_available() and model.device.type == "cuda":
            from . import _lib
            import ctypes as C
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                rc = _lib.lib.dgp_comm_unique_id(C.c_void_p(uid.data_ptr()))
                if rc != 0:
                    raise _lib.DGPError(f"dgp_comm_unique_id failed ({rc}): libnccl.so.2 could not be loaded")
            uid_dev = uid.to(model.device)
            dist.broadcast(uid_dev, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = uid_dev.cpu()
            self._ctx = _lib.get_context(model.device)
            self._ctx.call("dgp_comm_init", self.rank, self.world, C.c_void_p(uid.data_ptr()))

    def step(self, X_local, Y_local, n_offset, scale=1.0, want_grad=True, seed=None, zs=None):
        """Device-resident shard in, reduced flat device buffer out."""
        flat = self.model.elbo_flat((X_local, Y_local), want_grad=want_grad, scale=scale, kl_weight=1.0 / self.world,
                                    seed=seed, n_offset=n_offset, zs=zs)
        if self.world > 1:
            if self._ctx is not None:
                from . import _lib
                self._ctx.call("dgp_allreduce_grads", _lib.ptr(flat), flat.numel())
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def train_adam_step(self, X_local, Y_local, n_offset, params, state, t, lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                        scale=1.0, seed=None):
        """One data-parallel training iteration: sharded ELBO gradient, the allreduce, then the same dgp_adam_step on every
        rank (the reduced buffer is bit-identical across ranks, so the replicated parameters stay in lockstep without a
        broadcast). `params` / `state` as in DGP_Base._adam_state; t counts from 1. Returns the reduced flat buffer."""
        flat = self.step(X_local, Y_local, n_offset, scale=scale, want_grad=True, seed=seed)
        self.model._adam_step(params, flat, state, t, lr, beta_1, beta_2, epsilon)
        return flat

    def step_host(self, X_host: torch.Tensor, Y_host: torch.Tensor, n_offset, scale=1.0, want_grad=True, seed=None):
        """Pinned HOST shard in, HOST result out: H2D copies, the step, the allreduce and the D2H copy of the result."""
        dev = self.model.device
        if self.world == 1:
            if self._pinned_out is None:
                n, _ = self.model.grad_layout()
                self._pinned_out = torch.empty(n, dtype=torch.float64).pin_memory()
            out = self._pinned_out.numpy()
            self.model.elbo_flat_host(X_host.numpy(), Y_host.numpy(), want_grad=want_grad, scale=scale, kl_weight=1.0,
                                      seed=seed, n_offset=n_offset, out_host=out)
            return self._pinned_out
        X = X_host.to(dev, non_blocking=True)
        Y = Y_host.to(dev, non_blocking=True)
        flat = self.step(X, Y, n_offset, scale, want_grad, seed)
        if self._pinned_out is None or self._pinned_out.numel() != flat.numel():
            self._pinned_out = torch.empty(flat.numel(), dtype=torch.float64).pin_memory()
        self._pinned_out.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return self._pinned_out
