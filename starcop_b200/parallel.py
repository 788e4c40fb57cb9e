"""Data parallelism for the HyperSTARCOP train step: one process per GPU, tiles sharded across
ranks, ONE collective per step -- a sum all-reduce of the flat fp32 gradient arena (6.63 M
elements, 26.5 MB) over NCCL / NVLink.  BatchNorm statistics stay per GPU, like the reference
(scripts/train.py:120-138 passes neither ``strategy`` nor ``sync_batchnorm``).

The reference has no explicit collective: with ``devices > 1`` PyTorch-Lightning would wrap the
module in DistributedDataParallel (bucketed all-reduce of the same gradients, mean over ranks).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) -> (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


class GradSync:
    """grad_sync hook of ``ModelModule.train_step_fused``: sums the flat gradient arena across ranks
    and returns the scale (1/world) that the fused Adam applies, i.e. DDP's gradient mean.

    Two buckets, decoder first (what DDP's reverse-order buckets do): the arena is laid out in registration order
    (encoder | decoder | head), the backward pass produces it back to front.  ``early(flat_grads, split, main, side)``
    is called by the engine as soon as the head / decoder gradients -- elements ``[split:]``, 66 % of the bytes --
    have been issued; their all-reduce is launched from a communication stream and runs under the encoder's
    backward.  ``__call__`` then reduces what is left (``[:split]``) and joins.  ``policy`` for BatchNorm buffers is
    the reference's: none (running statistics stay per rank, ``scripts/train.py`` sets no ``sync_batchnorm``; DDP's
    ``broadcast_buffers`` re-broadcast of rank 0's buffers at each forward is NOT reproduced: evaluation uses each
    rank's own statistics, and checkpoints are written by rank 0)."""

    def __init__(self, world=None, group=None, bucketed=None):
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.group = group
        self.bucketed = (os.environ.get("STARCOP_NO_GRAD_BUCKETS", "") == "") if bucketed is None else bucketed
        self._split = None         # elements [split:] were reduced early in this step
        self._comm = None

    def early(self, flat_grads, split, main=None, side=None):
        if self.world <= 1 or not self.bucketed or split <= 0 or split >= flat_grads.numel():
            return
        tail = flat_grads[split:]
        if flat_grads.is_cuda:
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=flat_grads.device)
            main = main if main is not None else torch.cuda.current_stream(flat_grads.device)
            self._comm.wait_stream(main)
            if side is not None:
                self._comm.wait_stream(side)                  # the decoder's weight gradients run on the side stream
            with torch.cuda.stream(self._comm):
                dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=self.group)
        self._split = split

    def __call__(self, flat_grads):
        if self.world > 1:
            if self._split is not None:
                dist.all_reduce(flat_grads[:self._split], op=dist.ReduceOp.SUM, group=self.group)
                if flat_grads.is_cuda:
                    torch.cuda.current_stream(flat_grads.device).wait_stream(self._comm)
                self._split = None
            else:
                dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        return 1.0 / self.world


def broadcast_parameters(flat_params, buffers=(), src=0):
    """DDP's constructor broadcast: every rank starts from rank `src`'s weights and BN buffers."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src)
        for b in buffers:
            dist.broadcast(b, src)


def shard_tiles(n_tiles, rank, world):
    """Round-robin tile assignment: rank r takes tiles r, r+world, ..."""
    return list(range(rank, n_tiles, world))


def reduce_confusion(counts):
    """Validation: all-reduce of the int64 confusion-matrix counts (torchmetrics dist_reduce_fx="sum")."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
