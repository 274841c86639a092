"""Multi-GPU plumbing: one process per GPU, independent replay + transition stream per rank, and ONE
exchange step per update -- a SUM all-reduce of the flat float32 gradient; the 1/world
scaling is folded into the clip+Adam kernel.  The reference is single-process (no counterpart).

Everything here is device agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, world_size, local_rank); a no-op single-process answer when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        # NVLS (in-switch reduction) is left off unless the caller asks for it: a 4-of-8-GPU launch on this pool
        # hung in communicator setup with it on, while the 27 MB gradient all-reduce is overlapped with the conv
        # backward anyway (export NCCL_NVLS_ENABLE=1 to turn it back on)
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_seed(base_seed, rank):
    """Per-rank seed for the replay / noise / synthetic-stream generators (ranks must not share streams)."""
    return (int(base_seed) * 1000003 + 7919 * int(rank) + 1) & (2 ** 63 - 1)


class GradSync:
    """Gradient exchange for data-parallel learners.  With torch.distributed uninitialised (or world 1)
    every method is a no-op, so the single-GPU path is exactly the reference's single learner."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world_size = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.exchange = True   # False: somebody else (the peer-memory optimiser) moves the gradients

    def all_reduce_(self, flat_grad):
        """In-place SUM over ranks of the flat gradient buffer (one collective per update)."""
        if self.enabled and self.exchange:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return flat_grad

    def broadcast_(self, flat_param, src=0):
        """Make every rank start from rank `src`'s parameters."""
        if self.enabled:
            dist.broadcast(flat_param, src=src, group=self.group)
        return flat_param

    def max_(self, t):
        """MAX over ranks (used for timing: a step is as slow as the slowest rank)."""
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t
