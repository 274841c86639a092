"""Peer-memory optimiser state for multi-GPU learners (NVLink / NVSwitch, one process per GPU).

ROUND-1 STATUS: written and compiled (csrc/rb_peer.cu) but not exercised on hardware -- the round's GPU budget ran out
before a 2-GPU validation slot was left.  Nothing uses it unless `Agent(args, env)` is given `args.peer_optimizer = True`;
the measured multi-GPU path is the NCCL all-reduce of `rainbow_b200.dist.GradSync`.

Design (brief: "where a hot op is a compute step followed by a collective, write ONE kernel that does
both over peer memory"): the flat gradient and parameter buffers of every rank live in symmetric memory mapped on all
GPUs.  Per update `rb_peer_clip_adam` (1) reduce-scatters the gradient with peer loads, (2) clips by the global norm and
runs Adam on the owned 1/world slice only -- the Adam moments are sharded --, (3) all-gathers the updated parameters with
peer stores; cross-GPU ordering is epoch flags in the same symmetric allocation, never the host.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class PeerOptimizerState:
    """Symmetric allocation holding [flat_param | flat_grad | flags | norms] of one rank, rendezvoused over `group`."""

    def __init__(self, numel, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise _lib.RainbowB200Error("peer optimiser supports up to 8 ranks (one NVLink domain)")
        q = 4 * self.world
        self.numel = -(-numel // q) * q                      # P, multiple of 4 * world
        self.slice = self.numel // self.world
        pbytes = self.numel * 4
        fbytes = 3 * self.world * 8
        nbytes = self.world * 8
        self._off = (0, pbytes, 2 * pbytes, 2 * pbytes + 256 * (-(-fbytes // 256)))
        total = self._off[3] + 256 * (-(-nbytes // 256))
        grp = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, grp.group_name)
        self.flat_param = self.buf[self._off[0]:self._off[0] + pbytes].view(torch.float32)
        self.flat_grad = self.buf[self._off[1]:self._off[1] + pbytes].view(torch.float32)
        bases = [int(p) for p in self.handle.buffer_ptrs]
        n = self.world
        self._peer_param = (C.c_void_p * n)(*[b + self._off[0] for b in bases])
        self._peer_grad = (C.c_void_p * n)(*[b + self._off[1] for b in bases])
        self._peer_flags = (C.c_void_p * n)(*[b + self._off[2] for b in bases])
        self._peer_norms = (C.c_void_p * n)(*[b + self._off[3] for b in bases])
        f32 = torch.float32
        self.gred = torch.zeros(self.slice, dtype=f32, device=device)
        self.exp_avg = torch.zeros(self.slice, dtype=f32, device=device)      # this rank's shard of the moments
        self.exp_avg_sq = torch.zeros(self.slice, dtype=f32, device=device)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self.grad_norm = torch.zeros(1, dtype=f32, device=device)
        self._lib = _lib.load()
        self._scratch = torch.zeros(self._lib.rb_peer_scratch_bytes(), dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                   # every rank's zeroed flags are in place before the first step

    def step(self, max_norm, lr, betas, eps):
        _lib.check(self._lib.rb_peer_clip_adam(
            self._peer_grad, self._peer_param, self._peer_flags, self._peer_norms, self.world, self.rank, self.numel,
            _lib.ptr(self.gred), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), 1.0 / self.world, float(max_norm),
            float(lr), float(betas[0]), float(betas[1]), float(eps), _lib.ptr(self.step_count), _lib.ptr(self.epoch),
            _lib.ptr(self._scratch), _lib.ptr(self.grad_norm), _lib.stream()))
