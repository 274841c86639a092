"""Validate the peer-memory optimiser (rb_peer_reduce + rb_peer_adam_gather) against NCCL all-reduce + rb_clip_adam (needs 2, 4 or 8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/peer_adam_check.py

Every rank feeds the same per-rank random gradients to both optimisers for a few steps and compares the parameters
(they must agree to float rounding: the reduction order differs only in where the 1/world scaling is applied)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rainbow_b200 import _lib  # noqa: E402
from rainbow_b200.dist import init_from_env  # noqa: E402
from rainbow_b200.peer import PeerOptimizerState  # noqa: E402

rank, world, local = init_from_env("nccl")
dev = torch.device("cuda", local)
P = 6_868_928
L = _lib.load()
CONV_END = 78_272                                     # the learner's two segments: noisy head first (reduced early, on a
peer = PeerOptimizerState(P, dev, segments=[(CONV_END, P), (0, CONV_END)])   # side stream), conv parameters second
Pp = peer.numel
side = torch.cuda.Stream(device=dev)
torch.manual_seed(0)
p0 = torch.randn(Pp, device=dev) * 0.05
peer.flat_param.copy_(p0)
ref_p, ref_m, ref_v = p0.clone(), torch.zeros(Pp, device=dev), torch.zeros(Pp, device=dev)
ref_step = torch.zeros(1, dtype=torch.int64, device=dev)
ref_part = torch.zeros(L.rb_clip_adam_scratch_elems(), dtype=torch.float64, device=dev)
ref_norm = torch.zeros(1, device=dev)
torch.cuda.synchronize()
dist.barrier()
for it in range(6):
    torch.manual_seed(1000 * it + rank)
    g = torch.randn(Pp, device=dev) * (30.0 if it % 2 else 0.01) / Pp ** 0.5   # with and without clipping
    peer.flat_grad.copy_(g)
    if it % 3 != 2:                                       # like the learner: head segment early on a side stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            peer.reduce_segment(0)
        torch.cuda.current_stream().wait_stream(side)
    peer.step(10.0, 6.25e-5, (0.9, 0.999), 1.5e-4)
    gr = g.clone()
    dist.all_reduce(gr)
    _lib.check(L.rb_clip_adam(ref_p.data_ptr(), gr.data_ptr(), ref_m.data_ptr(), ref_v.data_ptr(), Pp, 1.0 / world, 10.0, 6.25e-5,
                              0.9, 0.999, 1.5e-4, ref_step.data_ptr(), ref_part.data_ptr(), ref_norm.data_ptr(), None,
                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    dp = (peer.flat_param - ref_p).abs().max().item()
    dn = abs(peer.grad_norm.item() - ref_norm.item()) / max(ref_norm.item(), 1e-12)
    dm = max((peer.exp_avg[sh] - ref_m[fl]).abs().max().item() for fl, sh in peer.shard_slices())
    print(f"rank {rank} step {it}: max|dp| {dp:.3e}  rel dnorm {dn:.3e}  max|dm| {dm:.3e}  step {int(peer.step_count.item())}", flush=True)
    assert dp < 5e-7 and dn < 1e-5 and dm < 1e-7, "peer optimiser disagrees with the NCCL path"
    chk = peer.flat_param.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    assert torch.equal(chk, peer.flat_param), "ranks hold different parameters"
dist.barrier()
if rank == 0:
    print(f"peer optimiser OK (world {world}, segments {peer.segments}, multicast all-gather {peer.multicast})")
dist.destroy_process_group()
