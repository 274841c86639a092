"""Probe (GPU): true per-stage GPU time of one learner update at bench config C2, each stage captured in its own
CUDA graph (N repetitions) and replayed at full clocks with warm caches.  ncu's per-launch durations on this pool are
taken at idle clocks with flushed caches, so they only give shares; this gives microseconds."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rainbow_b200.agent import Agent, c51_dueling_loss_grad  # noqa: E402
from rainbow_b200.memory import ReplayMemory, _SampleWorkspace  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
dev = torch.device("cuda:0")
args = bench.make_args(cfg, dev)
cap = min(cfg["cap"], 200_000)  # tree depth differs slightly from 1M; frames random either way
mem = ReplayMemory(args, cap)
meta = bench.synthetic_meta(cap, 1)
tr = mem.transitions
tr.load_arrays(timestep=meta["timestep"], action=meta["action"], reward=meta["reward"], nonterminal=meta["nonterminal"],
               index=meta["head"], full=True)
tr.frames.random_(0, 256)
leaf = torch.arange(cap, device=dev) + tr.tree_start
pri = torch.from_numpy(meta["priority"]).to(dev)
for s in range(0, cap, 1024):
    tr.update(leaf[s:s + 1024], pri[s:s + 1024])
ag = Agent(args, bench.FakeEnv())
B = cfg["B"]
ws = _SampleWorkspace(B, 4, dev)
on, tg = ag.online_net, ag.target_net


def timeit(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 1e3 / (5 * reps):8.1f} us", flush=True)


mem.sample_into(ws)
states, nstates = ws.states, ws.next_states
timeit("sample (K1)", lambda: mem._launch_sample(ws))
timeit("gather (K2)", lambda: mem._launch_gather(ws))
timeit("noise factors (online)", lambda: on.reset_noise())
with torch.no_grad():
    timeit("conv fwd x1 (B rows, no grad)", lambda: on.features(states))
    x_s = on.features(states)
    x_ns = on.features(nstates)
    timeit("head fwd online (2B rows)", lambda: on.head().forward(x_s, x_ns))
    timeit("head fwd target (B rows)", lambda: tg.head().forward(x_ns))
    from rainbow_b200 import _lib
    L = _lib.load()
    L.rb_head_debug(2)
    timeit("  fc1 only, online (2B rows)", lambda: on.head().forward(x_s, x_ns))
    timeit("  fc1 only, target (B rows)", lambda: tg.head().forward(x_ns))
    L.rb_head_debug(1)
    timeit("  fc2 only, online (2B rows)", lambda: on.head().forward(x_s, x_ns))
    timeit("  fc2 only, target (B rows)", lambda: tg.head().forward(x_ns))
    L.rb_head_debug(0)
    z_on, h_on, p_on = on.head().forward(x_s, x_ns)
    z_t, _, _ = tg.head().forward(x_ns)
    z_t = z_t.clone()
    run_c51 = lambda: c51_dueling_loss_grad(z_on, z_t, ag.action_space, ag.atoms, ws.actions, ws.returns, ws.nonterminals,
                                            ws.weights, ag.support, ag.Vmin, ag.Vmax, ag.delta_z, 0.99 ** 3)
    timeit("c51 dueling (K3)", run_c51)
    loss, dz = run_c51()
    dh = torch.empty(B, 2 * on.hidden_size, device=dev)
    dx = torch.empty_like(x_s)
    timeit("head backward (3 kernels)", lambda: on.head().backward(p_on, x_s, h_on[:B], dz, dh, dx))
    timeit("  wgrad2 only", lambda: on.head().backward(p_on, x_s, h_on[:B], dz, dh, dx, parts=1))
    timeit("  dh only", lambda: on.head().backward(p_on, x_s, h_on[:B], dz, dh, dx, parts=2))
    timeit("  bwd1 only", lambda: on.head().backward(p_on, x_s, h_on[:B], dz, dh, dx, parts=4))
    timeit("zero conv grads", lambda: ag.optimiser.zero_conv_grad())


def conv_fb():
    xs = on.features(states)
    xs.backward(dx)


timeit("conv fwd+bwd (B rows, autograd)", conv_fb, reps=10)
timeit("clip+adam (K7)", lambda: ag.optimiser.step())
timeit("tree update (K4)", lambda: mem.update_priorities(ws.tree_idx, loss))
st = torch.rand(4, 84, 84, device=dev)
timeit("append (K5)", lambda: mem.append(st, 1, 0.0, False))
