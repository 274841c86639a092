"""Device timeline of graph-replayed updates (CUPTI through torch.profiler; works under torchrun, unlike ncu).

    python tools/timeline.py [--config C2] [--steps 6] [--peer-optimizer] [--out gpurun_out/timeline.json]
    torchrun ... tools/timeline.py --out gpurun_out/timeline_n8.json        (every rank profiles, rank 0 writes)

For the LAST profiled update it lists every kernel with its start (us after the update's first kernel), duration and stream,
and summarises: step span, summed kernel time per name, time during which only NCCL kernels were running (exposed
communication).  Numbers taken under the profiler are diagnostics, never bench values."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--cap", type=int, default=0, help="override the replay capacity (faster set-up)")
    ap.add_argument("--peer-optimizer", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    o = ap.parse_args()
    cfg = dict(bench.CONFIGS[o.config])
    if o.cap:
        cfg["cap"] = o.cap
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from rainbow_b200.dist import init_from_env
        init_from_env("nccl")
    from rainbow_b200.agent import Agent
    from rainbow_b200.dist import shard_seed
    from rainbow_b200.memory import ReplayMemory
    import numpy as np
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(shard_seed(0, rank))
    args = bench.make_args(cfg, dev, peer_optimizer=o.peer_optimizer)
    cap = cfg["cap"]
    mem = ReplayMemory(args, cap, seed=shard_seed(17, rank))
    meta = bench.synthetic_meta(cap, 1 + rank)
    tr = mem.transitions
    tr.load_arrays(timestep=meta["timestep"], action=meta["action"], reward=meta["reward"], nonterminal=meta["nonterminal"],
                   index=meta["head"], full=True, t_episode=int(meta["timestep"][meta["head"] - 1]) + 1)
    tr.frames.copy_(torch.randint(0, 256, (1024, 7056), dtype=torch.uint8, device=dev).repeat((cap + 1023) // 1024, 1)[:cap])
    pri = torch.from_numpy(meta["priority"]).to(dev)
    leaf = torch.arange(cap, device=dev) + tr.tree_start
    for s in range(0, cap, 1024):
        tr.update(leaf[s:s + 1024], pri[s:s + 1024])
    agent = Agent(args, bench.FakeEnv())
    for _ in range(Agent.GRAPH_WARMUP + 6):
        agent.reset_noise()
        agent.learn(mem)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(o.steps):
            agent.reset_noise()
            agent.learn(mem)
        torch.cuda.synchronize(dev)
    class Ev:   # kineto activity record -> (name, start us, end us, stream)
        def __init__(self, name, start, end, stream):
            self.name, self.start, self.end, self.stream = name, start, end, stream

    evs = []
    try:
        for e in prof.profiler.kineto_results.events():
            if e.device_type() == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name().lower() and "memset" not in e.name().lower():
                evs.append(Ev(e.name(), e.start_ns() / 1e3, e.end_ns() / 1e3, int(e.device_resource_id())))
    except Exception:   # older / newer profiler object layouts: fall back to the public event list (no stream ids)
        evs = [Ev(e.name, e.time_range.start, e.time_range.end, -1) for e in prof.events()
               if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()]
    evs.sort(key=lambda e: e.start)
    # split into updates: the online reset_noise (k_noise_factors) is the first kernel of an update, two launches per update
    starts = [i for i, e in enumerate(evs) if "k_noise_factors" in e.name]
    firsts = starts[0::2]
    last = evs[firsts[-1]:]
    t0 = last[0].start
    rows = [dict(name=e.name[:90], start_us=round(e.start - t0, 2), dur_us=round(e.end - e.start, 2), stream=e.stream) for e in last]
    span = max(r["start_us"] + r["dur_us"] for r in rows)
    per_step = (evs[firsts[-1]].start - evs[firsts[1]].start) / (len(firsts) - 2) if len(firsts) > 2 else None
    # time covered only by NCCL kernels
    edges = sorted({r["start_us"] for r in rows} | {r["start_us"] + r["dur_us"] for r in rows})
    exposed = 0.0
    for a, b in zip(edges[:-1], edges[1:]):
        active = [r for r in rows if r["start_us"] <= a and r["start_us"] + r["dur_us"] >= b]
        if active and all("nccl" in r["name"].lower() for r in active):
            exposed += b - a
    idle = sum(b - a for a, b in zip(edges[:-1], edges[1:]) if not any(r["start_us"] <= a and r["start_us"] + r["dur_us"] >= b for r in rows))
    by = {}
    for r in rows:
        k = r["name"][:60]
        by.setdefault(k, [0, 0.0])
        by[k][0] += 1
        by[k][1] += r["dur_us"]
    out = dict(config=o.config, world=world, rank=rank, peer_optimizer=o.peer_optimizer, step_span_us=round(span, 1),
               us_between_update_starts=None if per_step is None else round(per_step, 1), nccl_only_us=round(exposed, 1),
               idle_us=round(idle, 1), kernels=len(rows), by_kernel={k: [v[0], round(v[1], 1)] for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])},
               timeline=rows)
    if rank == 0:
        os.makedirs(os.path.dirname(o.out), exist_ok=True)
        with open(o.out, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps({k: v for k, v in out.items() if k not in ("timeline", "by_kernel")}))
        for r in rows:
            print(f"{r['start_us']:8.1f} +{r['dur_us']:7.1f}  s{r['stream']:<3} {r['name'][:80]}")
    sys.stdout.flush()
    if world > 1:
        torch.distributed.barrier()
    os._exit(0)     # CUPTI + NCCL teardown has been seen to hang: everything is written, leave without it


if __name__ == "__main__":
    main()
