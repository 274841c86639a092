"""The UNMODIFIED reference as a bench arm -- TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's cpu legs import this).

`make -C oracle _ref` packs the reference's three hot-path modules (memory.py, agent.py, model.py), hash-checked and byte
for byte, into oracle/_ref/reference_modules.zip (git-ignored build output).  This module imports them from the archive
(zipimport) under the reference's own module names,
builds the BASELINE.md synthetic replay WITHOUT the reference's 242-second Python-list constructor (memory.py:19: the
SegmentTree object is assembled field by field, `data` as np.zeros(cap, Transition_dtype); every METHOD that runs
afterwards is the reference's own) and times `dqn.reset_noise(); dqn.learn(mem)` (main.py:150-151) on a device of choice.
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "reference_modules.zip")
_mods = None


def available():
    return os.path.isfile(REF_DIR)


def modules():
    """(memory, agent, model) of the reference, imported from oracle/_ref under their own names (agent.py does
    `from model import DQN`), then removed from sys.modules again so dropin/ is not shadowed for anybody else."""
    global _mods
    if _mods is None:
        saved = {k: sys.modules.pop(k, None) for k in ("memory", "agent", "model")}
        sys.path.insert(0, REF_DIR)
        try:
            m = tuple(importlib.import_module(k) for k in ("memory", "agent", "model"))
            assert all(os.path.abspath(x.__file__).startswith(REF_DIR) for x in m)
        finally:
            sys.path.remove(REF_DIR)
            for k, v in saved.items():
                sys.modules.pop(k, None)
                if v is not None:
                    sys.modules[k] = v
        _mods = m
    return _mods


def make_replay(args, cap, meta, frame_seed=1):
    """Reference ReplayMemory holding the synthetic fill -- fields set like memory.py:13-20,93-102."""
    memory = modules()[0]
    mem = memory.ReplayMemory.__new__(memory.ReplayMemory)
    mem.device, mem.capacity, mem.history, mem.discount, mem.n = args.device, cap, args.history_length, args.discount, args.multi_step
    mem.priority_weight, mem.priority_exponent, mem.t = args.priority_weight, args.priority_exponent, 0
    mem.n_step_scaling = torch.tensor([mem.discount ** i for i in range(mem.n)], dtype=torch.float32, device=mem.device)
    t = memory.SegmentTree.__new__(memory.SegmentTree)
    t.index, t.size, t.full = 0, cap, False
    t.tree_start = 2 ** (cap - 1).bit_length() - 1
    t.sum_tree = np.zeros((t.tree_start + cap,), dtype=np.float32)
    t.data = np.zeros(cap, dtype=memory.Transition_dtype)
    t.max = 1
    mem.transitions = t
    t.data["timestep"], t.data["action"], t.data["reward"] = meta["timestep"], meta["action"], meta["reward"]
    t.data["nonterminal"] = meta["nonterminal"].astype(np.bool_)
    pat = np.random.RandomState(frame_seed).randint(0, 256, (1024, 84, 84), dtype=np.uint8)   # touches every page of the ring
    for s in range(0, cap, 1024):
        e = min(cap, s + 1024)
        t.data["state"][s:e] = pat[:e - s]
    for s in range(0, cap, 4096):
        e = min(cap, s + 4096)
        t.update(np.arange(s, e) + t.tree_start, meta["priority"][s:e])    # the reference's own SegmentTree.update
    t.index, t.full = meta["head"], True
    return mem


class Session:
    """One synthetic reference replay (7 GB of host memory at 1M transitions), timed on one or more devices."""

    def __init__(self, args, actions, cap, meta, log=lambda *a: None):
        self.args, self.actions, self.log = args, actions, log
        self.mem = make_replay(args, cap, meta)
        log("reference arm: replay filled")

    def time(self, device, updates, warmup, with_appends, replay_frequency=4, budget_s=None):
        """`updates` (at least 3; fewer than asked if budget_s runs out) unmodified `dqn.reset_noise(); dqn.learn(mem)` pairs
        with args.device = `device` (the replay stays on the host either way, memory.py:106,137-145).
        Returns (updates/s, seconds, updates done)."""
        _, agent, _ = modules()
        actions, mem = self.actions, self.mem
        dev = torch.device(device)
        args = argparse.Namespace(**{**vars(self.args), "device": dev})
        mem.device = dev
        mem.n_step_scaling = mem.n_step_scaling.to(dev)

        class Env:
            def action_space(self):
                return actions

        dqn = agent.Agent(args, Env())
        frames = [torch.rand(4, 84, 84, device=dev) for _ in range(8)]

        def step(i):
            if with_appends:
                for j in range(replay_frequency):
                    mem.append(frames[(i + j) % 8], (i + j) % actions, float((i % 3) - 1), (i * 4 + j) % 1000 == 999)
            dqn.reset_noise()
            dqn.learn(mem)

        def sync():
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)

        for i in range(warmup):
            step(i)
        sync()
        t0 = time.perf_counter()
        done = 0
        while done < updates:
            step(warmup + done)
            done += 1
            if budget_s is not None and done >= 3 and time.perf_counter() - t0 > budget_s:
                break
        sync()
        dt = time.perf_counter() - t0
        return done / dt, dt, done
