"""GPU tests of the actor-side batched append (ReplayMemory(defer_appends=True) -> rb_append_batch)."""

import numpy as np
import pytest
import torch

from helpers import assert_bits_equal, golden
from test_gpu_parity import DEV, cpu, make_args

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flush_every", [1, 3, 8, 100])
@pytest.mark.parametrize("pinned_host", [False, True])
def test_deferred_appends_equal_immediate_appends(flush_every, pinned_host):
    """Same transitions through append() with and without deferral: identical tree, ring, metadata, running max."""
    from rainbow_b200.memory import ReplayMemory
    g = golden("append")
    a = ReplayMemory(make_args(), 8)
    b = ReplayMemory(make_args(), 8, defer_appends=True)
    for i in range(19):
        act, rew, term = g[f"a{i}_args"]
        frame = torch.from_numpy(g["last_frames_f32"][i])
        st_a = torch.zeros(4, 84, 84, device=DEV)
        st_a[-1] = frame.to(DEV)
        st_b = torch.zeros(4, 84, 84).pin_memory() if pinned_host else torch.zeros(4, 84, 84, device=DEV)
        st_b[-1] = frame
        a.append(st_a, int(act), float(rew), bool(term))
        b.append(st_b, int(act), float(rew), bool(term))
        if (i + 1) % flush_every == 0:
            b.flush_appends()
        if i == 9:
            for m in (a, b):   # update_priorities flushes first
                m.update_priorities(np.array([m.transitions.tree_start + 2]), np.array([9.0], np.float32))
        assert (a.transitions.index, a.transitions.full, a.t) == (b.transitions.index, b.transitions.full, b.t)
    b.flush_appends()
    torch.cuda.synchronize()
    assert_bits_equal(b.transitions.sum_tree, a.transitions.sum_tree, "tree")
    assert_bits_equal(b.transitions.sum_tree, g["a18_tree"], "tree vs reference")
    for k in ("state", "timestep", "action", "reward", "nonterminal"):
        assert np.array_equal(b.transitions.data[k], a.transitions.data[k]), k
    assert list(cpu(b.transitions.ring_state)[:3]) == list(cpu(a.transitions.ring_state)[:3])
    assert b.transitions.max == a.transitions.max


def test_deferred_appends_flush_before_sampling():
    from rainbow_b200.memory import ReplayMemory
    mem = ReplayMemory(make_args(), 64, defer_appends=True)
    rs = np.random.RandomState(0)
    for i in range(40):
        st = torch.from_numpy(rs.uniform(0, 1, (4, 84, 84)).astype(np.float32)).to(DEV)
        mem.append(st, i % 6, 0.0, i % 17 == 16)
    assert len(mem._queue) == 0 or len(mem._queue) < 8
    out = mem.sample(4)           # flushes the remainder first
    mem.check_last_sample()
    assert len(mem._queue) == 0 and int(mem.transitions.ring_state[3].item()) == 40
    assert out[1].shape == (4, 4, 84, 84)
