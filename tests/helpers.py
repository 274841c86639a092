"""Shared helpers for the parity tests: golden loading and oracle-tree construction."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLD, name + ".npz")))
    return _cache[name]


def manifest():
    with open(os.path.join(GOLD, "MANIFEST.json")) as f:
        return json.load(f)


def oracle_ring_from_golden(g, pfx):
    """OracleTree loaded with the ring + tree the reference held (dump_ring in oracle/gen_golden.py)."""
    import oracle
    meta = g[pfx + "meta"]
    t = oracle.OracleTree(int(meta[3]))
    t.sum_tree[:] = g[pfx + "sum_tree"]
    t.frames[:] = g[pfx + "frames"]
    t.timestep[:] = g[pfx + "timestep"]
    t.action[:] = g[pfx + "action"]
    t.reward[:] = g[pfx + "reward"]
    t.nonterminal[:] = g[pfx + "nonterminal"]
    t.index, t.full = int(meta[0]), bool(meta[1])
    t.max[0] = g[pfx + "max"]
    return t


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    if a.dtype.kind == "f":
        bad = np.flatnonzero(bits(a).ravel() != bits(b).ravel())
    else:
        bad = np.flatnonzero(a.ravel() != b.ravel())
    assert bad.size == 0, f"{what}: {bad.size} mismatches, first at {bad[:5]}: {a.ravel()[bad[:5]]} vs {b.ravel()[bad[:5]]}"


def sample_stride(numel):
    """Sub-sampling rule of the full-update fixtures (same function as in oracle/gen_golden.py)."""
    return 1 if numel <= 4096 else (5 if numel <= 40000 else (23 if numel <= 200000 else 199))


def update_case(name):
    for c in manifest()["update_cases"]:
        if c["name"] == name:
            return c
    raise KeyError(name)


def update_case_ring(case):
    """Frames of a full-update fixture's ring, re-drawn from the seed exactly like oracle/gen_golden.py::gen_full_update
    does (the fixture stores only their SHA-256)."""
    import hashlib
    rs = np.random.RandomState(case["seed"] + 7)
    frames = rs.randint(0, 256, (case["fill"], 84, 84), dtype=np.uint8)
    ring = np.zeros((case["cap"], 84, 84), np.uint8)
    for i in range(case["fill"]):
        ring[i % case["cap"]] = frames[i]
    assert hashlib.sha256(np.ascontiguousarray(ring).tobytes()).hexdigest() == case["frames_sha"], "frame stream differs"
    return ring
