"""CPU property tests of the oracle sum tree (hypothesis): structural invariants that hold for every even capacity
and every update stream, complementing the fixed golden vectors (tests/test_oracle_golden.py)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


@settings(max_examples=40, deadline=None)
@given(half=st.integers(1, 300), seed=st.integers(0, 2 ** 31 - 1), batches=st.integers(1, 6))
def test_tree_invariants(half, seed, batches):
    cap = 2 * half
    rs = np.random.RandomState(seed)
    t = oracle.OracleTree(cap, with_data=False)
    ts = t.tree_start
    leaves = np.zeros(cap, np.float32)
    for _ in range(batches):
        B = int(rs.randint(1, 2 * cap + 1))
        idx = rs.randint(0, cap, B).astype(np.int64)
        val = rs.uniform(0, 3, B).astype(np.float32)
        val[rs.uniform(size=B) < 0.1] = 0.0
        t.update(idx + ts, val)
        for i, v in zip(idx, val):      # last write wins
            leaves[i] = v
    tree = t.sum_tree
    assert np.array_equal(tree[ts:ts + cap], leaves)
    # every internal node whose two children exist is fl32(left + right); nodes above missing leaves stay 0
    n_int = (tree.size - 1) // 2
    par = np.arange(n_int)
    assert np.array_equal(tree[par], tree[2 * par + 1] + tree[2 * par + 2])
    assert float(t.max[0]) >= 1.0 and float(t.max[0]) >= float(leaves.max(initial=0.0)) - 0.0
    total = float(tree[0])
    if total > 0:
        # find(): the returned leaf's prefix interval contains the value (float32 partial sums vs float64 prefix:
        # tolerance 1e-4 of the total), indices are in range, probabilities are the leaf values
        vals = rs.uniform(0, total, 64)
        p, di, ti = t.find(vals)
        assert np.all(di >= 0) and np.all(di < cap) and np.array_equal(ti, di + ts)
        assert np.array_equal(p, leaves[di])
        prefix = np.concatenate([[0.0], np.cumsum(leaves.astype(np.float64))])
        tol = 1e-4 * total + 1e-6
        assert np.all(prefix[di] <= vals + tol) and np.all(vals <= prefix[di + 1] + tol)
        # values beyond the total clip to the last array element (memory.py:70-71)
        p2, di2, ti2 = t.find(np.array([total * 4 + 1.0]))
        assert ti2[0] == tree.size - 1 or leaves[di2[0]:].sum() == leaves[di2[0]]


@settings(max_examples=30, deadline=None)
@given(B=st.integers(1, 64), seed=st.integers(0, 2 ** 31 - 1))
def test_stratified_samples_and_weights(B, seed):
    rs = np.random.RandomState(seed)
    total = np.float32(rs.uniform(0.1, 1e5))
    u = rs.random_sample(B)
    v = oracle.segment_samples(total, B, u)
    seg = float(np.float32(total) / np.float32(B))
    assert np.all(np.diff(v) > -1e-12) and np.all(v >= np.arange(B) * seg) and np.all(v <= (np.arange(B) + 1) * seg)
    probs = rs.uniform(1e-3, 2, B).astype(np.float32)
    w = oracle.is_weights(probs, total, 1000, 0.4)
    assert w.max() == np.float32(1.0) and np.all(w > 0)
    assert np.argmax(w) == np.argmin(probs)        # the rarest sample carries the largest weight


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), A=st.integers(1, 8), B=st.integers(1, 16))
def test_c51_projection_conserves_mass(seed, A, B):
    rs = np.random.RandomState(seed)
    Z = 51
    q = [rs.normal(0, 3, (B, A, Z)).astype(np.float32) for _ in range(3)]
    actions = rs.randint(0, A, B).astype(np.int64)
    returns = rs.uniform(-15, 15, B).astype(np.float32)        # includes clamped targets
    nonterm = (rs.uniform(size=(B, 1)) > 0.3).astype(np.float32)
    w = rs.uniform(0.1, 1, B).astype(np.float32)
    support = np.linspace(-10, 10, Z).astype(np.float32)
    loss, grad, m, astar = oracle.c51(q[0], q[1], q[2], actions, returns, nonterm, w, support, -10.0, 10.0, 0.4, 0.99 ** 3)
    np.testing.assert_allclose(m.sum(1), 1.0, atol=2e-6)       # a probability distribution (agent.py:89-92)
    assert np.all(m >= 0) and np.all(loss > 0) and np.all((astar >= 0) & (astar < A))
    # the gradient lives on the taken action's row only and sums to ~0 over the atoms (softmax Jacobian)
    for i in range(B):
        mask = np.ones(A, bool)
        mask[actions[i]] = False
        assert not grad[i, mask].any()
        assert abs(grad[i, actions[i]].sum()) < 1e-6
    # terminal transitions put all mass on at most two adjacent atoms around the clamped return
    term = np.flatnonzero(nonterm[:, 0] == 0)
    for i in term:
        b = (np.clip(returns[i], -10, 10) + 10) / 0.4
        nz = np.flatnonzero(m[i])
        assert 1 <= nz.size <= 2 and nz.max() - nz.min() <= 1
        assert abs(nz.mean() - b) <= 1.0 + 1e-3


def _warp_walk(tree, ts, leaves, values):
    """numpy emulation of the register algorithm of k_tree_update_warp / k_append_batch: every lane prefetches the
    siblings of its path BEFORE anything is written, lanes that share a parent take the sibling value from the partner
    lane if the sibling is itself being updated, the lowest lane of a group writes."""
    L = int(np.log2(ts + 1))
    lanes = len(leaves)
    # duplicates: highest lane wins
    val = np.array(values, np.float32)
    for i in range(lanes):
        same = [j for j in range(lanes) if leaves[j] == leaves[i]]
        val[i] = np.float32(values[max(same)])
    sib = np.zeros((lanes, L), np.float32)
    for i in range(lanes):
        for l in range(L):
            nl = ((leaves[i] + 1) >> l) - 1
            sn = nl + 1 if nl & 1 else nl - 1
            sib[i, l] = tree[sn]
    node = np.array(leaves, np.int64)
    for i in range(lanes):
        tree[node[i]] = val[i]
    for l in range(L):
        parent = (node - 1) >> 1
        is_left = (node & 1) == 1
        new = val.copy()
        for i in range(lanes):
            other = [j for j in range(lanes) if parent[j] == parent[i] and is_left[j] != is_left[i]]
            sv = val[min(other)] if other else sib[i, l]
            new[i] = np.float32(val[i] + sv)
        val, node = new, parent
        for i in range(lanes):
            tree[node[i]] = val[i]
    return tree


@settings(max_examples=60, deadline=None)
@given(half=st.integers(2, 200), seed=st.integers(0, 2 ** 31 - 1), k=st.integers(1, 8), start=st.integers(0, 10 ** 6))
def test_batched_walk_equals_sequential_appends(half, seed, k, start):
    """The claim behind rb_append_batch: k sequential single-leaf walks (memory.py:36-41) leave the same float32 tree as
    ONE batched walk over the k (consecutive, wrapping) leaves -- and the same for arbitrary leaf sets with duplicates
    (rb_tree_update's single-warp path)."""
    cap = 2 * half
    k = min(k, cap)
    rs = np.random.RandomState(seed)
    t = oracle.OracleTree(cap, with_data=False)
    ts = t.tree_start
    t.update(np.arange(cap) + ts, rs.uniform(0, 3, cap).astype(np.float32))
    vmax = np.float32(t.max[0])
    head = start % cap
    seq = t.sum_tree.copy()
    for j in range(k):                                   # k sequential appends of the running max
        oracle.lib().orc_tree_set_leaf(seq, ts + (head + j) % cap, float(vmax))
    leaves = [ts + (head + j) % cap for j in range(k)]
    bat = _warp_walk(t.sum_tree.copy(), ts, leaves, [vmax] * k)
    assert np.array_equal(bat, seq)
    # arbitrary update batch with duplicates vs the oracle's level-synchronous update
    B = int(rs.randint(1, 33))
    idx = (rs.randint(0, cap, B) + ts).tolist()
    if B > 2:
        idx[-1] = idx[0]
    vals = rs.uniform(0, 4, B).astype(np.float32)
    ref = oracle.OracleTree(cap, with_data=False)
    ref.sum_tree[:] = t.sum_tree
    ref.update(np.array(idx, np.int64), vals)
    got = _warp_walk(t.sum_tree.copy(), ts, idx, vals.tolist())
    assert np.array_equal(got, ref.sum_tree)
