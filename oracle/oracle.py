"""ctypes front-end of oracle/rb_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE: a CPU restatement of reference memory.py / agent.py /
model.py arithmetic, pinned against vectors generated from the unmodified
reference (oracle/gen_golden.py -> tests/golden/).  See rb_oracle.c for the
reference file:line each function follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librb_oracle.so")
_lib = None

FRAME = 84 * 84

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile rb_oracle.c with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "rb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "librb_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_tree_start.restype = C.c_int64
        L.orc_tree_start.argtypes = [C.c_int64]
        L.orc_pow_priorities.argtypes = [_f32p, C.c_float, C.c_int, _f32p]
        L.orc_tree_update.argtypes = [_f32p, _i64p, _f32p, C.c_int, _f32p]
        L.orc_tree_set_leaf.argtypes = [_f32p, C.c_int64, C.c_float]
        L.orc_tree_find.argtypes = [_f32p, C.c_int64, C.c_int64, _f64p, C.c_int, _f32p, _i64p, _i64p]
        L.orc_segment_samples.argtypes = [C.c_float, C.c_int, _f64p, _f64p]
        L.orc_batch_valid.restype = C.c_int
        L.orc_batch_valid.argtypes = [_i64p, _f32p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.orc_is_weights.argtypes = [_f32p, C.c_float, C.c_int64, C.c_float, C.c_int, _f32p]
        L.orc_quantise_frame.argtypes = [_f32p, _u8p]
        L.orc_gather.argtypes = [_u8p, _i32p, _i32p, _f32p, _u8p, C.c_int64, _i64p, C.c_int, C.c_int, C.c_int,
                                 _f32p, _f32p, _f32p, _i64p, _f32p, _f32p]
        L.orc_iter_state.argtypes = [_u8p, _i32p, C.c_int64, C.c_int64, C.c_int, _f32p]
        L.orc_c51.argtypes = [_f32p, _f32p, _f32p, _i64p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float,
                              C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_void_p,
                              C.c_void_p]
        L.orc_noisy.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p]
        L.orc_clip_adam.restype = C.c_float
        L.orc_clip_adam.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_int64]
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def tree_start(size):
    return int(lib().orc_tree_start(int(size)))


def pow_priorities(raw, omega):
    raw = _c(raw, np.float32)
    out = np.empty_like(raw)
    lib().orc_pow_priorities(raw, float(omega), raw.size, out)
    return out


class OracleTree:
    """float32 heap-layout sum tree + SoA transition ring (reference memory.py:12-89)."""

    def __init__(self, size, with_data=True):
        self.size = int(size)
        self.tree_start = tree_start(size)
        self.sum_tree = np.zeros(self.tree_start + self.size, np.float32)
        self.index = 0
        self.full = False
        self.max = np.ones(1, np.float32)  # memory.py:20
        if with_data:
            self.frames = np.zeros((self.size, FRAME), np.uint8)
            self.timestep = np.zeros(self.size, np.int32)
            self.action = np.zeros(self.size, np.int32)
            self.reward = np.zeros(self.size, np.float32)
            self.nonterminal = np.zeros(self.size, np.uint8)

    def update(self, tree_idx, values):
        tree_idx = _c(tree_idx, np.int64)
        values = _c(values, np.float32)
        lib().orc_tree_update(self.sum_tree, tree_idx, values, tree_idx.size, self.max)

    def append(self, timestep, frame_u8, action, reward, nonterminal, value=None):
        i = self.index
        self.timestep[i] = timestep
        self.frames[i] = np.asarray(frame_u8, np.uint8).reshape(-1)
        self.action[i] = action
        self.reward[i] = reward
        self.nonterminal[i] = 1 if nonterminal else 0
        v = float(self.max[0]) if value is None else float(value)
        lib().orc_tree_set_leaf(self.sum_tree, i + self.tree_start, v)
        self.index = (i + 1) % self.size
        self.full = self.full or self.index == 0
        self.max[0] = max(np.float32(v), self.max[0])

    def find(self, values):
        values = _c(values, np.float64)
        B = values.size
        probs = np.empty(B, np.float32)
        didx = np.empty(B, np.int64)
        tidx = np.empty(B, np.int64)
        lib().orc_tree_find(self.sum_tree, self.tree_start, self.size, values, B, probs, didx, tidx)
        return probs, didx, tidx

    def total(self):
        return self.sum_tree[0]


def segment_samples(p_total, B, u01):
    u01 = _c(u01, np.float64)
    out = np.empty(B, np.float64)
    lib().orc_segment_samples(float(np.float32(p_total)), int(B), u01, out)
    return out


def batch_valid(didx, probs, head, capacity, n, history):
    didx = _c(didx, np.int64)
    probs = _c(probs, np.float32)
    return bool(lib().orc_batch_valid(didx, probs, didx.size, int(head), int(capacity), int(n), int(history)))


def is_weights(probs, p_total, count, beta):
    probs = _c(probs, np.float32)
    w = np.empty_like(probs)
    lib().orc_is_weights(probs, float(np.float32(p_total)), int(count), float(np.float32(beta)), probs.size, w)
    return w


def quantise_frame(frame_f32):
    f = _c(frame_f32, np.float32).reshape(-1)
    out = np.empty(FRAME, np.uint8)
    lib().orc_quantise_frame(f, out)
    return out


def gather(tree, didx, history, n, gamma_pow):
    didx = _c(didx, np.int64)
    B = didx.size
    gp = _c(gamma_pow, np.float32)
    states = np.empty((B, history, 84, 84), np.float32)
    nstates = np.empty((B, history, 84, 84), np.float32)
    actions = np.empty(B, np.int64)
    returns = np.empty(B, np.float32)
    nonterm = np.empty(B, np.float32)
    lib().orc_gather(tree.frames, tree.timestep, tree.action, tree.reward, tree.nonterminal, tree.size, didx, B,
                     int(history), int(n), gp, states.reshape(-1), nstates.reshape(-1), actions, returns, nonterm)
    return states, actions, returns, nstates, nonterm.reshape(B, 1)


def iter_state(tree, cur, history):
    out = np.empty((history, 84, 84), np.float32)
    lib().orc_iter_state(tree.frames, tree.timestep, tree.size, int(cur), int(history), out.reshape(-1))
    return out


def c51(q_on_s, q_on_ns, q_tg_ns, actions, returns, nonterminals, weights, support, vmin, vmax, delta_z, gamma_n):
    q_on_s = _c(q_on_s, np.float32)
    B, A, Z = q_on_s.shape
    loss = np.empty(B, np.float32)
    grad = np.empty((B, A, Z), np.float32)
    m = np.empty((B, Z), np.float32)
    astar = np.empty(B, np.int64)
    lib().orc_c51(q_on_s.reshape(-1), _c(q_on_ns, np.float32).reshape(-1), _c(q_tg_ns, np.float32).reshape(-1),
                  _c(actions, np.int64), _c(returns, np.float32).reshape(-1),
                  _c(nonterminals, np.float32).reshape(-1), _c(weights, np.float32), _c(support, np.float32),
                  float(np.float32(vmin)), float(np.float32(vmax)), float(np.float32(delta_z)),
                  float(np.float32(gamma_n)), B, A, Z, loss, grad.reshape(-1), m.ctypes.data, astar.ctypes.data)
    return loss, grad, m, astar


def noisy(x_in, x_out):
    x_in = _c(x_in, np.float32)
    x_out = _c(x_out, np.float32)
    w = np.empty((x_out.size, x_in.size), np.float32)
    b = np.empty(x_out.size, np.float32)
    lib().orc_noisy(x_in, x_out, x_in.size, x_out.size, w.reshape(-1), b)
    return w, b


def clip_adam(param, grad, exp_avg, exp_avg_sq, max_norm, lr, beta1, beta2, eps, step):
    """In-place on the four flat float32 arrays; returns the pre-clip gradient norm."""
    for a in (param, grad, exp_avg, exp_avg_sq):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    return float(lib().orc_clip_adam(param, grad, exp_avg, exp_avg_sq, param.size, float(max_norm), float(lr),
                                     float(beta1), float(beta2), float(eps), int(step)))
