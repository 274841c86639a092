"""Prioritised replay living entirely in B200 HBM.

Mirrors the class surface of the reference's memory.py (SegmentTree memory.py:12-89, ReplayMemory
memory.py:91-180) so main.py / test.py / agent.py of the reference run against it unchanged, but the
state is a structure of arrays on the device and every operation is one hand-written CUDA kernel
called through the C ABI in include/rainbow_b200.h:

    append             -> rb_append        (K5)   memory.py:105-108, 56-61
    sample             -> rb_tree_sample   (K1)   memory.py:124-132, 148-154
                          rb_gather        (K2)   memory.py:111-121, 134-146
    update_priorities  -> rb_tree_update   (K4)   memory.py:157-159, 23-48
    __next__           -> rb_iter_states          memory.py:166-178

HBM layout (per ReplayMemory): float32 sum tree [tree_start+size] in heap order (one pad float in
front so level runs are 128-byte aligned), uint8 frames [size,7056], int32 timestep/action,
float32 reward, uint8 nonterminal: 7069 B per transition like the reference's packed record, i.e.
7.07 GB at the default 1M capacity.

Randomness: `rng="philox"` (default) draws the stratified uniforms on the device (Philox4x32-10,
counter kept on the device so CUDA-graph replays advance it); `rng="numpy"` consumes the global
legacy numpy generator exactly like memory.py:129 does, which makes sampled indices bit-identical to
the reference for the same seed and tree (at the price of one small H2D copy and a status read).
"""
import numpy as np
import torch

from . import _lib

FRAME = 84 * 84

# memory.py:7 -- only used to exchange state with reference-format consumers (pickles)
Transition_dtype = np.dtype([("timestep", np.int32), ("state", np.uint8, (84, 84)), ("action", np.int32),
                             ("reward", np.float32), ("nonterminal", np.bool_)])


def _require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.RainbowB200Error(
            f"rainbow_b200.ReplayMemory needs a CUDA device, got '{device}': the replay lives in HBM and there "
            "is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class SegmentTree:
    """Device-resident counterpart of memory.py:12-89.  `index`, `full`, `size`, `tree_start` are host
    mirrors (cheap ints, advanced in lock step with the device copy in `ring_state`); `max`, `sum_tree`
    and `data` read the device and therefore synchronise -- they exist for inspection and pickling."""

    def __init__(self, size, device):
        size = int(size)
        if size <= 0 or size % 2:
            # the reference itself raises IndexError in _propagate_index for odd sizes
            raise ValueError("SegmentTree size must be a positive even number")
        self.device = _require_cuda(device)
        self.size = size
        self.tree_start = 2 ** (size - 1).bit_length() - 1
        self.index = 0
        self.full = False
        n = self.tree_start + size
        # one pad element in front: node j lives at element j+1, so the 32-node run of a tree level starts
        # on a 128-byte boundary
        self._tree_store = torch.zeros(n + 1, dtype=torch.float32, device=self.device)
        self.tree = self._tree_store[1:]
        self.frames = torch.zeros((size, FRAME), dtype=torch.uint8, device=self.device)
        self.timestep = torch.zeros(size, dtype=torch.int32, device=self.device)
        self.action = torch.zeros(size, dtype=torch.int32, device=self.device)
        self.reward = torch.zeros(size, dtype=torch.float32, device=self.device)
        self.nonterminal = torch.zeros(size, dtype=torch.uint8, device=self.device)
        self.ring_state = torch.zeros(5, dtype=torch.int64, device=self.device)  # head, full, t_episode, appended, ticket
        self.running_max = torch.ones(1, dtype=torch.float32, device=self.device)  # memory.py:20
        self._status = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._lib = _lib.load()

    # ---- pickling: the reference pickles the whole object graph (main.py:85-100) ------------------------------
    def __getstate__(self):
        """The reference's own field layout (memory.py:13-20): index, size, full, tree_start, sum_tree, data, max."""
        return dict(index=self.index, size=self.size, full=self.full, tree_start=self.tree_start, sum_tree=self.sum_tree,
                    data=self.data, max=self.max)

    def __setstate__(self, state):
        """Accepts the reference's SegmentTree.__dict__ (a file written by the reference, loaded with dropin/ shadowing
        the `memory` module).  The device arrays are created by _materialise() once the owning ReplayMemory knows the device."""
        if not all(k in state for k in ("index", "size", "full", "sum_tree", "data", "max")):
            raise _lib.RainbowB200Error("unknown SegmentTree pickle layout")
        self._pending = dict(state)
        self.size, self.index, self.full = int(state["size"]), int(state["index"]), bool(state["full"])
        self.tree_start = 2 ** (self.size - 1).bit_length() - 1

    def _materialise(self, device, t_episode=0):
        fields = self.__dict__.pop("_pending")
        SegmentTree.__init__(self, fields["size"], device)
        self.load_arrays(**reference_fields_to_ring(fields, t=t_episode))

    # ---- reference-style accessors -------------------------------------------------------------
    @property
    def max(self):
        return float(self.running_max.item())

    @property
    def sum_tree(self):
        return self.tree.cpu().numpy()

    @property
    def data(self):
        out = np.zeros(self.size, dtype=Transition_dtype)
        out["timestep"] = self.timestep.cpu().numpy()
        out["state"] = self.frames.cpu().numpy().reshape(self.size, 84, 84)
        out["action"] = self.action.cpu().numpy()
        out["reward"] = self.reward.cpu().numpy()
        out["nonterminal"] = self.nonterminal.cpu().numpy().astype(np.bool_)
        return out

    def total(self):
        return self.tree[0].item()

    # ---- operations ----------------------------------------------------------------------------
    def _as_dev(self, x, dtype):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device, non_blocking=True)

    def update(self, indices, values, omega=None, gate=None):
        """memory.py:44-48.  `values` are tree values; with `omega` given they are raw priorities and the
        kernel applies ^omega (memory.py:158).  `gate`: optional device int32 tensor; the launch is a no-op when its
        first element is 0 (the status word of a rejected sample batch)."""
        idx = self._as_dev(indices, torch.int64).reshape(-1)
        val = self._as_dev(values, torch.float32).reshape(-1)
        if idx.numel() != val.numel():
            raise ValueError("indices and values differ in length")
        _lib.check(self._lib.rb_tree_update(
            _lib.ptr(self.tree), self.tree_start, self.size, _lib.ptr(idx), _lib.ptr(val),
            0.0 if omega is None else float(omega), 1 if omega is None else 0, idx.numel(),
            _lib.ptr(self.running_max), _lib.ptr(self._status), _lib.ptr(gate), _lib.stream()))

    def find(self, values):
        """memory.py:79-82: float64 values -> (leaf values, data indices, tree indices), as device tensors."""
        v = self._as_dev(values, torch.float64).reshape(-1)
        B = v.numel()
        probs = torch.empty(B, dtype=torch.float32, device=self.device)
        didx = torch.empty(B, dtype=torch.int64, device=self.device)
        tidx = torch.empty(B, dtype=torch.int64, device=self.device)
        _lib.check(self._lib.rb_tree_find(_lib.ptr(self.tree), self.tree_start, self.size, _lib.ptr(v), B,
                                          _lib.ptr(probs), _lib.ptr(didx), _lib.ptr(tidx), _lib.stream()))
        return probs, didx, tidx

    def append_frame(self, last_frame, action, reward, terminal):
        """memory.py:56-61 with the record fields passed separately; the leaf gets the running max.
        `last_frame`: float32 [84*84] in device memory or PINNED host memory (read in place by the kernel)."""
        fptr = last_frame.data_ptr() if (not last_frame.is_cuda and last_frame.is_pinned()) else _lib.ptr(last_frame)
        _lib.check(self._lib.rb_append(
            _lib.ptr(self.tree), self.tree_start, self.size, _lib.ptr(self.frames), _lib.ptr(self.timestep),
            _lib.ptr(self.action), _lib.ptr(self.reward), _lib.ptr(self.nonterminal), _lib.ptr(self.ring_state),
            _lib.ptr(self.running_max), fptr, int(action), float(reward), 1 if terminal else 0,
            _lib.stream()))
        self.index = (self.index + 1) % self.size
        self.full = self.full or self.index == 0

    def get(self, data_index):
        """memory.py:85-86 (host copy of the selected records; inspection only)."""
        idx = np.asarray(data_index) % self.size
        flat = torch.as_tensor(idx.reshape(-1), dtype=torch.int64, device=self.device)
        out = np.zeros(idx.size, dtype=Transition_dtype)
        out["timestep"] = self.timestep[flat].cpu().numpy()
        out["state"] = self.frames[flat].cpu().numpy().reshape(-1, 84, 84)
        out["action"] = self.action[flat].cpu().numpy()
        out["reward"] = self.reward[flat].cpu().numpy()
        out["nonterminal"] = self.nonterminal[flat].cpu().numpy().astype(np.bool_)
        return out.reshape(idx.shape)

    # ---- bulk state exchange (tests, pickling, synthetic fill) -----------------------------------
    def load_arrays(self, sum_tree=None, frames=None, timestep=None, action=None, reward=None, nonterminal=None,
                    index=None, full=None, t_episode=0, max_value=None):
        def put(dst, src, dt):
            if src is not None:
                dst.copy_(torch.as_tensor(np.ascontiguousarray(src), dtype=dt).reshape(dst.shape))

        put(self.tree, sum_tree, torch.float32)
        put(self.frames, frames, torch.uint8)
        put(self.timestep, timestep, torch.int32)
        put(self.action, action, torch.int32)
        put(self.reward, reward, torch.float32)
        put(self.nonterminal, None if nonterminal is None else np.asarray(nonterminal).astype(np.uint8), torch.uint8)
        if index is not None:
            self.index = int(index)
        if full is not None:
            self.full = bool(full)
        self.ring_state.copy_(torch.tensor([self.index, int(self.full), int(t_episode), 0, 0], dtype=torch.int64))
        if max_value is not None:
            self.running_max.fill_(float(max_value))


class _SampleWorkspace:
    """Output buffers of one sample() call for a given batch size."""

    def __init__(self, B, history, device):
        f32, i64 = torch.float32, torch.int64
        self.B = B
        self.probs = torch.empty(B, dtype=f32, device=device)
        self.data_idx = torch.empty(B, dtype=i64, device=device)
        self.tree_idx = torch.empty(B, dtype=i64, device=device)
        self.weights = torch.empty(B, dtype=f32, device=device)
        # one allocation, s rows then s' rows: the learner can push [s; s'] through the online conv body in a single pass
        self.both_states = torch.empty((2 * B, history, 84, 84), dtype=f32, device=device)
        self.states, self.next_states = self.both_states[:B], self.both_states[B:]
        self.actions = torch.empty(B, dtype=i64, device=device)
        self.returns = torch.empty(B, dtype=f32, device=device)
        self.nonterminals = torch.empty((B, 1), dtype=f32, device=device)
        self.status = torch.zeros(4, dtype=torch.int32, device=device)   # ok flag, draws used, rejected batches so far, -

    def as_tuple(self):
        return (self.tree_idx, self.states, self.actions, self.returns, self.next_states, self.nonterminals,
                self.weights)


class ReplayMemory:
    """Drop-in for memory.py:91-180 (`ReplayMemory(args, capacity)`).

    Extra keyword arguments (not in the reference): rng ("philox" | "numpy"), seed, max_attempts,
    strict (read the kernel's status word after every sample and raise if the batch was rejected
    max_attempts times -- costs a device synchronisation)."""

    APPEND_BATCH = 8  # RB_APPEND_BATCH

    def __init__(self, args, capacity, rng="philox", seed=None, max_attempts=64, strict=False, defer_appends=False):
        self.device = _require_cuda(args.device)
        self.capacity = int(capacity)
        self.history = int(args.history_length)
        self.discount = args.discount
        self.n = int(args.multi_step)
        self.priority_weight = args.priority_weight  # beta; main.py:161 overwrites this attribute every step
        self.priority_exponent = args.priority_exponent
        if self.history + self.n > 64:
            raise ValueError("history_length + multi_step must not exceed 64")
        if rng not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' or 'numpy'")
        self.rng = rng
        self.max_attempts = int(max_attempts)
        self.strict = bool(strict)
        # defer_appends: append() only queues (frame reference + fields); the queue is written by ONE rb_append_batch
        # launch before the next read of the replay (sample / update_priorities / iteration / pickling) or when it holds
        # APPEND_BATCH transitions.  The caller must not modify a queued frame tensor before the flush.
        # [round-1 status: off by default, not yet exercised on hardware]
        self.defer_appends = bool(defer_appends)
        self._queue = []
        self.t = 0  # in-episode step of the next append (memory.py:100); device copy in ring_state[2]
        # memory.py:101: [gamma**i] in Python doubles, then float32
        self.n_step_scaling = torch.tensor([self.discount ** i for i in range(self.n)], dtype=torch.float32,
                                           device=self.device)
        self.transitions = SegmentTree(self.capacity, self.device)
        if seed is None:   # data-parallel ranks launched with one torch seed must not draw the same stratified uniforms
            from .dist import GradSync, shard_seed
            rank = GradSync().rank
            seed = torch.initial_seed() if rank == 0 else shard_seed(torch.initial_seed(), rank)
        self.seed = int(seed) & (2 ** 64 - 1)
        self._rng_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._beta_dev = torch.full((1,), float(self.priority_weight), dtype=torch.float32, device=self.device)
        self._beta_pushed = float(self.priority_weight)
        self._lib = _lib.load()
        self._last = None
        self._stage = None

    def push_beta(self):
        """Mirror the host attribute `priority_weight` (main.py:161 rewrites it every step) into the device
        scalar the sampling kernel reads, so a captured CUDA graph sees the current beta."""
        b = float(self.priority_weight)
        if b != self._beta_pushed:
            self._beta_dev.fill_(b)
            self._beta_pushed = b

    # ---- append --------------------------------------------------------------------------------
    STAGE_SLOTS = 16   # owned pinned staging frames for host-resident states (2 x APPEND_BATCH)

    def _stage_host_frame(self, last):
        """Copy a HOST frame into an owned pinned slot and return the slot (float32 [84*84], 16-byte aligned).

        The reference copies the frame synchronously (memory.py:106); an asynchronous H2D copy straight from the caller's
        buffer would race with an env that rewrites its (pinned) frame buffer in place.  The slot is reused only after the
        kernel that consumed it has finished (event per slot), so the caller may do whatever it likes with `state` as
        soon as append() returns."""
        if self._stage is None:
            self._stage = torch.empty((self.STAGE_SLOTS, FRAME), dtype=torch.float32).pin_memory()
            self._stage_evt = [None] * self.STAGE_SLOTS
            self._stage_next = 0
        k = self._stage_next
        self._stage_next = (k + 1) % self.STAGE_SLOTS
        if self._stage_evt[k] is not None:
            self._stage_evt[k].synchronize()
            self._stage_evt[k] = None
        slot = self._stage[k]
        slot.view(84, 84).copy_(last)      # host memcpy (+ dtype conversion / de-striding if needed)
        return k, slot

    def _release_stage_slots(self, slots):
        if slots:
            evt = torch.cuda.Event()
            evt.record(torch.cuda.current_stream(self.device))
            for k in slots:
                self._stage_evt[k] = evt

    def append(self, state, action, reward, terminal):
        """memory.py:105-108.  `state` is the float32 [history,84,84] frame stack in [0,1]; only the newest
        frame is stored (quantised to uint8 on the device).  Host frames are staged through owned pinned memory and read
        by the kernel in place (no separate H2D copy launch); device frames are read in place."""
        last = state[-1]
        slot_id = None
        if not last.is_cuda:
            slot_id, last = self._stage_host_frame(last)
        else:
            last = last.to(torch.float32).contiguous()
            if last.data_ptr() % 16:
                last = last.clone()
        if self.defer_appends:
            # queued by reference: a DEVICE frame must not be modified by the caller before the flush (main.py's env
            # builds a fresh state tensor every step); host frames are already copied into the staging ring
            self._queue.append((last, int(action), float(reward), bool(terminal), slot_id))
            tr = self.transitions
            tr.index = (tr.index + 1) % tr.size       # host mirrors advance now, the device copy at the flush
            tr.full = tr.full or tr.index == 0
            self.t = 0 if terminal else self.t + 1
            if len(self._queue) >= self.APPEND_BATCH:
                self.flush_appends()
            return
        self.transitions.append_frame(last, action, reward, terminal)
        if slot_id is not None:
            self._release_stage_slots([slot_id])
        self.t = 0 if terminal else self.t + 1

    def flush_appends(self):
        """Write the queued transitions (defer_appends=True) with one rb_append_batch launch."""
        if not self._queue:
            return
        import ctypes as C
        q, self._queue = self._queue, []
        k = len(q)
        tr = self.transitions
        frames = (C.c_void_p * k)(*[e[0].data_ptr() for e in q])   # device or pinned-host pointers (UVA)
        acts = (C.c_int32 * k)(*[e[1] for e in q])
        rews = (C.c_float * k)(*[e[2] for e in q])
        terms = (C.c_int32 * k)(*[1 if e[3] else 0 for e in q])
        _lib.check(self._lib.rb_append_batch(
            _lib.ptr(tr.tree), tr.tree_start, tr.size, _lib.ptr(tr.frames), _lib.ptr(tr.timestep), _lib.ptr(tr.action),
            _lib.ptr(tr.reward), _lib.ptr(tr.nonterminal), _lib.ptr(tr.ring_state), _lib.ptr(tr.running_max), frames, acts,
            rews, terms, k, _lib.stream()))
        self._release_stage_slots([e[4] for e in q if e[4] is not None])
        self._flushed_refs = q   # keep device frames alive until the next flush (the launch is asynchronous)

    # ---- sample --------------------------------------------------------------------------------
    def _launch_sample(self, ws, u01=None, attempts=0):
        tr = self.transitions
        L = self._lib
        _lib.check(L.rb_tree_sample(
            _lib.ptr(tr.tree), tr.tree_start, tr.size, _lib.ptr(tr.ring_state), self.n, self.history,
            _lib.ptr(u01), attempts, self.seed, _lib.ptr(self._rng_counter), ws.B, float(self.priority_weight),
            _lib.ptr(self._beta_dev), self.max_attempts, _lib.ptr(ws.probs), _lib.ptr(ws.data_idx), _lib.ptr(ws.tree_idx),
            _lib.ptr(ws.weights), _lib.ptr(ws.status), _lib.stream()))

    def _launch_gather(self, ws):
        tr = self.transitions
        _lib.check(self._lib.rb_gather(
            _lib.ptr(tr.frames), _lib.ptr(tr.timestep), _lib.ptr(tr.action), _lib.ptr(tr.reward),
            _lib.ptr(tr.nonterminal), tr.size, _lib.ptr(ws.data_idx), ws.B, self.history, self.n,
            _lib.ptr(self.n_step_scaling), _lib.ptr(ws.states), _lib.ptr(ws.next_states), _lib.ptr(ws.actions),
            _lib.ptr(ws.returns), _lib.ptr(ws.nonterminals), _lib.stream()))

    def sample_into(self, ws):
        """Device-RNG sample into caller-owned buffers: two launches, no synchronisation (graph capturable).
        The caller is responsible for push_beta() and flush_appends() (outside any graph capture)."""
        self._launch_sample(ws)
        self._launch_gather(ws)
        self._last = ws
        return ws.as_tuple()

    def sample(self, batch_size):
        """memory.py:148-155.  Returns (tree_idxs, states, actions, returns, next_states, nonterminals, weights),
        all device tensors (the reference returns tree_idxs as numpy; update_priorities takes either)."""
        ws = _SampleWorkspace(int(batch_size), self.history, self.device)
        self.flush_appends()
        self.push_beta()
        if self.rng == "numpy":
            # consume the legacy global generator exactly like np.random.uniform(0, seg, [B]) does
            for _ in range(100000):
                u = torch.from_numpy(np.random.random_sample(ws.B)).to(self.device)
                self._launch_sample(ws, u01=u, attempts=1)
                if int(ws.status[0].item()) == 1:
                    break
            else:  # pragma: no cover
                raise _lib.RainbowB200Error("no valid batch after 100000 draws")
            self._launch_gather(ws)
            self._last = ws
            return ws.as_tuple()
        out = self.sample_into(ws)
        if self.strict:
            self.check_last_sample()
        return out

    def sample_gate(self):
        """Device int32 status words of the most recent sample (None in numpy-rng mode, where the host loop only ever
        returns valid batches): element 0 is 0 when that batch was rejected `max_attempts` times.  The kernel has then
        zeroed the batch's importance weights, and handing this tensor as `gate` to update_priorities() / the optimiser
        makes the whole update a no-op -- the device-side stand-in for the reference's unbounded redraw loop
        (memory.py:128-132)."""
        return None if (self._last is None or self.rng == "numpy") else self._last.status

    def rejected_batches(self):
        """Synchronises.  Number of device-RNG batches (since this workspace was created) that stayed invalid after
        `max_attempts` redraws and were therefore skipped."""
        return 0 if self._last is None else int(self._last.status[2].item())

    def check_last_sample(self):
        """Synchronises; raises if the most recent device-RNG sample exhausted max_attempts redraws (strict mode)."""
        if self._last is not None and int(self._last.status[0].item()) != 1:
            raise _lib.RainbowB200Error(
                f"replay sampling rejected {self.max_attempts} consecutive batches (buffer too empty around the "
                "write head, or zero-priority leaves): that batch was skipped (zero weights, no update)")

    # ---- priorities ----------------------------------------------------------------------------
    def update_priorities(self, idxs, priorities, gate=None):
        """memory.py:157-159: raw per-sample losses -> ^omega -> leaves -> propagate to the root.
        `gate` (optional, see sample_gate()): skip the write-back of a rejected batch on the device."""
        if self._queue and not torch.cuda.is_current_stream_capturing():
            self.flush_appends()
        self.transitions.update(idxs, priorities, omega=self.priority_exponent, gate=gate)

    # ---- validation iterator (memory.py:162-180) -------------------------------------------------
    _ITER_CHUNK = 64

    def __iter__(self):
        self.flush_appends()
        self.current_idx = 0
        self._iter_buf = None
        self._iter_base = 0
        return self

    def iter_states(self, first, count):
        """Iterator states (memory.py:166-178) for current_idx = first .. first+count-1 in one launch:
        device float32 [count, history, 84, 84] (backward-only blanking, negative indices wrap)."""
        self.flush_appends()
        tr = self.transitions
        buf = torch.empty((count, self.history, 84, 84), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.rb_iter_states(_lib.ptr(tr.frames), _lib.ptr(tr.timestep), tr.size, int(first), int(count),
                                            self.history, _lib.ptr(buf), _lib.stream()))
        return buf

    def __next__(self):
        if self.current_idx == self.capacity:
            raise StopIteration
        if self._iter_buf is None or self.current_idx >= self._iter_base + self._iter_buf.shape[0]:
            count = min(self._ITER_CHUNK, self.capacity - self.current_idx)
            self._iter_buf, self._iter_base = self.iter_states(self.current_idx, count), self.current_idx
        state = self._iter_buf[self.current_idx - self._iter_base]
        self.current_idx += 1
        return state

    next = __next__

    # ---- pickling (main.py:85-100 pickles the whole object) --------------------------------------
    def __getstate__(self):
        """Own compact layout (plain numpy arrays, structure of arrays).  For a file the REFERENCE can load use
        save_reference_pickle()."""
        self.flush_appends()
        tr = self.transitions
        return dict(
            version=1, capacity=self.capacity, history=self.history, discount=self.discount, n=self.n,
            priority_weight=self.priority_weight, priority_exponent=self.priority_exponent, t=self.t, rng=self.rng,
            seed=self.seed, max_attempts=self.max_attempts, strict=self.strict, device=str(self.device),
            rng_counter=int(self._rng_counter.item()), index=tr.index, full=tr.full, max=tr.max,
            sum_tree=tr.sum_tree, frames=tr.frames.cpu().numpy(), timestep=tr.timestep.cpu().numpy(),
            action=tr.action.cpu().numpy(), reward=tr.reward.cpu().numpy(),
            nonterminal=tr.nonterminal.cpu().numpy())

    def _init_runtime(self, rng_counter=0):
        self.n_step_scaling = torch.tensor([self.discount ** i for i in range(self.n)], dtype=torch.float32,
                                           device=self.device)
        self.defer_appends, self._queue = False, []
        self._rng_counter = torch.tensor([int(rng_counter)], dtype=torch.int64, device=self.device)
        self._beta_dev = torch.full((1,), float(self.priority_weight), dtype=torch.float32, device=self.device)
        self._beta_pushed = float(self.priority_weight)
        self._lib = _lib.load()
        self._last = None
        self._stage = None

    def __setstate__(self, s):
        if "version" not in s and "transitions" in s:
            return self._setstate_reference(s)
        self.device = _require_cuda(s["device"])
        self.capacity, self.history, self.discount, self.n = s["capacity"], s["history"], s["discount"], s["n"]
        self.priority_weight, self.priority_exponent, self.t = s["priority_weight"], s["priority_exponent"], s["t"]
        self.rng, self.seed, self.max_attempts, self.strict = s["rng"], s["seed"], s["max_attempts"], s["strict"]
        self.transitions = SegmentTree(self.capacity, self.device)
        self.transitions.load_arrays(s["sum_tree"], s["frames"], s["timestep"], s["action"], s["reward"],
                                     s["nonterminal"], s["index"], s["full"], s["t"], s["max"])
        self._init_runtime(s["rng_counter"])

    def _setstate_reference(self, s):
        """A memory file written by the REFERENCE (its ReplayMemory.__dict__, memory.py:93-102, with the AoS
        Transition_dtype `data` and the truncated `sum_tree` of memory.py:13-20): rebuilt in HBM.  A CPU device in the
        file is mapped to the current CUDA device (the replay has no host variant)."""
        dev = torch.device(s.get("device", "cuda"))
        self.device = _require_cuda(dev if dev.type == "cuda" else "cuda")
        self.capacity, self.history, self.discount, self.n = int(s["capacity"]), int(s["history"]), s["discount"], int(s["n"])
        self.priority_weight, self.priority_exponent, self.t = s["priority_weight"], s["priority_exponent"], int(s["t"])
        self.rng, self.max_attempts, self.strict = "philox", 64, False
        self.seed = int(torch.initial_seed()) & (2 ** 64 - 1)
        tr = s["transitions"]
        if isinstance(tr, SegmentTree):
            tr._materialise(self.device, self.t)
        else:   # any object carrying the reference's fields
            fields = {k: getattr(tr, k) for k in ("index", "size", "full", "sum_tree", "data", "max")}
            tr = SegmentTree(int(fields["size"]), self.device)
            tr.load_arrays(**reference_fields_to_ring(fields, t=self.t))
        self.transitions = tr
        self._init_runtime()

    def reference_state(self, device="cpu"):
        """(ReplayMemory.__dict__, SegmentTree.__dict__) exactly as the reference's objects hold them."""
        self.flush_appends()
        dev = torch.device(device)
        mem = dict(device=dev, capacity=self.capacity, history=self.history, discount=self.discount, n=self.n,
                   priority_weight=self.priority_weight, priority_exponent=self.priority_exponent, t=self.t,
                   n_step_scaling=self.n_step_scaling.to(dev))
        return mem, self.transitions.__getstate__()


def save_reference_pickle(mem, file, device="cpu", protocol=None):
    """Write `mem` as a pickle the UNMODIFIED reference loads with pickle.load (main.py:85-91): the stream names the
    classes `memory.ReplayMemory` / `memory.SegmentTree` and carries the reference's own field layout, so inside the
    reference process it unpickles into the reference's classes (and, with dropin/ on the path, into ours)."""
    import pickle
    import sys
    import types
    mem_state, tree_state = mem.reference_state(device)
    stand_in = types.ModuleType("memory")
    tree_cls = type("SegmentTree", (), {"__module__": "memory"})
    mem_cls = type("ReplayMemory", (), {"__module__": "memory"})
    stand_in.SegmentTree, stand_in.ReplayMemory = tree_cls, mem_cls
    tree = tree_cls()
    tree.__dict__.update(tree_state)
    obj = mem_cls()
    obj.__dict__.update(mem_state, transitions=tree)
    saved = sys.modules.get("memory")
    sys.modules["memory"] = stand_in      # pickle verifies that memory.ReplayMemory is the class being written
    try:
        pickle.dump(obj, file, protocol=protocol)
    finally:
        if saved is None:
            del sys.modules["memory"]
        else:
            sys.modules["memory"] = saved


def ring_to_reference_fields(state):
    """Host-side helper for format exchange: the pickled dict above -> the fields of the reference's
    SegmentTree (memory.py:13-20): index, size, full, tree_start, sum_tree, data (AoS Transition_dtype), max."""
    size = state["capacity"]
    data = np.zeros(size, dtype=Transition_dtype)
    data["timestep"] = state["timestep"]
    data["state"] = np.asarray(state["frames"]).reshape(size, 84, 84)
    data["action"] = state["action"]
    data["reward"] = state["reward"]
    data["nonterminal"] = np.asarray(state["nonterminal"]).astype(np.bool_)
    return dict(index=state["index"], size=size, full=state["full"], tree_start=2 ** (size - 1).bit_length() - 1,
                sum_tree=np.asarray(state["sum_tree"], dtype=np.float32), data=data, max=state["max"])


def reference_fields_to_ring(fields, t=0):
    """Inverse of ring_to_reference_fields: SegmentTree attributes of the reference -> load_arrays kwargs."""
    data = fields["data"]
    size = int(fields["size"])
    return dict(sum_tree=fields["sum_tree"], frames=np.ascontiguousarray(data["state"]).reshape(size, FRAME),
                timestep=data["timestep"], action=data["action"], reward=data["reward"],
                nonterminal=data["nonterminal"].astype(np.uint8), index=fields["index"], full=fields["full"],
                t_episode=t, max_value=fields["max"])
