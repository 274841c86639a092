"""Rainbow learner: the reference's Agent API (agent.py:12-118) over B200-native kernels.

One `learn(mem)` (agent.py:61-100) is:

    K1 rb_tree_sample + K2 rb_gather          (mem.sample, memory.py:148-155)
    3 x conv body (torch: cuDNN)              (agent.py:66,71,75 -> model.py:70-71)
    K6 rb_noisy_resample (target net)         (agent.py:74)
    fused noisy dueling heads (rb_head_forward) on the conv features -- online net on [s; s'], target on s'
    K3 rb_c51_dueling_loss_grad               (agent.py:67,72-73,76-96 + softmax halves of model.py:76-79 + model.py:75)
    rb_head_backward (16 head gradients + d conv features), torch autograd backward through the online convs
    [NCCL all-reduce of the flat gradient when world_size > 1]
    K7 rb_clip_adam                           (agent.py:97-98)
    K4 rb_tree_update                         (agent.py:100 -> memory.py:157-159)

Nothing in that chain synchronises with the host, so the whole update is captured into one CUDA graph
(`cuda_graph=True`, the default) and replayed: the update is launch-latency bound otherwise
(the reference issues ~600 ATen ops per update).
"""
import os

import numpy as np
import torch

from . import _lib
from .dist import GradSync
from .memory import ReplayMemory, _SampleWorkspace
from .model import DQN


def c51_loss_grad(q_online_s, q_online_ns, q_target_ns, actions, returns, nonterminals, weights, support, vmin, vmax,
                  delta_z, gamma_n, loss=None, grad=None, m_out=None, astar_out=None):
    """Launch K3 on pre-softmax logits [B,A,Z]; returns (loss[B], grad[B,A,Z])."""
    B, A, Z = q_online_s.shape
    dev = q_online_s.device
    if loss is None:
        loss = torch.empty(B, dtype=torch.float32, device=dev)
    if grad is None:
        grad = torch.empty((B, A, Z), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().rb_c51_loss_grad(
        _lib.ptr(q_online_s), _lib.ptr(q_online_ns), _lib.ptr(q_target_ns), _lib.ptr(actions), _lib.ptr(returns),
        _lib.ptr(nonterminals), _lib.ptr(weights), _lib.ptr(support), float(vmin), float(vmax), float(delta_z),
        float(gamma_n), B, A, Z, _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(m_out), _lib.ptr(astar_out), _lib.stream()))
    return loss, grad


def c51_dueling_loss_grad(z_online, z_target, actions_n, atoms, actions, returns, nonterminals, weights, support, vmin, vmax,
                          delta_z, gamma_n, m_out=None, astar_out=None):
    """K3 fed straight by the fused heads (rb_c51_dueling_loss_grad): z_online [2B, Z(1+A)] (s rows, then s' rows),
    z_target [B, Z(1+A)]; returns (loss[B], dz[B, Z(1+A)]) with dz = d mean(w*loss) / d (z_value | z_advantage)."""
    B = actions.shape[0]
    loss = torch.empty(B, dtype=torch.float32, device=actions.device)
    dz = torch.empty((B, atoms * (1 + actions_n)), dtype=torch.float32, device=actions.device)
    _lib.check(_lib.load().rb_c51_dueling_loss_grad(
        _lib.ptr(z_online), _lib.ptr(z_target), actions_n, atoms, _lib.ptr(actions), _lib.ptr(returns),
        _lib.ptr(nonterminals), _lib.ptr(weights), _lib.ptr(support), float(vmin), float(vmax), float(delta_z),
        float(gamma_n), B, _lib.ptr(loss), _lib.ptr(dz), _lib.ptr(m_out), _lib.ptr(astar_out), _lib.stream()))
    return loss, dz


class FusedClipAdam:
    """clip_grad_norm_ + Adam (agent.py:46,97-98) over ONE flat parameter buffer.

    The network's parameters are re-pointed at slices of `flat_param` (each slice starts on a 256-byte
    boundary; padding stays zero), their .grad at slices of `flat_grad`, so the optimiser step is two kernel
    launches (sum of squares, then clip+Adam) and the multi-GPU gradient exchange is a single all-reduce."""

    ALIGN = 64  # elements

    def __init__(self, net, lr, eps, max_norm, betas=(0.9, 0.999), peer=False):
        named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        self.offsets, off = [], 0
        self.conv_end = None  # flat offset where the first noisy-head parameter starts (convs come first)
        for n, p in named:
            if self.conv_end is None and n.startswith("fc_"):
                self.conv_end = off
            self.offsets.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.numel = off
        if self.conv_end is None:
            self.conv_end = off
        self.lr, self.eps, self.max_norm, self.betas = float(lr), float(eps), float(max_norm), betas
        self.peer = None
        if peer:   # [round-1 status: untested on hardware] buffers in symmetric memory, optimiser fused with the exchange
            from .peer import PeerOptimizerState
            self.peer = PeerOptimizerState(off, dev)
            self.numel = off = self.peer.numel
            self.flat_param, self.flat_grad = self.peer.flat_param, self.peer.flat_grad
            self.exp_avg, self.exp_avg_sq = self.peer.exp_avg, self.peer.exp_avg_sq
        else:
            self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
            self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
            self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.step_count = self.peer.step_count if self.peer is not None else torch.zeros(1, dtype=torch.int64, device=dev)
        self.grad_norm = self.peer.grad_norm if self.peer is not None else torch.zeros(1, dtype=torch.float32, device=dev)
        self._lib = _lib.load()
        self._partial = torch.zeros(self._lib.rb_clip_adam_scratch_elems(), dtype=torch.float64, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.flat_param[o:o + n].copy_(p.reshape(-1))
                p.data = self.flat_param[o:o + n].view_as(p)
                p.grad = self.flat_grad[o:o + n].view_as(p)

    def zero_grad(self):
        self.flat_grad.zero_()

    def zero_conv_grad(self):
        """Only the conv parameters accumulate through autograd on the fused path; the head kernels overwrite theirs."""
        self.flat_grad[:self.conv_end].zero_()

    def step(self, grad_scale=1.0):
        if self.peer is not None:   # reduce-scatter + clip + Adam + all-gather over peer memory (1/world folded in)
            self.peer.step(self.max_norm, self.lr, self.betas, self.eps)
            return
        _lib.check(self._lib.rb_clip_adam(
            _lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
            self.numel, float(grad_scale), self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps,
            _lib.ptr(self.step_count), _lib.ptr(self._partial), _lib.ptr(self.grad_norm), _lib.stream()))

    def state_dict(self):
        return dict(step=int(self.step_count.item()), exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(),
                    lr=self.lr, eps=self.eps, betas=self.betas, max_norm=self.max_norm)

    def load_state_dict(self, sd):
        self.step_count.fill_(int(sd["step"]))
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


class Agent:
    def __init__(self, args, env):
        self.device = torch.device(args.device)
        if self.device.type != "cuda":
            raise _lib.RainbowB200Error(f"rainbow_b200.Agent needs a CUDA device, got '{self.device}' (no CPU fallback)")
        self.action_space = env.action_space()
        self.atoms = args.atoms
        self.Vmin = args.V_min
        self.Vmax = args.V_max
        self.support = torch.linspace(args.V_min, args.V_max, self.atoms).to(device=self.device)  # agent.py:18
        self.delta_z = (args.V_max - args.V_min) / (self.atoms - 1)
        self.batch_size = args.batch_size
        self.n = args.multi_step
        self.discount = args.discount
        self.norm_clip = args.norm_clip

        self.online_net = DQN(args, self.action_space).to(device=self.device)
        model_path = getattr(args, "model", None)
        if model_path:  # agent.py:26-36: pretrained weights, with the old conv key names re-mapped
            if not os.path.isfile(model_path):
                raise FileNotFoundError(model_path)
            state_dict = torch.load(model_path, map_location="cpu")
            for i, old in enumerate(("conv1", "conv2", "conv3")):
                for kind in ("weight", "bias"):
                    if f"{old}.{kind}" in state_dict:
                        state_dict[f"convs.{2 * i}.{kind}"] = state_dict.pop(f"{old}.{kind}")
            self.online_net.load_state_dict(state_dict)
            print("Loading pretrained model: " + model_path)
        self.online_net.train()

        self.target_net = DQN(args, self.action_space).to(device=self.device)
        self.sync = GradSync()  # no-op unless torch.distributed is initialised with world_size > 1
        # distinct noise streams per net and per rank
        self.online_net.noise_seed = (self.online_net.noise_seed * 2 + 1 + 7919 * self.sync.rank) & (2 ** 63 - 1)
        self.target_net.noise_seed = (self.target_net.noise_seed * 2 + 2 + 7919 * self.sync.rank) & (2 ** 63 - 1)

        self.peer_optimizer = bool(getattr(args, "peer_optimizer", False)) and self.sync.enabled
        self.optimiser = FusedClipAdam(self.online_net, lr=args.learning_rate, eps=args.adam_eps, max_norm=self.norm_clip,
                                       peer=self.peer_optimizer)
        self.sync.broadcast_(self.optimiser.flat_param)  # identical initial parameters on every rank
        if self.peer_optimizer:
            self.sync.exchange = False   # the optimiser step does the gradient exchange itself
        self.update_target_net()
        self.target_net.train()
        for p in self.target_net.parameters():
            p.requires_grad = False

        self.use_cuda_graph = bool(getattr(args, "cuda_graph", True))
        self.use_fused_head = bool(getattr(args, "fused_head", True))
        self._streams = None
        self._graph = None
        self._graph_key = None
        self._ws = None
        self._learn_calls = 0
        self.last_loss = None  # per-sample losses of the most recent update (device tensor)

    # ---- acting / evaluation (agent.py:49-59,110-118) ---------------------------------------------
    def reset_noise(self):
        self.online_net.reset_noise()

    def act(self, state):
        with torch.no_grad():
            return (self.online_net(state.unsqueeze(0)) * self.support).sum(2).argmax(1).item()

    def act_e_greedy(self, state, epsilon=0.001):
        return np.random.randint(0, self.action_space) if np.random.random() < epsilon else self.act(state)

    def evaluate_q(self, state):
        with torch.no_grad():
            return (self.online_net(state.unsqueeze(0)) * self.support).sum(2).max(1)[0].item()

    def train(self):
        self.online_net.train()

    def eval(self):
        self.online_net.eval()

    def update_target_net(self):
        self.target_net.load_state_dict(self.online_net.state_dict())

    def save(self, path, name="model.pth"):
        torch.save(self.online_net.state_dict(), os.path.join(path, name))

    # ---- the update ------------------------------------------------------------------------------
    def _fused_path(self, B):
        on = self.online_net
        return self.use_fused_head and B <= 32 and on.training and on.fused_ok(2 * B) and self.target_net.fused_ok(B)

    def _side_streams(self):
        if self._streams is None:
            self._streams = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        return self._streams

    def _update_fused(self, batch, target_noise=None, after_loss=None):
        """agent.py:66-98 with the fused head: torch only runs the conv bodies (forward x3, backward x1).
        The three network passes are independent until the loss, and every conv kernel of this size leaves most of the
        148 SMs idle, so they run as three concurrent branches (fork/join with events; inside the captured CUDA graph
        they become parallel branches): online(s) with autograd on the caller's stream, online(s') and the whole
        target pass (noise draw, convs, head) on two side streams."""
        idxs, states, actions, returns, next_states, nonterminals, weights = batch
        on, tg = self.online_net, self.target_net
        B = states.shape[0]
        main = torch.cuda.current_stream(self.device)
        s_ns, s_tg = self._side_streams()
        fork = torch.cuda.Event()
        fork.record(main)
        with torch.cuda.stream(s_tg), torch.no_grad():
            s_tg.wait_event(fork)
            if target_noise is None:
                tg.reset_noise()                                            # agent.py:74
            else:
                tg.reset_noise(*target_noise)
            x_t = tg.features_nograd(next_states)
            z_t, _, _ = tg.head().forward(x_t)
            done_tg = torch.cuda.Event()
            done_tg.record(s_tg)
        with torch.cuda.stream(s_ns), torch.no_grad():
            s_ns.wait_event(fork)
            x_ns = on.features_nograd(next_states)
            done_ns = torch.cuda.Event()
            done_ns.record(s_ns)
        manual = on.manual_conv_ok(states)
        if manual:
            with torch.no_grad():
                acts = on.conv_forward_saving(states)  # library kernels, backward scheduled by hand below
            x_s = acts[-1].view(B, -1)
        else:
            x_s = on.features(states)                  # autograd graph: convs only
        with torch.no_grad():
            xs_d = x_s.detach()
            main.wait_event(done_ns)
            x_ns.record_stream(main)
            z_on, h_on, p_on = on.head().forward(xs_d, x_ns)              # rows [0,B) = s, [B,2B) = s'
            main.wait_event(done_tg)
            loss, dz = c51_dueling_loss_grad(z_on, z_t, self.action_space, self.atoms, actions, returns, nonterminals, weights,
                                             self.support, self.Vmin, self.Vmax, self.delta_z, self.discount ** self.n)
            wb_done = None
            if after_loss is not None:
                # the priority write-back (agent.py:100) needs nothing but the per-sample losses: it runs on a side
                # stream beside the whole backward instead of at the end of the critical path
                loss_ready = torch.cuda.Event()
                loss_ready.record(main)
                with torch.cuda.stream(s_ns):
                    s_ns.wait_event(loss_ready)
                    after_loss(loss)
                    wb_done = torch.cuda.Event()
                    wb_done.record(s_ns)
            dh = torch.empty((B, 2 * on.hidden_size), dtype=torch.float32, device=self.device)
            dx = torch.empty_like(xs_d)
            if manual:
                # dx comes back already masked by the last conv layer's ReLU; conv gradients are overwritten.
                # The layer-2 weight gradient feeds nothing but the optimiser: it runs beside the dh -> layer-1 chain.
                hd = on.head()
                dz_ready = torch.cuda.Event()
                dz_ready.record(main)
                with torch.cuda.stream(s_tg):
                    s_tg.wait_event(dz_ready)
                    hd.backward(p_on, xs_d, h_on[:B], dz, dh, dx, parts=hd.BWD_WGRAD2)
                    w2_done = torch.cuda.Event()
                    w2_done.record(s_tg)
                hd.backward(p_on, xs_d, h_on[:B], dz, dh, dx, relu_mask_x=True, parts=hd.BWD_DH | hd.BWD_LAYER1)
                main.wait_event(w2_done)
                head_ready = None
                if self.sync.enabled:
                    # 99 % of the gradient bytes (the noisy head) are final here: start their all-reduce on a side
                    # stream so it overlaps the conv backward; the conv slice (a few hundred KB) follows afterwards
                    head_ready = torch.cuda.Event()
                    head_ready.record(main)
                    with torch.cuda.stream(s_tg):
                        s_tg.wait_event(head_ready)
                        self.sync.all_reduce_(self.optimiser.flat_grad[self.optimiser.conv_end:])
                        head_reduced = torch.cuda.Event()
                        head_reduced.record(s_tg)
                grads_done = on.conv_backward_into_grads(acts, dx.view_as(acts[-1]), s_ns)
                main.wait_event(grads_done)
                if self.sync.enabled:
                    self.sync.all_reduce_(self.optimiser.flat_grad[:self.optimiser.conv_end])
                    main.wait_event(head_reduced)
            else:
                self.optimiser.zero_conv_grad()
                on.head().backward(p_on, xs_d, h_on[:B], dz, dh, dx)       # writes the 16 head gradients + dx
        if not manual:
            x_s.backward(dx)
            self.sync.all_reduce_(self.optimiser.flat_grad)
        self.optimiser.step(grad_scale=1.0 / self.sync.world_size)
        if wb_done is not None:
            main.wait_event(wb_done)
        return loss

    def _update_from_batch(self, batch, target_noise=None, after_loss=None):
        """agent.py:66-98 on an already sampled batch; returns per-sample losses (device).  `after_loss(loss)`, if
        given, is called as soon as the losses exist (the fused path runs it on a side stream)."""
        if self._fused_path(batch[1].shape[0]):
            return self._update_fused(batch, target_noise, after_loss)
        idxs, states, actions, returns, next_states, nonterminals, weights = batch
        q_s = self.online_net.logits(states)
        with torch.no_grad():
            q_ns = self.online_net.logits(next_states)
            if target_noise is None:
                self.target_net.reset_noise()
            else:
                self.target_net.reset_noise(*target_noise)
            q_t = self.target_net.logits(next_states)
            loss, grad = c51_loss_grad(q_s.detach(), q_ns, q_t, actions, returns, nonterminals, weights, self.support,
                                       self.Vmin, self.Vmax, self.delta_z, self.discount ** self.n)
        self.optimiser.zero_grad()
        q_s.backward(grad)
        self.sync.all_reduce_(self.optimiser.flat_grad)
        self.optimiser.step(grad_scale=1.0 / self.sync.world_size)
        if after_loss is not None:
            after_loss(loss)
        return loss

    def _learn_eager(self, mem):
        batch = mem.sample(self.batch_size)
        if isinstance(mem, ReplayMemory):
            return self._update_from_batch(batch, after_loss=lambda loss: mem.update_priorities(batch[0], loss))
        loss = self._update_from_batch(batch)
        mem.update_priorities(batch[0], loss.detach().cpu().numpy())  # a foreign (reference-style, host) memory: agent.py:100
        return loss

    def _capture(self, mem):
        """Record one whole update (sample -> ... -> priority write-back) into a CUDA graph.  Capturing does
        not execute; the caller replays."""
        ws = _SampleWorkspace(self.batch_size, mem.history, self.device)
        mem.flush_appends()
        mem.push_beta()  # outside the capture: a captured fill_ would freeze beta at today's value
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            batch = mem.sample_into(ws)
            loss = self._update_from_batch(batch, after_loss=lambda l: mem.update_priorities(batch[0], l))
        self._graph, self._ws, self.last_loss = graph, ws, loss

    GRAPH_WARMUP = 2  # eager updates before capture (cuDNN/cuBLAS plan selection, autograd buffers)

    def learn(self, mem):
        """agent.py:61-100.  Exactly one update per call: the first GRAPH_WARMUP calls run eagerly on a side
        stream (torch's documented warm-up recipe for whole-step capture), the next call captures the graph and
        every call from then on is one graph launch."""
        # the captured graph bakes in: this memory's buffers, the batch size and training-mode (noisy) weights
        graphable = (self.use_cuda_graph and isinstance(mem, ReplayMemory) and mem.rng == "philox" and
                     self.online_net.training)
        key = (mem, self.batch_size)   # holds a reference: a recycled id() can never alias a dead memory's graph
        if graphable and (self._graph_key is None or self._graph_key[0] is not mem or self._graph_key[1] != self.batch_size):
            self._graph, self._graph_key, self._warm = None, key, 0
        if not graphable:
            self.last_loss = self._learn_eager(mem)
        elif self._graph is None and self._warm < self.GRAPH_WARMUP:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self.last_loss = self._learn_eager(mem)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._warm += 1
        else:
            if self._graph is None:
                self._capture(mem)
            mem.flush_appends()   # no-op unless the memory defers its appends
            mem.push_beta()
            self._graph.replay()
        self._learn_calls += 1
        if self._learn_calls % 4096 == 0 and isinstance(mem, ReplayMemory):
            mem.check_last_sample()
