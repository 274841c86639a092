"""CPU port of the reference learner step -- TEST / BASELINE INFRASTRUCTURE ONLY.

`OracleLearner` restates one reference update (main.py:150-151,163-164: dqn.reset_noise(); dqn.learn(mem),
agent.py:61-100) on the host: the replay side runs through oracle/rb_oracle.c (sum tree, stratified
sampling with the legacy numpy generator, window gather, IS weights, priority write-back), the network
side is eager torch-CPU like the reference's (conv2d / linear on composed noisy weights, autograd
backward, clip_grad_norm_, Adam; all intra-op threads) with the C51 target/loss/gradient coming from the
C oracle instead of the reference's ~25 small ATen ops.  The reference itself is pure Python and cannot travel to the GPU box, so this port is
what bench.py times as `cpu_baseline` / `--impl reference` ("kind": "port").  Its tree code is C, i.e.
FASTER than the reference's numpy tree (which is ~5 % of the reference's CPU step), so the
baseline it gives is, if anything, flattering to the CPU.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc

_ARCH = {"canonical": (((32, 8, 4), (64, 4, 2), (64, 3, 1)), 3136), "data-efficient": (((32, 5, 5), (64, 5, 5)), 576)}


class _Net:
    """Parameters of one dueling noisy network as plain tensors (reference model.py:49-80 layout)."""

    def __init__(self, arch, history, hidden, atoms, actions, noisy_std, requires_grad):
        specs, flat = _ARCH[arch]
        self.atoms, self.actions, self.flat = atoms, actions, flat
        self.conv, c_in = [], history
        for c_out, k, s in specs:
            conv = torch.nn.Conv2d(c_in, c_out, k, stride=s)
            self.conv.append((conv.weight.detach().clone(), conv.bias.detach().clone(), s))
            c_in = c_out
        self.fc = {}
        for name, fin, fout in (("h_v", flat, hidden), ("h_a", flat, hidden), ("z_v", hidden, atoms),
                                ("z_a", hidden, actions * atoms)):
            r = 1 / math.sqrt(fin)
            self.fc[name] = dict(w_mu=torch.empty(fout, fin).uniform_(-r, r), w_sig=torch.full((fout, fin), noisy_std / math.sqrt(fin)),
                                 b_mu=torch.empty(fout).uniform_(-r, r), b_sig=torch.full((fout,), noisy_std / math.sqrt(fout)),
                                 w_eps=torch.zeros(fout, fin), b_eps=torch.zeros(fout))
        for t in self.parameters():
            t.requires_grad_(requires_grad)
        self.reset_noise()

    def parameters(self):
        out = []
        for w, b, _ in self.conv:
            out += [w, b]
        for name in ("h_v", "h_a", "z_v", "z_a"):
            d = self.fc[name]
            out += [d["w_mu"], d["w_sig"], d["b_mu"], d["b_sig"]]
        return out

    def copy_from(self, other):
        with torch.no_grad():
            for a, b in zip(self.parameters(), other.parameters()):
                a.copy_(b)
            for name in self.fc:
                self.fc[name]["w_eps"].copy_(other.fc[name]["w_eps"])
                self.fc[name]["b_eps"].copy_(other.fc[name]["b_eps"])

    def reset_noise(self):  # model.py:32-40,82-85
        for name in ("h_v", "h_a", "z_v", "z_a"):
            d = self.fc[name]
            x_in, x_out = torch.randn(d["w_mu"].shape[1]), torch.randn(d["w_mu"].shape[0])
            e_in, e_out = x_in.sign().mul_(x_in.abs().sqrt_()), x_out.sign().mul_(x_out.abs().sqrt_())
            d["w_eps"].copy_(e_out.ger(e_in))
            d["b_eps"].copy_(e_out)

    def _noisy(self, name, x):  # model.py:42-44
        d = self.fc[name]
        return F.linear(x, d["w_mu"] + d["w_sig"] * d["w_eps"], d["b_mu"] + d["b_sig"] * d["b_eps"])

    def logits(self, x):  # model.py:69-75 (pre-softmax dueling combination)
        for w, b, s in self.conv:
            x = F.relu(F.conv2d(x, w, b, stride=s))
        x = x.view(-1, self.flat)
        v = self._noisy("z_v", F.relu(self._noisy("h_v", x))).view(-1, 1, self.atoms)
        a = self._noisy("z_a", F.relu(self._noisy("h_a", x))).view(-1, self.actions, self.atoms)
        return v + a - a.mean(1, keepdim=True)


class OracleReplay:
    """Host replay on the C oracle (reference memory.py:91-159)."""

    def __init__(self, capacity, history, n, discount, beta, omega):
        self.cap, self.history, self.n, self.beta, self.omega = capacity, history, n, beta, omega
        self.tree = orc.OracleTree(capacity)
        self.gamma = np.array([discount ** i for i in range(n)], np.float32)
        self.t = 0

    def append(self, state_f32, action, reward, terminal):  # memory.py:105-108
        self.tree.append(self.t, orc.quantise_frame(state_f32[-1]), action, np.float32(reward), not terminal)
        self.t = 0 if terminal else self.t + 1

    def sample(self, B):  # memory.py:124-155
        t = self.tree
        total = t.total()
        while True:
            u = np.random.random_sample(B)  # the unit uniforms np.random.uniform(0, seg, [B]) consumes (memory.py:129)
            vals = orc.segment_samples(total, B, u)
            probs, didx, tidx = t.find(vals)
            if orc.batch_valid(didx, probs, t.index, self.cap, self.n, self.history):
                break
        states, actions, returns, nstates, nonterm = orc.gather(t, didx, self.history, self.n, self.gamma)
        w = orc.is_weights(probs, total, self.cap if t.full else t.index, self.beta)
        return (tidx, torch.from_numpy(states), torch.from_numpy(actions), torch.from_numpy(returns),
                torch.from_numpy(nstates), torch.from_numpy(nonterm), torch.from_numpy(w))

    def update_priorities(self, tidx, raw):  # memory.py:157-159
        self.tree.update(tidx, orc.pow_priorities(raw, self.omega))


class OracleLearner:
    """reset_noise() + learn(mem): agent.py:49-50,61-100 on torch-CPU."""

    def __init__(self, args, actions):
        self.B, self.atoms, self.n, self.discount = args.batch_size, args.atoms, args.multi_step, args.discount
        self.Vmin, self.Vmax, self.norm_clip = args.V_min, args.V_max, args.norm_clip
        self.support = torch.linspace(args.V_min, args.V_max, self.atoms)
        self.delta_z = (args.V_max - args.V_min) / (self.atoms - 1)
        mk = lambda g: _Net(args.architecture, args.history_length, args.hidden_size, self.atoms, actions, args.noisy_std, g)
        self.online, self.target = mk(True), mk(False)
        self.target.copy_from(self.online)
        self.opt = torch.optim.Adam(self.online.parameters(), lr=args.learning_rate, eps=args.adam_eps)
        self.last_loss = None

    def reset_noise(self):
        self.online.reset_noise()

    def learn(self, mem):
        """Three network passes in torch (autograd for the online pass), the distributional target, loss and
        d loss / d logits from the pinned C oracle (orc.c51 restates agent.py:67-96), then backward from the
        logits, clip and Adam."""
        idxs, states, actions, returns, next_states, nonterminals, weights = mem.sample(self.B)
        q_s = self.online.logits(states)
        with torch.no_grad():
            q_ns = self.online.logits(next_states)
            self.target.reset_noise()  # agent.py:74
            q_t = self.target.logits(next_states)
        loss, grad, _, _ = orc.c51(q_s.detach().numpy(), q_ns.numpy(), q_t.numpy(), actions.numpy(), returns.numpy(),
                                   nonterminals.numpy(), weights.numpy(), self.support.numpy(), self.Vmin, self.Vmax,
                                   self.delta_z, self.discount ** self.n)
        self.opt.zero_grad()
        q_s.backward(torch.from_numpy(grad))
        torch.nn.utils.clip_grad_norm_(self.online.parameters(), self.norm_clip)  # agent.py:97
        self.opt.step()  # agent.py:98
        self.last_loss = loss
        mem.update_priorities(idxs, loss)  # agent.py:100
