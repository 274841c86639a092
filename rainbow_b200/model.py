"""Dueling distributional network with factorised-noise linear layers.

API and state_dict layout follow the reference's model.py (NoisyLinear model.py:10-46, DQN
model.py:49-85) so checkpoints interchange: convs.{0,2,4}.{weight,bias} and
fc_{h_v,h_a,z_v,z_a}.{weight_mu,weight_sigma,bias_mu,bias_sigma,weight_epsilon,bias_epsilon}.

Per the north star the conv/GEMM body stays a cuDNN/cuBLAS torch forward.  What is ours:
  * reset_noise(): all four NoisyLinear layers of a net are resampled by ONE rb_noisy_resample launch
    (device Philox + Box-Muller, f(x)=sign(x)sqrt|x|, outer product streamed straight into the
    weight_epsilon buffers) instead of the reference's ~52 ATen ops (model.py:32-40, 82-85);
  * logits(): the pre-softmax dueling combination (model.py:75), which the fused C51 kernel consumes
    (the softmax / log_softmax of model.py:76-79 are folded into that kernel).
"""
import ctypes as C
import math

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib

_ARCH = {
    # name: (conv specs (out_channels, kernel, stride), flattened conv output size)   model.py:55-63
    "canonical": (((32, 8, 4), (64, 4, 2), (64, 3, 1)), 3136),
    "data-efficient": (((32, 5, 5), (64, 5, 5)), 576),
}


class NoisyLinear(nn.Module):
    """y = x (mu_w + sigma_w * eps_w)^T + (mu_b + sigma_b * eps_b) in training mode, mu only in eval mode."""

    def __init__(self, in_features, out_features, std_init=0.5):
        super().__init__()
        self.in_features, self.out_features, self.std_init = in_features, out_features, std_init
        self.weight_mu = nn.Parameter(torch.empty(out_features, in_features))
        self.weight_sigma = nn.Parameter(torch.empty(out_features, in_features))
        self.register_buffer("weight_epsilon", torch.zeros(out_features, in_features))
        self.bias_mu = nn.Parameter(torch.empty(out_features))
        self.bias_sigma = nn.Parameter(torch.empty(out_features))
        self.register_buffer("bias_epsilon", torch.zeros(out_features))
        self.reset_parameters()
        self._construction_noise()

    def _construction_noise(self):
        """Construction-time only (host, before the module is moved to the GPU): the reference draws one
        noise sample in __init__ (model.py:23).  Doing the same keeps the torch RNG stream -- and therefore
        every later layer's initial weights -- identical to the reference for the same seed.  The learner
        never uses this path: DQN.reset_noise() is the CUDA kernel."""
        with torch.no_grad():
            x_in, x_out = torch.randn(self.in_features), torch.randn(self.out_features)
            f_in, f_out = x_in.sign() * x_in.abs().sqrt(), x_out.sign() * x_out.abs().sqrt()
            self.weight_epsilon.copy_(torch.outer(f_out, f_in))
            self.bias_epsilon.copy_(f_out)

    def reset_parameters(self):  # model.py:25-30
        bound = 1.0 / math.sqrt(self.in_features)
        with torch.no_grad():
            self.weight_mu.uniform_(-bound, bound)
            self.bias_mu.uniform_(-bound, bound)
            self.weight_sigma.fill_(self.std_init / math.sqrt(self.in_features))
            self.bias_sigma.fill_(self.std_init / math.sqrt(self.out_features))

    def forward(self, x):
        if self.training:  # model.py:43-44
            w = torch.addcmul(self.weight_mu, self.weight_sigma, self.weight_epsilon)
            b = torch.addcmul(self.bias_mu, self.bias_sigma, self.bias_epsilon)
            return F.linear(x, w, b)
        return F.linear(x, self.weight_mu, self.bias_mu)


def resample_noise(layers, seed, rng_counter, x_in=None, x_out=None):
    """One rb_noisy_resample launch over `layers` (NoisyLinear modules on one CUDA device).
    x_in / x_out: optional injected raw normals (parity mode), concatenated over layers."""
    lib = _lib.load()
    n = len(layers)
    w = (C.c_void_p * n)(*[_lib.ptr(m.weight_epsilon) for m in layers])
    b = (C.c_void_p * n)(*[_lib.ptr(m.bias_epsilon) for m in layers])
    fin = (C.c_int * n)(*[m.in_features for m in layers])
    fout = (C.c_int * n)(*[m.out_features for m in layers])
    _lib.check(lib.rb_noisy_resample(w, b, fin, fout, n, _lib.ptr(x_in), _lib.ptr(x_out), seed,
                                     _lib.ptr(rng_counter), _lib.stream()))


class DQN(nn.Module):
    def __init__(self, args, action_space):
        super().__init__()
        self.atoms = args.atoms
        self.action_space = action_space
        if args.architecture not in _ARCH:
            raise ValueError(f"unknown architecture '{args.architecture}'")
        specs, self.conv_output_size = _ARCH[args.architecture]
        mods, c_in = [], args.history_length
        for c_out, k, s in specs:
            mods += [nn.Conv2d(c_in, c_out, k, stride=s, padding=0), nn.ReLU()]
            c_in = c_out
        self.convs = nn.Sequential(*mods)
        self.fc_h_v = NoisyLinear(self.conv_output_size, args.hidden_size, std_init=args.noisy_std)
        self.fc_h_a = NoisyLinear(self.conv_output_size, args.hidden_size, std_init=args.noisy_std)
        self.fc_z_v = NoisyLinear(args.hidden_size, self.atoms, std_init=args.noisy_std)
        self.fc_z_a = NoisyLinear(args.hidden_size, action_space * self.atoms, std_init=args.noisy_std)
        self.noise_seed = int(torch.initial_seed()) & (2 ** 63 - 1)
        self.register_buffer("_noise_counter", torch.zeros(1, dtype=torch.int64), persistent=False)

    def noisy_layers(self):
        """Layers in the reference's reset order (named_children containing 'fc', model.py:83-85)."""
        return [m for name, m in self.named_children() if "fc" in name]

    def logits(self, x):
        """Pre-softmax q [B, A, Z] (model.py:69-75)."""
        x = self.convs(x).view(-1, self.conv_output_size)
        v = self.fc_z_v(F.relu(self.fc_h_v(x))).view(-1, 1, self.atoms)
        a = self.fc_z_a(F.relu(self.fc_h_a(x))).view(-1, self.action_space, self.atoms)
        return v + a - a.mean(1, keepdim=True)

    def forward(self, x, log=False):
        q = self.logits(x)
        return F.log_softmax(q, dim=2) if log else F.softmax(q, dim=2)

    def reset_noise(self, x_in=None, x_out=None):
        """model.py:82-85: one kernel launch for all layers.  Needs the network on a CUDA device."""
        layers = self.noisy_layers()
        if not layers[0].weight_epsilon.is_cuda:
            raise _lib.RainbowB200Error("DQN.reset_noise needs the network on a CUDA device (no CPU fallback)")
        resample_noise(layers, self.noise_seed, self._noise_counter, x_in, x_out)
