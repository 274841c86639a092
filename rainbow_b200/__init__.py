"""rainbow_b200 -- B200-native Rainbow learner hot path (prioritised replay in HBM, fused C51 loss,
on-device noise resampling, fused clip+Adam) behind the Agent / ReplayMemory / DQN API of Kaixhin/Rainbow.

Importing the package does not touch the GPU; the CUDA extension (librainbow_b200.so) is loaded on first
use and there is no CPU fallback for any kernel."""
from ._lib import RainbowB200Error  # noqa: F401

__all__ = ["Agent", "ReplayMemory", "SegmentTree", "DQN", "NoisyLinear", "RainbowB200Error"]


def __getattr__(name):
    if name in ("ReplayMemory", "SegmentTree"):
        from . import memory
        return getattr(memory, name)
    if name in ("DQN", "NoisyLinear"):
        from . import model
        return getattr(model, name)
    if name == "Agent":
        from . import agent
        return agent.Agent
    raise AttributeError(name)
