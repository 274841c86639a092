"""Drop-in module: `PYTHONPATH=/path/to/rainbow-b200/dropin:/path/to/rainbow-b200 python main.py ...` makes the
reference's unmodified main.py / test.py import the B200-native classes under the reference's module name."""
from rainbow_b200.model import DQN, NoisyLinear  # noqa: F401
