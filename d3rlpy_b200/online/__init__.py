"""Online side of the update path: the HBM replay buffer (d3rlpy/online/buffers.py), explorers and the
`train_single_env` loop (d3rlpy/online/iterators.py) that feeds `algo.update` from it."""
from .buffers import ReplayBuffer  # noqa: F401
from .explorers import ConstantEpsilonGreedy, LinearDecayEpsilonGreedy, NormalNoise  # noqa: F401
from .iterators import train_single_env  # noqa: F401
