"""Online side of the update path: the HBM replay buffer (d3rlpy/online/buffers.py)."""
from .buffers import ReplayBuffer  # noqa: F401
