"""d3rlpy_b200 — B200-native (sm_100a) implementation of d3rlpy's offline-RL update path.

Host side mirrors the reference's Python API for that path (algos.CQL / TD3PlusBC, impl hooks,
dataset.TransitionMiniBatch, torch_utility.soft_sync); the work is done by hand-written CUDA
kernels behind the C ABI in include/d3rlpy_b200.h.  No CPU fallback.
"""
__version__ = "0.1.0"
