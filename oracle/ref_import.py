"""Import the UNMODIFIED reference (d3rlpy 1.1.0) from /root/reference.

TEST INFRASTRUCTURE ONLY — used by ``tests/golden/make_golden.py`` and by the
``reference``-marked differential tests that run in the build container.  The
reference tree is read-only and absent on the GPU box, so nothing on the
``-m gpu`` path, ``smoke()`` or ``bench.py`` imports this module.

Shims, none of which edits reference source:
  1. ``d3rlpy.dataset`` (Cython) is served from ``oracle/_ref/`` (built by
     ``oracle/build_ref.py``) through a meta-path finder.
  2. ``gym``, ``structlog``, ``tensorboardX``, ``GPUtil``, ``h5py`` are absent
     from the image: tiny in-memory stub modules stand in for them.
  3. ``WANDB_MODE=disabled`` (``d3rlpy/logger.py:85`` calls ``wandb.init``).
"""
import importlib.abc
import importlib.util
import os
import sys
import types

REF = os.environ.get("D3RLPY_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "d3rlpy", "dataset.pyx"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    try:
        import gym  # noqa: F401
    except ImportError:
        class _Space:
            def __init__(self, *a, **k):
                self.shape = k.get("shape", ())

        class Box(_Space):
            pass

        class Discrete(_Space):
            def __init__(self, n=0):
                self.n = n
                self.shape = ()

        class Env:
            pass

        class Wrapper(Env):
            def __init__(self, env=None):
                self.env = env

        class ObservationWrapper(Wrapper):
            pass

        class TransformReward(Wrapper):
            def __init__(self, env=None, f=None):
                super().__init__(env)

        spaces = _stub("gym.spaces", Box=Box, Discrete=Discrete)
        discrete = _stub("gym.spaces.discrete", Discrete=Discrete)
        spaces.discrete = discrete
        wrappers = _stub("gym.wrappers", TransformReward=TransformReward)
        _stub("gym", Env=Env, Wrapper=Wrapper, ObservationWrapper=ObservationWrapper,
              spaces=spaces, wrappers=wrappers, make=lambda *a, **k: None)
    try:
        import structlog  # noqa: F401
    except ImportError:
        class BoundLogger:
            def _nop(self, *a, **k):
                return None
            debug = info = warning = error = critical = _nop

        _stub("structlog", BoundLogger=BoundLogger, get_logger=lambda *a, **k: BoundLogger())
    try:
        import tensorboardX  # noqa: F401
    except ImportError:
        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def add_hparams(self, *a, **k):
                pass

            def close(self):
                pass

        _stub("tensorboardX", SummaryWriter=SummaryWriter)
    try:
        import GPUtil  # noqa: F401
    except ImportError:
        _stub("GPUtil", getGPUs=lambda: [])
    try:
        import h5py  # noqa: F401
    except ImportError:
        def _nofile(*a, **k):
            raise RuntimeError("h5py stub: HDF5 I/O is outside the hot path")

        _stub("h5py", File=_nofile)


class _DatasetFinder(importlib.abc.MetaPathFinder):
    def __init__(self, so_path: str):
        self._so = so_path

    def find_spec(self, name, path, target=None):
        if name == "d3rlpy.dataset":
            return importlib.util.spec_from_file_location(name, self._so)
        return None


def load():
    """Returns the reference ``d3rlpy`` package (imported once per process)."""
    if "d3rlpy" in sys.modules and getattr(sys.modules["d3rlpy"], "__file__", "").startswith(REF):
        return sys.modules["d3rlpy"]
    if not available():
        raise RuntimeError("reference tree %s not present" % REF)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    try:
        import build_ref
    finally:
        sys.path.pop(0)
    so = build_ref.build()
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    _install_stubs()
    sys.meta_path.insert(0, _DatasetFinder(so))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import d3rlpy  # noqa: E402

    return d3rlpy
