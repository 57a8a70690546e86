"""Host staging of a numpy minibatch (`DeviceBatch.stage_host`): the pinned buffer must hold exactly what the
reference's `_convert_to_torch` would upload (d3rlpy/torch_utility.py:146-149: float32 casts of the six arrays; uint8
frames stay uint8), for every array shape / dtype a `TransitionMiniBatch`-like object can carry."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from d3rlpy_b200.algos.torch.base import DeviceBatch


def _expected(db, b):
    """The layout restated with plain slicing: [obs | next_obs | act | rew | term | n_steps], each 4-float aligned."""
    h = np.zeros(db.nfloat, np.float32)
    B = db.B

    def put(name, arr, n):
        h[db.off[name]:db.off[name] + n] = np.asarray(arr, dtype=np.float32).reshape(-1)

    if not db.pixel_shape:
        put("obs", b.observations, B * db.O)
        put("next_obs", b.next_observations, B * db.O)
    put("act", b.actions, B * (1 if db.discrete else db.A))
    put("rew", b.rewards, B)
    put("term", b.terminals, B)
    put("nsteps", b.n_steps, B)
    return h


def _vector_batch(rs, B, O, A, dtype=np.float32, flat_scalars=False):
    s = (B,) if flat_scalars else (B, 1)
    return SimpleNamespace(observations=rs.randn(B, O).astype(dtype), next_observations=rs.randn(B, O).astype(dtype),
                           actions=rs.uniform(-1, 1, (B, A)).astype(dtype), rewards=rs.randn(*s).astype(dtype),
                           terminals=(rs.rand(*s) < 0.1).astype(dtype), n_steps=rs.randint(1, 4, s).astype(dtype))


@pytest.mark.parametrize("B,O,A", [(256, 17, 6), (1, 11, 3), (33, 5, 1), (0, 4, 2)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("flat_scalars", [False, True])
def test_vector_batch_staging(B, O, A, dtype, flat_scalars):
    rs = np.random.RandomState(B + O)
    db = DeviceBatch(B, O, A, torch.device("cpu"))
    b = _vector_batch(rs, B, O, A, dtype, flat_scalars)
    db.stage_host(b)
    np.testing.assert_array_equal(db.host_np, _expected(db, b))
    # a second batch overwrites every field (no stale rows)
    b2 = _vector_batch(rs, B, O, A, dtype, flat_scalars)
    db.stage_host(b2)
    np.testing.assert_array_equal(db.host_np, _expected(db, b2))


def test_non_contiguous_and_list_inputs():
    rs = np.random.RandomState(3)
    B, O, A = 16, 6, 2
    db = DeviceBatch(B, O, A, torch.device("cpu"))
    big = rs.randn(B, 2 * O).astype(np.float32)
    b = _vector_batch(rs, B, O, A)
    b.observations = big[:, ::2]                       # strided view
    b.next_observations = np.asfortranarray(b.next_observations)
    b.rewards = [float(x) for x in rs.randn(B)]        # a plain list
    b.actions = torch.from_numpy(b.actions)            # a CPU tensor
    db.stage_host(b)
    np.testing.assert_array_equal(db.host_np, _expected(db, b))


def test_discrete_actions_are_cast_to_float():
    rs = np.random.RandomState(4)
    B, O = 32, 8
    db = DeviceBatch(B, O, 4, torch.device("cpu"), discrete=True)
    for shape in ((B,), (B, 1)):
        b = _vector_batch(rs, B, O, 1)
        b.actions = rs.randint(0, 4, shape).astype(np.int32)
        db.stage_host(b)
        np.testing.assert_array_equal(db.host_np, _expected(db, b))
        assert db.host_np[db.off["act"]:db.off["act"] + B].tolist() == b.actions.reshape(-1).astype(float).tolist()


def test_pixel_batch_staging_keeps_uint8():
    rs = np.random.RandomState(5)
    B, shape = 8, (4, 12, 12)
    db = DeviceBatch(B, 0, 4, torch.device("cpu"), pixel_shape=shape, discrete=True)
    b = SimpleNamespace(observations=rs.randint(0, 256, (B, *shape)).astype(np.uint8),
                        next_observations=rs.randint(0, 256, (B, *shape)).astype(np.uint8),
                        actions=rs.randint(0, 4, B).astype(np.int32), rewards=rs.randn(B, 1).astype(np.float32),
                        terminals=np.zeros((B, 1), np.float32), n_steps=np.ones((B, 1), np.float32))
    db.stage_host(b)
    assert db.pix_host_np.dtype == np.uint8
    np.testing.assert_array_equal(db.pix_host_np[:db.npix], b.observations.reshape(-1))
    np.testing.assert_array_equal(db.pix_host_np[db.npix:], b.next_observations.reshape(-1))
    np.testing.assert_array_equal(db.host_np, _expected(db, b))
    # frames handed over flattened per row are accepted too
    b.observations = b.observations.reshape(B, -1)
    db.stage_host(b)
    np.testing.assert_array_equal(db.pix_host_np[:db.npix], b.observations.reshape(-1))


def test_wrong_batch_size_is_an_error():
    rs = np.random.RandomState(6)
    db = DeviceBatch(16, 4, 2, torch.device("cpu"))
    with pytest.raises(ValueError):
        db.stage_host(_vector_batch(rs, 8, 4, 2))
