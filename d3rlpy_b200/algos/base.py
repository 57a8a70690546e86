"""User-facing algorithm base: mirrors LearnableBase/AlgoBase (d3rlpy/base.py:124-758,
d3rlpy/algos/base.py) for the part of the API on the update path: constructor kwargs,
`create_impl`, `build_with_dataset`, `update(batch)` with the pre-increment `grad_step`
schedule (base.py:746-758), and a minimal `fit()` loop over a device-resident replay."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

IMPL_NOT_INITIALIZED_ERROR = "The neural network parameters are not initialized. Pleaes call build_with_dataset, build_with_env, or directly call fit or fit_online method."


def _hidden_units(factory, default):
    """Accepts "default", a sequence of hidden sizes, or an object with `hidden_units`
    (VectorEncoderFactory, d3rlpy/models/encoders.py:170-260).  BN / dropout / dense encoders are
    outside the hot path and rejected here, before any kernel runs."""
    if factory is None or factory == "default":
        return list(default)
    if isinstance(factory, (list, tuple)):
        return list(factory)
    hu = getattr(factory, "hidden_units", None) or getattr(factory, "_hidden_units", None)
    if hu is None:
        raise ValueError(f"unsupported encoder factory {factory!r}")
    for flag in ("use_batch_norm", "_use_batch_norm", "use_dense", "_use_dense"):
        if getattr(factory, flag, False):
            raise ValueError("batch-norm / dense encoders are not supported by the B200 path")
    if getattr(factory, "dropout_rate", None) or getattr(factory, "_dropout_rate", None):
        raise ValueError("dropout encoders are not supported by the B200 path")
    return list(hu)


class VectorEncoderFactory:
    """Same constructor as d3rlpy.models.encoders.VectorEncoderFactory (hidden_units only)."""

    TYPE = "vector"

    def __init__(self, hidden_units: Optional[Sequence[int]] = None, activation: str = "relu",
                 use_batch_norm: bool = False, dropout_rate: Optional[float] = None, use_dense: bool = False):
        if activation != "relu" or use_batch_norm or dropout_rate is not None or use_dense:
            raise ValueError("B200 path supports ReLU MLP encoders without BN/dropout/dense")
        self.hidden_units = list(hidden_units) if hidden_units is not None else [256, 256]


_ADAM_DEFAULTS = {"optim_cls": "Adam", "betas": [0.9, 0.999], "eps": 1e-08, "weight_decay": 0, "amsgrad": False}


class _AdamWeightDecay:
    """Stand-in for the reference's AdamFactory(weight_decay=...) when an algorithm is rebuilt from params.json."""

    def __init__(self, weight_decay: float):
        self.weight_decay = weight_decay

    def get_params(self, deep: bool = False):
        return dict(_ADAM_DEFAULTS, weight_decay=self.weight_decay)


def _q_func_to_json(factory) -> Dict[str, Any]:
    """QFunctionFactory -> {"type", "params"} (models/q_functions.py:60-77,117-121,156-161)."""
    if factory is None or factory == "mean" or getattr(factory, "TYPE", None) == "mean":
        return {"type": "mean", "params": {"share_encoder": False}}
    n = 32 if factory == "qr" else int(factory.n_quantiles)
    return {"type": "qr", "params": {"share_encoder": False, "n_quantiles": n}}


def _q_func_from_json(value, cls):
    kind = value["type"] if isinstance(value, dict) else value
    params = value.get("params", {}) if isinstance(value, dict) else {}
    if params.get("share_encoder", False):
        raise ValueError("share_encoder is not on the accelerated path")
    if kind == "mean":
        return "mean"
    if kind == "qr" and getattr(cls, "SUPPORTS_QR", False):
        from .dqn import QRQFunctionFactory

        return QRQFunctionFactory(n_quantiles=int(params.get("n_quantiles", 32)))
    raise ValueError(f"q_func_factory {kind!r} is not on the accelerated path of {cls.__name__}")


def _encoder_to_json(factory) -> Dict[str, Any]:
    """EncoderFactory -> {"type", "params"} as `_serialize_params` writes it (d3rlpy/base.py:78-98 with
    models/encoders.py get_type/get_params)."""
    if factory is None or factory == "default":
        return {"type": "default", "params": {"activation": "relu", "use_batch_norm": False, "dropout_rate": None}}
    if hasattr(factory, "filters"):
        return {"type": "pixel", "params": {"filters": [list(f) for f in factory.filters],
                                            "feature_size": factory.feature_size, "activation": "relu",
                                            "use_batch_norm": False, "dropout_rate": None}}
    hidden = list(factory) if isinstance(factory, (list, tuple)) else _hidden_units(factory, [256, 256])
    return {"type": "vector", "params": {"hidden_units": hidden, "activation": "relu", "use_batch_norm": False,
                                         "dropout_rate": None, "use_dense": False}}


def _encoder_from_json(doc):
    if not isinstance(doc, dict):
        return doc
    kind, params = doc["type"], dict(doc.get("params", {}))
    if params.get("activation", "relu") != "relu" or params.get("use_batch_norm") or params.get("dropout_rate") or \
            params.get("use_dense"):
        raise ValueError(f"encoder configuration outside the accelerated path: {doc}")
    if kind == "default":
        return "default"
    if kind == "vector":
        return VectorEncoderFactory(hidden_units=params.get("hidden_units"))
    if kind == "pixel":
        from .dqn import PixelEncoderFactory

        return PixelEncoderFactory(filters=[tuple(f) for f in params["filters"]] if params.get("filters") else None,
                                   feature_size=params.get("feature_size", 512))
    raise ValueError(f"unsupported encoder type {kind!r}")


def _jsonable(v):
    """default_json_encoder (d3rlpy/logger.py:21-28): numpy scalars and arrays as python numbers / nested lists."""
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, np.integer):
        return int(v)
    if isinstance(v, np.floating):
        return float(v)
    return v


def _scaler_to_json(scaler):
    """_serialize_params (base.py:78-98): {"type": get_type(), "params": get_params()}."""
    if scaler is None:
        return None
    kind = getattr(scaler, "TYPE", None)
    if kind is None or not hasattr(scaler, "get_params"):
        raise ValueError(f"unsupported scaler {scaler!r}")
    return {"type": kind, "params": {k: _jsonable(v) for k, v in scaler.get_params().items()}}


def _scaler_from_json(doc, create=None):
    """_deseriealize_params (base.py:101-117): create_scaler / create_action_scaler / create_reward_scaler."""
    if doc is None or not isinstance(doc, dict):
        return doc
    from .. import preprocessing

    create = create or preprocessing.create_scaler
    try:
        return create(doc["type"], **doc.get("params", {}))
    except AssertionError:
        raise ValueError(f"scaler {doc['type']!r} is not on the accelerated path") from None


def random_iterator_indices(rng, n_transitions: int, n_steps: int, batch_size: int) -> np.ndarray:
    """Index stream of RandomIterator for one epoch: one vectorised draw == n_steps*batch_size sequential
    `np.random.randint(n)` calls (the legacy RandomState consumes its stream element by element)."""
    return rng.randint(n_transitions, size=(n_steps, batch_size)).astype(np.int64)


def round_iterator_indices(rng, n_transitions: int, batch_size: int, shuffle: bool = True) -> np.ndarray:
    """Index stream of RoundIterator for one epoch (iterators/round_iterator.py:39-55)."""
    perm = np.arange(n_transitions)
    if shuffle:
        rng.shuffle(perm)
    n_batches = n_transitions // batch_size
    return perm[:n_batches * batch_size].reshape(n_batches, batch_size).astype(np.int64)


class AlgoBase:
    _impl = None
    DISCRETE_ACTIONS = False   # get_action_type(): ActionSpace.CONTINUOUS; the DQN family overrides it

    def __init__(self, batch_size: int, n_frames: int, n_steps: int, gamma: float, scaler=None, action_scaler=None,
                 reward_scaler=None, use_gpu=0, kwargs: Optional[Dict[str, Any]] = None):
        self._batch_size, self._n_frames, self._n_steps, self._gamma = batch_size, n_frames, n_steps, gamma
        from ..preprocessing import check_action_scaler, check_reward_scaler, check_scaler

        # check_scaler / check_action_scaler / check_reward_scaler (base.py:166-168): instance, registered name or None
        self._scaler = check_scaler(scaler)
        self._action_scaler = check_action_scaler(action_scaler)
        self._reward_scaler = check_reward_scaler(reward_scaler)
        self._use_gpu = use_gpu
        self._grad_step = 0
        self._kwargs = kwargs or {}

    # ------------------------------------------------------------------ reference API
    @property
    def batch_size(self):
        return self._batch_size

    @property
    def n_frames(self):
        return self._n_frames

    @property
    def n_steps(self):
        return self._n_steps

    @property
    def gamma(self):
        return self._gamma

    @property
    def scaler(self):
        return self._scaler

    @property
    def impl(self):
        return self._impl

    @property
    def grad_step(self) -> int:
        return self._grad_step

    def set_grad_step(self, grad_step: int) -> None:
        self._grad_step = grad_step

    def create_impl(self, observation_shape: Sequence[int], action_size: int) -> None:
        if self._impl:
            return
        self._create_impl(tuple(observation_shape), action_size)

    def _get_shape(self, observation_shape):
        """_process_observation_shape (base.py:735-744): frame stacking multiplies channels."""
        if len(observation_shape) == 3:
            return (self._n_frames * observation_shape[0],) + tuple(observation_shape[1:])
        return tuple(observation_shape)

    def build_with_dataset(self, dataset) -> None:
        self.create_impl(self._get_shape(dataset.get_observation_shape()), dataset.get_action_size())

    def build_with_env(self, env) -> None:
        """LearnableBase.build_with_env (base.py:707-720): shapes from the gym-like environment's spaces."""
        space = env.action_space
        action_size = int(space.n) if hasattr(space, "n") else int(space.shape[0])
        self.create_impl(self._get_shape(env.observation_space.shape), action_size)

    @property
    def action_size(self) -> Optional[int]:
        return self._impl.action_size if self._impl is not None else None

    @property
    def observation_shape(self):
        return self._impl.observation_shape if self._impl is not None else None

    @property
    def action_scaler(self):
        return self._action_scaler

    @property
    def reward_scaler(self):
        return self._reward_scaler

    def fit_online(self, env, buffer=None, explorer=None, n_steps: int = 1000000, n_steps_per_epoch: int = 10000,
                   update_interval: int = 1, update_start_step: int = 0, random_steps: int = 0,
                   timelimit_aware: bool = True, callback=None, **unused: Any) -> List[Dict[str, float]]:
        """AlgoBase.fit_online (algos/base.py:161-247): `train_single_env` with a ReplayBuffer of 1M transitions by
        default; logging / evaluation arguments of the reference are accepted and ignored."""
        from ..online import ReplayBuffer, train_single_env

        if buffer is None:
            buffer = ReplayBuffer(1000000, env=env)
        return train_single_env(self, env, buffer, explorer, n_steps, n_steps_per_epoch, update_interval,
                                update_start_step, random_steps, timelimit_aware, callback)

    def update(self, batch) -> Dict[str, float]:
        """LearnableBase.update (base.py:746-758): `_update` then grad_step += 1."""
        loss = self._update(batch)
        self._grad_step += 1
        return loss

    def _update(self, batch) -> Dict[str, float]:
        raise NotImplementedError

    # ------------------------------------------------------------------ params.json (base.py:188-232, 823-850)
    def get_params(self, deep: bool = True) -> Dict[str, Any]:
        """Constructor arguments by name (LearnableBase.get_params, base.py:266-319): every `_x` attribute that is a
        hyper-parameter, the factories as objects."""
        skip = {"_impl", "_grad_step", "_kwargs", "_factories", "_use_gpu", "_n_quantiles", "_actor_weight_decay"}
        out: Dict[str, Any] = {}
        for key, value in vars(self).items():
            if key in skip or not key.startswith("_") or key.endswith("_hidden"):
                continue
            out[key[1:]] = value
        out.update(getattr(self, "_factories", {}))
        out.update(self._kwargs)
        out["use_gpu"] = self._use_gpu
        return out

    def _params_document(self) -> Dict[str, Any]:
        """The document `save_params` writes: reference keys and value encodings (`_serialize_params`), plus this
        package's own constructor keys (`seed`, `precision`), which the reference's `from_json` passes through
        `**kwargs` untouched."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        doc: Dict[str, Any] = {"generated_maxlen": 100000, "real_ratio": 1.0}
        if getattr(self, "HAS_Q_FUNC_FACTORY", True):   # IQL's constructor has none (algos/iql.py:109-135)
            doc["q_func_factory"] = {"type": "mean", "params": {"share_encoder": False}}
        for key, value in self.get_params().items():
            if key.endswith("encoder_factory"):
                doc[key] = _encoder_to_json(value)
            elif key == "q_func_factory":
                doc[key] = _q_func_to_json(value)
            elif key in ("scaler", "action_scaler", "reward_scaler"):
                doc[key] = _scaler_to_json(value)
            elif key == "use_gpu":
                doc[key] = None if value is None or value is False else (0 if value is True else int(value))
            else:
                doc[key] = value
            if key.endswith("learning_rate"):
                okey = key[:-len("learning_rate")] + "optim_factory"
                doc[okey] = dict(_ADAM_DEFAULTS)
                if okey in getattr(self, "WEIGHT_DECAY_OPTIMS", ()):
                    doc[okey]["weight_decay"] = getattr(self, "_" + okey[:-len("optim_factory")] + "weight_decay")
        doc = dict(sorted(doc.items()))
        doc["algorithm"] = type(self).__name__
        doc["observation_shape"] = list(self._impl.observation_shape)
        doc["action_size"] = self._impl.action_size
        return doc

    def save_params(self, logger_or_path) -> None:
        """LearnableBase.save_params (base.py:823-850): hands the document to `logger.add_params`, or writes it to the
        given path as `params.json`."""
        doc = self._params_document()
        if hasattr(logger_or_path, "add_params"):
            logger_or_path.add_params(doc)
            return
        import json

        with open(logger_or_path, "w") as f:
            json.dump(doc, f, indent=2)

    @classmethod
    def from_json(cls, fname: str, use_gpu=0, **overrides: Any):
        """LearnableBase.from_json (base.py:188-232): rebuilds the algorithm from a `params.json` written by this
        package or by the reference, then `create_impl`.  Configurations outside the accelerated path raise."""
        import json

        with open(fname, "r") as f:
            params = json.load(f)
        observation_shape = tuple(params.pop("observation_shape"))
        action_size = params.pop("action_size")
        name = params.pop("algorithm", cls.__name__)
        if name != cls.__name__:
            raise ValueError(f"{fname} was written by {name}, not {cls.__name__}")
        for key in ("generated_maxlen", "real_ratio"):
            params.pop(key, None)
        for key in list(params):
            value = params[key]
            if key.endswith("encoder_factory"):
                params[key] = _encoder_from_json(value)
            elif key.endswith("optim_factory") and key in getattr(cls, "WEIGHT_DECAY_OPTIMS", ()):
                # AdamFactory(weight_decay=...) is part of this algorithm's definition (AWAC's actor, awac.py:105)
                wd = dict(_ADAM_DEFAULTS, weight_decay=(value or {}).get("weight_decay", 0))
                if value is not None and any(value.get(k, v) != v for k, v in wd.items() if not isinstance(v, list)):
                    raise ValueError(f"{key}: only Adam with weight decay is on the accelerated path, got {value}")
                params[key] = None if value is None else _AdamWeightDecay(float(wd["weight_decay"]))
            elif key.endswith("optim_factory"):
                if value is not None and any(value.get(k, v) != v and list(value.get(k, v)) != v
                                             for k, v in _ADAM_DEFAULTS.items() if not isinstance(v, list)) or \
                        (value is not None and list(value.get("betas", [0.9, 0.999])) != [0.9, 0.999]):
                    raise ValueError(f"{key}: only Adam defaults are on the accelerated path, got {value}")
                params[key] = None
            elif key == "q_func_factory":
                params[key] = _q_func_from_json(value, cls)
            elif key == "scaler":
                params[key] = _scaler_from_json(value)
            elif key == "action_scaler":
                from ..preprocessing import create_action_scaler

                params[key] = _scaler_from_json(value, create_action_scaler)
            elif key == "reward_scaler":
                from ..preprocessing import create_reward_scaler

                params[key] = _scaler_from_json(value, create_reward_scaler)
        params["use_gpu"] = use_gpu
        params.update(overrides)
        algo = cls(**params)
        algo.create_impl(observation_shape, action_size)
        return algo

    # ------------------------------------------------------------------ evaluation (algos/base.py:predict/predict_value)
    def predict(self, x):
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.predict_best_action(x)

    def predict_value(self, x, action, with_std: bool = False):
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.predict_value(x, action, with_std)

    def sample_action(self, x):
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.sample_action(x)

    def save_policy(self, fname: str) -> None:
        """AlgoBase.save_policy (algos/base.py): greedy policy as TorchScript (.pt) or ONNX (.onnx)."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        self._impl.save_policy(fname)

    def save_model(self, fname: str) -> None:
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        self._impl.save_model(fname)

    def load_model(self, fname: str) -> None:
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        self._impl.load_model(fname)

    # ------------------------------------------------------------------ fit over an HBM-resident replay
    def fit(self, dataset, n_epochs: Optional[int] = None, n_steps: Optional[int] = None,
            n_steps_per_epoch: int = 10000, **kwargs: Any):
        """LearnableBase.fit (base.py:349-434): `list(self.fitter(...))` -- one `(epoch, metrics)` pair per epoch."""
        return list(self.fitter(dataset, n_epochs, n_steps, n_steps_per_epoch, **kwargs))

    def fitter(self, dataset, n_epochs: Optional[int] = None, n_steps: Optional[int] = None,
               n_steps_per_epoch: int = 10000, shuffle: bool = True, seed: Optional[int] = None,
               eval_episodes=None, scorers: Optional[Dict[str, Any]] = None, callback=None, **unused: Any):
        """LearnableBase.fitter (base.py:436-687) without the logger: a generator over epochs yielding
        `(epoch, metrics)`, metrics = the epoch's mean of every loss `update` returned plus one entry per scorer
        (`scorer(algo, eval_episodes)`, base.py:760-771); `callback(algo, epoch, total_step)` runs after every step.
        `n_steps` selects the RandomIterator index stream (one `np.random.randint` per sample,
        iterators/random_iterator.py:38-41), `n_epochs` the RoundIterator one (per epoch a `np.random.shuffle`d
        permutation cut into len // batch_size consecutive batches, the remainder dropped,
        iterators/round_iterator.py:39-55) -> device gather -> update.  `seed` (ours) draws the index stream from a
        private RandomState instead of numpy's global one; the reference's logging arguments are accepted and
        ignored."""
        if (n_epochs is None) == (n_steps is None):
            raise ValueError("Either of n_epochs or n_steps must be given.")  # base.py:548-549
        subset = None
        if not hasattr(dataset, "device_replay"):   # List[Episode] / List[Transition] (base.py:494-507)
            from ..dataset import Episode, Transition
            from ..preprocessing import TransitionSubset

            if not dataset:
                raise ValueError("empty dataset is not supported.")
            if not isinstance(dataset[0], (Episode, Transition)):
                raise ValueError(f"invalid dataset type: {type(dataset)}")
            subset = TransitionSubset(dataset)
            dataset = subset._ds
        discrete = getattr(self, "DISCRETE_ACTIONS", None)   # base.py:509-519
        if discrete is not None:
            assert dataset.is_action_discrete() == discrete, \
                ("The action-space of the given dataset is not compatible with the algorithm. Please use "
                 + ("discrete" if discrete else "continuous") + " action-space algorithms.")
        for sc in (self._scaler, self._action_scaler, self._reward_scaler):   # base.py:566-585
            if sc is not None:
                sc.fit(dataset if subset is None else subset)
        self.build_with_dataset(dataset)
        replay = dataset.device_replay(self._impl._device)
        rng = np.random if seed is None else np.random.RandomState(seed)
        B = self._batch_size
        n = len(replay) if subset is None else len(subset)
        pick = (lambda idx: idx) if subset is None else (lambda idx: subset._t_index[idx])
        if n_steps is not None:
            assert n_steps >= n_steps_per_epoch  # base.py:523
            n_epochs = n_steps // n_steps_per_epoch
            draw = lambda: random_iterator_indices(rng, n, n_steps_per_epoch, B)   # noqa: E731
        else:
            draw = lambda: round_iterator_indices(rng, n, B, shuffle)              # noqa: E731
        progress = {"epoch": 0, "total_step": 0}

        def after_step():
            progress["total_step"] += 1
            callback(self, progress["epoch"], progress["total_step"])

        for epoch in range(1, n_epochs + 1):
            progress["epoch"] = epoch
            metrics = self._fit_epoch(replay, pick(draw()), after_step if callback else None)
            if scorers and eval_episodes:
                for name, scorer in scorers.items():
                    metrics[name] = scorer(self, eval_episodes)
            yield epoch, metrics

    def _fit_epoch(self, replay, idx: np.ndarray, after_step=None) -> Dict[str, float]:
        """One epoch over the index matrix [steps, batch_size]; the epoch's metric means (base.py:660-676)."""
        from ..dataset import TransitionMiniBatch

        impl, B = self._impl, self._batch_size
        if (not replay.is_image) and (not replay.discrete) and self._n_frames == 1:
            acc = self._fit_chunk_device(replay, idx, after_step)
        else:
            acc: Dict[str, List[float]] = {}
            for i in range(idx.shape[0]):
                batch = TransitionMiniBatch.from_indices(replay, idx[i], n_frames=self._n_frames,
                                                         n_steps=self._n_steps, gamma=self._gamma,
                                                         scaler=self._scaler, out=impl.device_batch(B))
                impl.scale_actions_rewards(batch._device_batch)
                batch.scaled = {"obs", "act_rew"}
                for k, v in self.update(batch).items():
                    acc.setdefault(k, []).append(float(v))
                if after_step:
                    after_step()
        return {k: float(np.mean(v)) for k, v in acc.items()}

    def _fit_chunk_device(self, replay, idx: np.ndarray, after_step=None) -> Dict[str, List[float]]:
        """Vector observations: the chunk's indices are uploaded once; every step is one gather launch + one graph
        replay + one 256-byte device-side copy of the metric slots; metrics come back with ONE D2H per chunk
        (the reference syncs on every loss)."""
        import torch
        from types import SimpleNamespace

        impl, B = self._impl, self._batch_size
        L, st = impl._lib, impl._stream
        chunk = idx.shape[0]
        idx_dev = torch.from_numpy(np.ascontiguousarray(idx)).to(impl._device)
        hist = torch.zeros(chunk, 64, dtype=torch.float32, device=impl._device)
        db = impl.device_batch(B)
        holder = SimpleNamespace(_device_batch=db, scaled={"obs", "act_rew"})
        sc = (None, None, 0.0)
        if self._scaler is not None and hasattr(self._scaler, "affine_f32"):
            m, s, e = replay.scaler_tensors(self._scaler)
            sc = (m.data_ptr(), s.data_ptr(), e)
        torch.cuda.current_stream(impl._device).synchronize()
        names_per_step = []
        O, A = replay.obs_shape[0], replay.act_dim
        for i in range(chunk):
            L.gather_vector(replay.obs.data_ptr(), O, replay.actions.data_ptr(), A, 0, replay.rewards.data_ptr(),
                            replay.meta.data_ptr(), idx_dev.data_ptr() + 8 * B * i, B, self._n_steps, float(self._gamma),
                            db.ptr("obs"), db.ptr("act"), db.ptr("rew"), db.ptr("next_obs"), db.ptr("term"),
                            db.ptr("nsteps"), sc[0], sc[1], sc[2], st)
            impl.scale_actions_rewards(db)
            names_per_step.append(self._update_async(holder))
            self._grad_step += 1
            L.copy_d2d(hist.data_ptr() + 256 * i, impl._slots.data_ptr(), 256, st)
            if after_step:
                after_step()
        impl.sync()
        h = hist.cpu().numpy()
        acc: Dict[str, List[float]] = {}
        for i, names in enumerate(names_per_step):
            for slot, key in names:
                acc.setdefault(key, []).append(float(h[i, slot]))
        return acc
