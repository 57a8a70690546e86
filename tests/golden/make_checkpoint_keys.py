"""Records the structure (keys, nested state_dict keys, tensor shapes) of the UNMODIFIED reference's
`impl.save_model()` checkpoints (d3rlpy/algos/torch/base.py:137-142, torch_utility.py:97-110) for the algorithms on
the path.  Run in the build container:  python tests/golden/make_checkpoint_keys.py"""
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_import  # noqa: E402

ref_import.load()
from d3rlpy.algos import AWAC, BEAR, CRR, DDPG, DQN, IQL, BCQ, CQL, PLAS, SAC, TD3, DiscreteCQL, TD3PlusBC  # noqa: E402
from d3rlpy.models.encoders import VectorEncoderFactory  # noqa: E402


def describe(obj):
    if isinstance(obj, torch.Tensor):
        return {"shape": list(obj.shape), "dtype": str(obj.dtype)}
    if isinstance(obj, dict):
        return {str(k): describe(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [describe(v) for v in obj]
    return obj if isinstance(obj, (int, float, bool, str, type(None))) else str(type(obj).__name__)


enc = VectorEncoderFactory([32, 32])
cases = {
    "cql": (CQL(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "td3bc": (TD3PlusBC(actor_encoder_factory=enc, critic_encoder_factory=enc, scaler=None), (6,), 3),
    "bcq": (BCQ(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc), (6,), 3),
    "dcql": (DiscreteCQL(encoder_factory=enc, n_critics=2), (6,), 4),
    "sac": (SAC(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "td3": (TD3(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "ddpg": (DDPG(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "iql": (IQL(actor_encoder_factory=enc, critic_encoder_factory=enc, value_encoder_factory=enc), (6,), 3),
    "dqn_qr": (DQN(encoder_factory=enc, q_func_factory="qr"), (6,), 4),
    "awac": (AWAC(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "crr": (CRR(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "plas": (PLAS(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc), (6,), 3),
    "bear": (BEAR(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc), (6,), 3),
}
out = {}
for name, (algo, obs, act) in cases.items():
    algo.create_impl(obs, act)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "m.pt")
        algo.impl.save_model(f)
        ck = torch.load(f, map_location="cpu", weights_only=False)
    out[name] = describe(ck)
json.dump(out, open(os.path.join(HERE, "checkpoint_keys.json"), "w"), indent=1)  # key ORDER is part of the layout
print({k: sorted(v.keys()) for k, v in out.items()})
