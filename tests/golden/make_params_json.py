"""Records the `params.json` documents the UNMODIFIED reference writes (`LearnableBase.save_params`,
d3rlpy/base.py:823-850, through `_serialize_params` :78-98) for the algorithms on the path, so that our
`save_params` / `from_json` can be pinned to the same format.  Run in the build container:
    python tests/golden/make_params_json.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_import  # noqa: E402

ref_import.load()
from d3rlpy.algos import AWAC, BEAR, CRR, DDPG, DQN, IQL, NFQ, BCQ, CQL, PLAS, SAC, TD3, DiscreteCQL, TD3PlusBC  # noqa: E402
from d3rlpy.models.encoders import VectorEncoderFactory  # noqa: E402
from d3rlpy.models.q_functions import QRQFunctionFactory  # noqa: E402


class _Logger:
    def add_params(self, params):
        self.params = params


enc = VectorEncoderFactory([32, 32])
cases = {
    "cql": (CQL(actor_encoder_factory=enc, critic_encoder_factory=enc, n_action_samples=4), (6,), 3),
    "td3bc": (TD3PlusBC(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "bcq": (BCQ(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc), (6,), 3),
    "dcql": (DiscreteCQL(encoder_factory=enc, n_critics=2), (6,), 4),
    "dcql_pixel": (DiscreteCQL(n_frames=4, scaler="pixel"), (4, 84, 84), 4),
    "dcql_qr": (DiscreteCQL(encoder_factory=enc, q_func_factory=QRQFunctionFactory(n_quantiles=16)), (6,), 4),
    "sac": (SAC(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "td3": (TD3(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "ddpg": (DDPG(actor_encoder_factory=enc, critic_encoder_factory=enc), (6,), 3),
    "iql": (IQL(actor_encoder_factory=enc, critic_encoder_factory=enc, value_encoder_factory=enc), (6,), 3),
    "dqn_qr": (DQN(encoder_factory=enc, q_func_factory="qr"), (6,), 4),
    "nfq": (NFQ(encoder_factory=enc), (6,), 4),
    "awac": (AWAC(actor_encoder_factory=enc, critic_encoder_factory=enc, n_action_samples=2), (6,), 3),
    "crr": (CRR(actor_encoder_factory=enc, critic_encoder_factory=enc, advantage_type="max", weight_type="binary"), (6,), 3),
    "plas": (PLAS(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc, lam=0.6), (6,), 3),
    "bear": (BEAR(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc,
                  mmd_kernel="gaussian", n_mmd_action_samples=3), (6,), 3),
}
out = {}
for name, (algo, obs, act) in cases.items():
    algo.create_impl(obs, act)
    lg = _Logger()
    algo.save_params(lg)
    out[name] = json.loads(json.dumps(lg.params, default=lambda o: f"<{type(o).__name__}>"))
json.dump(out, open(os.path.join(HERE, "params_json.json"), "w"), indent=1)
print(json.dumps(out["cql"], indent=1))
print(json.dumps(out["dcql_pixel"], indent=1))
