"""Which first-layer critic weights deviate from the oracle after each of three c2 updates (fp32 mode, both engines)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402
from d3rlpy_b200.algos import CQL  # noqa: E402
from oracle import update as ou  # noqa: E402

O, A, B, N, H = 17, 6, 256, 10, [256, 256, 256]
torch.set_num_threads(8)
for eng, ops in ((1, 7), (1, 3), (1, 4), (1, 1), (1, 2), (0, 7)):
    lib().set_fp32_engine(eng)
    lib().tc32_set_ops(ops)
    orc = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=5)
    algo = CQL(actor_encoder_factory=H, critic_encoder_factory=H, n_action_samples=N)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi), (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    rs = np.random.RandomState(0)
    key = "_q_funcs.0._encoder._fcs.0.weight"
    for s in range(3):
        arrays = dict(observations=rs.randn(B, O).astype(np.float32), actions=rs.uniform(-1, 1, (B, A)).astype(np.float32),
                      rewards=rs.randn(B, 1).astype(np.float32), next_observations=rs.randn(B, O).astype(np.float32),
                      terminals=(rs.rand(B, 1) < 0.05).astype(np.float32), n_steps=np.ones((B, 1), np.float32))
        noise = ou.Noise(seed=100 + s)
        orc.update(ou.Batch(arrays), noise)
        impl.inject_noise(noise.log, B)
        algo.update(SimpleNamespace(**arrays))
        g = impl.q_function.state_dict()[key].cpu()
        r = orc.q[key].detach()
        d = (g - r).abs()
        idx = (d > 2e-5).nonzero()
        m = impl._q_func.arena.state_dict("exp_avg")[key].cpu()
        mr = orc.critic_optim.state[orc.q[key]]["exp_avg"]
        vr = orc.critic_optim.state[orc.q[key]]["exp_avg_sq"]
        errs = []
        msd = impl._q_func.arena.state_dict("exp_avg")
        for k2 in orc.q:
            mr2 = orc.critic_optim.state[orc.q[k2]]["exp_avg"]
            errs.append(f"{k2.replace('_q_funcs.', 'q').replace('_encoder._fcs.', 'L').replace('weight', 'w').replace('bias', 'b')}={float((msd[k2].cpu() - mr2).norm() / mr2.norm()):.1e}")
        print(f"engine {'tc32' if eng else 'simt'} ops {ops} step {s}: {idx.shape[0]} elements > 2e-5; max {float(d.max()):.2e}; exp_avg rel-L2: " + " ".join(errs))
lib().set_fp32_engine(1)
lib().tc32_set_ops(7)
