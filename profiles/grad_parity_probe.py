"""Per-tensor gradient parity of the first CQL update (c2 shape) vs the fp32 oracle, for both arithmetic
modes: relative L2 error and cosine of Adam's exp_avg (= 0.1 * grad after step 1)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200.algos import CQL  # noqa: E402
from oracle import update as ou  # noqa: E402

O, A, B, N, H = 17, 6, 256, 10, [256, 256, 256]
rs = np.random.RandomState(0)
arrays = dict(observations=rs.randn(B, O).astype(np.float32), actions=rs.uniform(-1, 1, (B, A)).astype(np.float32),
              rewards=rs.randn(B, 1).astype(np.float32), next_observations=rs.randn(B, O).astype(np.float32),
              terminals=(rs.rand(B, 1) < 0.05).astype(np.float32), n_steps=np.ones((B, 1), np.float32))
for precision in ("fp32", "bf16"):
    orc = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=5)
    algo = CQL(actor_encoder_factory=H, critic_encoder_factory=H, n_action_samples=N, precision=precision)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    noise = ou.Noise(seed=100)
    ref = orc.update(ou.Batch(arrays), noise)
    impl.inject_noise(noise.log, B)
    got = algo.update(SimpleNamespace(**arrays))
    print(precision, {k: (float(got[k]), ref[k]) for k in ref})
    for name, net, params, opt in (("critic", impl._q_func, orc.q, orc.critic_optim),
                                   ("policy", impl._policy, orc.pi, orc.actor_optim)):
        m = net.arena.state_dict("exp_avg")
        for k, p in params.items():
            r = opt.state[p]["exp_avg"]
            g = m[k].cpu()
            rel = float((g - r).norm() / (r.norm() + 1e-30))
            cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
            print(f"  {precision} {name:6s} {k:42s} rel_l2={rel:.4f} cos={cos:.5f} |g|={float(r.norm()) * 10:.3e}")
