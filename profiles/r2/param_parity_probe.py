"""Post-update parameter parity of three c2 CQL updates vs the fp32 oracle, per network and engine:
max |p - p_ref|, relative L2 of the update delta (p - p0), relative L2 of Adam's moments."""
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
modes = [("fp32", 0), ("fp32", 1), ("bf16", 1)]
for precision, eng in modes:
    lib().set_fp32_engine(eng)
    orc = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=5)
    init = {"q": {k: v.detach().clone() for k, v in orc.q.items()}, "pi": {k: v.detach().clone() for k, v in orc.pi.items()}}
    algo = CQL(actor_encoder_factory=H, critic_encoder_factory=H, n_action_samples=N, precision=precision)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi), (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    rs = np.random.RandomState(0)
    for s in range(3):
        arrays = dict(observations=rs.randn(B, O).astype(np.float32), actions=rs.uniform(-1, 1, (B, A)).astype(np.float32),
                      rewards=rs.randn(B, 1).astype(np.float32), next_observations=rs.randn(B, O).astype(np.float32),
                      terminals=(rs.rand(B, 1) < 0.05).astype(np.float32), n_steps=np.ones((B, 1), np.float32))
        noise = ou.Noise(seed=100 + s)
        ref = orc.update(ou.Batch(arrays), noise)
        impl.inject_noise(noise.log, B)
        got = algo.update(SimpleNamespace(**arrays))
    print(f"== {precision} engine {'tc32' if eng else 'simt'}: metric rel err",
          {k: f"{abs(float(got[k]) - ref[k]) / max(1, abs(ref[k])):.1e}" for k in ref})
    for name, view, refp, net, opt in (("q", impl.q_function, orc.q, impl._q_func, orc.critic_optim),
                                       ("pi", impl.policy, orc.pi, impl._policy, orc.actor_optim)):
        sd = view.state_dict()
        m_sd, v_sd = net.arena.state_dict("exp_avg"), net.arena.state_dict("exp_avg_sq")
        num = den = 0.0
        mx = 0.0
        mnum = mden = vnum = vden = 0.0
        worst = None
        for k, r in refp.items():
            g = sd[k].detach().cpu()
            r = r.detach()
            d_ref, d_got = r - init[name][k], g - init[name][k]
            num += float((d_got - d_ref).pow(2).sum()); den += float(d_ref.pow(2).sum())
            e = float((g - r).abs().max())
            if e > mx:
                mx, worst = e, k
            st = opt.state[refp[k]]
            mnum += float((m_sd[k].cpu() - st["exp_avg"]).pow(2).sum()); mden += float(st["exp_avg"].pow(2).sum())
            vnum += float((v_sd[k].cpu() - st["exp_avg_sq"]).pow(2).sum()); vden += float(st["exp_avg_sq"].pow(2).sum())
        print(f"   {name:3s} max|dp| {mx:.2e} ({worst})  rel-L2(delta) {np.sqrt(num / den):.2e}  rel-L2(exp_avg) {np.sqrt(mnum / mden):.2e}"
              f"  rel-L2(exp_avg_sq) {np.sqrt(vnum / vden):.2e}")
lib().set_fp32_engine(1)
