import os, sys
from types import SimpleNamespace
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200.algos import CRR
from oracle import update as ou
from tests.golden_io import Case, load_awac
case = Case(load_awac(), "crr_binary_max_soft"); c = case.cfg; B = 16
algo = CRR(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B, beta=1.0, n_action_samples=4,
           advantage_type="max", weight_type="binary", max_weight=20.0, target_update_type="soft")
algo.create_impl((6,), 3); impl = algo.impl; impl.use_graph = False
for v, g in ((impl.q_function, "q"), (impl.targ_q_function, "q"), (impl.policy, "pi"), (impl.targ_policy, "pi")):
    v.load_state_dict(case.group("init", g))
orc = ou.CRR(6, 3, critics=case.group("init", "q"), policy=case.group("init", "pi"), beta=1.0, n_action_samples=4,
             advantage_type="max", weight_type="binary", max_weight=20.0, target_update_type="soft", target_update_interval=100)
b = ou.Batch(case.batch(0)); noise = ou.Noise(injected=case.noise(0))
orc.critic_optim.zero_grad()
cl = ou.td_error_continuous(orc.q, b.observations, b.actions, b.rewards, orc.compute_target(b, noise), b.terminals, orc.gamma ** b.n_steps)
cl.backward(); orc.critic_optim.step()
adv = orc.compute_advantage(b, noise).view(-1)
impl.inject_noise(case.noise(0), B)
bb = SimpleNamespace(**case.batch(0))
print("critic", float(impl.update_critic(bb)), float(cl))
print("actor", float(impl.update_actor(bb)))
q = [v for k, v in impl._ws.items() if k[0] == "wq_q"][0].cpu()[0]
w = [v for k, v in impl._ws.items() if k[0] == "aw"][0].cpu()
x = [v for k, v in impl._ws.items() if k[0] == "xw"][0].cpu()
qd, qs = q[:B], q[B:].view(B, 4)
print("ours adv", (qd - qs.max(1).values))
print("ref  adv", adv)
print("weights", w)
# reference sampled actions
dist = orc._dist(orc.pi, b.observations)
print("x rows 16..20 actions", x[16:20, 6:])
