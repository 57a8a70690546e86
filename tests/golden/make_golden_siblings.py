"""Golden fixtures for the sibling algorithms that reuse the update path's hooks: SAC and
TD3, recorded from the LIVE unmodified reference exactly like tests/golden/make_golden.py does for the four algorithms
on the path (same helpers: noise tap, reference-vs-oracle agreement check, case packing).

    python tests/golden/make_golden_siblings.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through oracle/ref_import)

oupdate = mg.oupdate


def main():
    from d3rlpy.algos import DDPG, IQL, SAC, TD3, TD3PlusBC
    from d3rlpy.models.encoders import VectorEncoderFactory

    out, cases = {}, []
    rs = np.random.RandomState(11)
    steps = 4  # TD3: two actor steps (grad_step 0 and 2)

    # ---- SAC: n_steps = 3 batches (gamma ** n), 3-layer encoders like c2
    O, A, B = 6, 3, 16
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(5)
    enc = VectorEncoderFactory([32, 32, 32])
    algo = SAC(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_steps=3)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.SAC(O, A, critics=init["q"], policy=init["pi"])
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=3) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "targ_pi": mg.sd(impl._targ_policy), "log_temp": mg.sd(impl._log_temp)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi), ("log_temp", orc.log_temp)):
        mg.assert_params_close(final[g], p, f"sac {g}")
    mg.pack_case("sac", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, h2=32), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("sac")

    # ---- TD3: no scaler (the reference default), delayed actor
    O, A, B = 5, 3, 16
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(6)
    enc = VectorEncoderFactory([32, 32])
    algo = TD3(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.TD3(O, A, critics=init["q"], policy=init["pi"])
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B)) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "targ_pi": mg.sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
        mg.assert_params_close(final[g], p, f"td3 {g}")
    mg.pack_case("td3", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("td3")

    # ---- DDPG: one critic, actor + soft syncs on every step, no noise draw at all
    O, A, B = 7, 2, 16
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(12)
    enc = VectorEncoderFactory([32, 32])
    algo = DDPG(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_steps=2)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.DDPG(O, A, critics=init["q"], policy=init["pi"], hidden=(32, 32))
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(3)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    assert all(len(n) == 0 for n in noises)
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "targ_pi": mg.sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
        mg.assert_params_close(final[g], p, f"ddpg {g}")
    mg.pack_case("ddpg", out, dict(obs=O, act=A, batch=B, steps=3, h0=32, h1=32), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("ddpg")

    # ---- IQL: value function + expectile regression, advantage-weighted actor, one Adam over critics + value function
    O, A, B = 6, 3, 16
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(13)
    enc = VectorEncoderFactory([32, 32])
    algo = IQL(actor_encoder_factory=enc, critic_encoder_factory=enc, value_encoder_factory=enc, batch_size=B, n_steps=2,
               weight_temp=3.0, max_weight=5.0)   # a max_weight some samples actually hit
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "v": mg.sd(impl._value_func)}
    orc = oupdate.IQL(O, A, critics=init["q"], policy=init["pi"], value=init["v"], max_weight=5.0)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(3)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    assert all(len(n) == 0 for n in noises)
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "v": mg.sd(impl._value_func),
             "targ_q": mg.sd(impl._targ_q_func), "targ_pi": mg.sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("v", orc.v), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
        mg.assert_params_close(final[g], p, f"iql {g}")
    mg.pack_case("iql", out, dict(obs=O, act=A, batch=B, steps=3, h0=32, h1=32, max_weight=5.0), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("iql")

    # ---- quantile-regression critics on the TD3 family (ContinuousQRQFunction, qr_q_function.py:91-165)
    from d3rlpy.models.q_functions import QRQFunctionFactory

    for name, cls, kw, okw in (
            ("td3bc_qr", TD3PlusBC, dict(q_func_factory=QRQFunctionFactory(n_quantiles=8), scaler=None, n_steps=2),
             dict(cls=oupdate.TD3PlusBC, nq=8)),
            ("ddpg_qr", DDPG, dict(q_func_factory="qr", n_critics=2), dict(cls=oupdate.DDPG, nq=32))):
        O, A, B = 6, 3, 16
        o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
        trs = mg.ref_transitions(o, a, r, t)
        torch.manual_seed(14 if name == "td3bc_qr" else 15)
        enc = VectorEncoderFactory([32, 32])
        algo = cls(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, **kw)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
        assert init["q"]["_q_funcs.0._fc.weight"].shape[0] == okw["nq"]
        orc = okw["cls"](O, A, critics=init["q"], policy=init["pi"], hidden=(32, 32), n_critics=2)
        n_steps = kw.get("n_steps", 1)
        batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=n_steps) for _ in range(4)]
        metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
        final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
                 "targ_pi": mg.sd(impl._targ_policy)}
        for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
            mg.assert_params_close(final[g], p, f"{name} {g}")
        mg.pack_case(name, out, dict(obs=O, act=A, batch=B, steps=4, h0=32, h1=32, n_quantiles=okw["nq"],
                                     n_steps=n_steps), init, [mg.batch_arrays(b) for b in batches], noises, metrics,
                     final)
        cases.append(name)

    out["cases"] = np.array(cases)
    path = os.path.join(HERE, "update_siblings.npz")
    np.savez_compressed(path, **out)
    print("update_siblings.npz:", cases, "%.1f KB" % (os.path.getsize(path) / 1024), [m for m in metrics])


if __name__ == "__main__":
    torch.set_num_threads(1)
    main()
