"""Golden fixtures for AWAC, CRR, PLAS, BEAR, DiscreteBCQ, DiscreteSAC, TD3PlusRelation and BC / DiscreteBC, recorded from the LIVE unmodified reference like
tests/golden/make_golden_siblings.py.  The CUDA path for AWAC is not built yet; this pins the oracle class
(oracle/update.py:AWAC) that path will be held to: non-squashed Gaussian policy with a logstd parameter in [-6, 0],
batch-softmax advantage weights with sampled state values, actor Adam with weight decay.

    python tests/golden/make_golden_awac.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

oupdate = mg.oupdate


def main():
    from d3rlpy.algos import AWAC
    from d3rlpy.models.encoders import VectorEncoderFactory

    out, cases = {}, []
    rs = np.random.RandomState(51)
    for name, n_samples, interval, lam, seed in (("awac", 1, 1, 1.0, 61), ("awac_n4", 4, 2, 0.5, 62)):
        O, A, B, steps = 6, 3, 16, 4
        o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
        trs = mg.ref_transitions(o, a, r, t)
        torch.manual_seed(seed)
        enc = VectorEncoderFactory([32, 32])
        algo = AWAC(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_steps=2,
                    n_action_samples=n_samples, update_actor_interval=interval, lam=lam)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
        orc = oupdate.AWAC(O, A, critics=init["q"], policy=init["pi"], n_action_samples=n_samples,
                           update_actor_interval=interval, lam=lam)
        batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
        metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
        final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
                 "targ_pi": mg.sd(impl._targ_policy)}
        for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
            mg.assert_params_close(final[g], p, f"{name} {g}")
        mg.pack_case(name, out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_action_samples=n_samples,
                                     update_actor_interval=interval, lam=lam), init,
                     [mg.batch_arrays(b) for b in batches], noises, metrics, final)
        cases.append(name)
    # ---- CRR: hard target copies every 2 steps with exp weights / mean advantage; soft targets with binary / max
    from d3rlpy.algos import CRR

    for name, kw, seed in (("crr", dict(target_update_type="hard", target_update_interval=2, beta=0.5, max_weight=4.0,
                                        n_action_samples=3), 63),
                           ("crr_binary_max_soft", dict(target_update_type="soft", weight_type="binary",
                                                        advantage_type="max", n_critics=2), 64)):
        O, A, B, steps = 6, 3, 16, 4
        o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
        trs = mg.ref_transitions(o, a, r, t)
        torch.manual_seed(seed)
        enc = VectorEncoderFactory([32, 32])
        algo = CRR(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_steps=2, **kw)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
        okw = {k: v for k, v in kw.items() if k != "n_critics"}
        orc = oupdate.CRR(O, A, critics=init["q"], policy=init["pi"], **okw)
        batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
        metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
        final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
                 "targ_pi": mg.sd(impl._targ_policy)}
        for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
            mg.assert_params_close(final[g], p, f"{name} {g}")
        cfg = dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, hard=float(kw["target_update_type"] == "hard"),
                   target_update_interval=kw.get("target_update_interval", 100), beta=kw.get("beta", 1.0),
                   max_weight=kw.get("max_weight", 20.0), n_action_samples=kw.get("n_action_samples", 4),
                   binary=float(kw.get("weight_type", "exp") == "binary"),
                   adv_max=float(kw.get("advantage_type", "mean") == "max"))
        mg.pack_case(name, out, cfg, init, [mg.batch_arrays(b) for b in batches], noises, metrics, final)
        cases.append(name)

    # ---- PLAS: two VAE warm-up steps, then critic every step and actor every other step
    from d3rlpy.algos import PLAS

    O, A, B, steps = 6, 3, 16, 6
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(65)
    enc, venc = VectorEncoderFactory([32, 32]), VectorEncoderFactory([48, 48])
    algo = PLAS(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=venc, batch_size=B,
                n_steps=2, warmup_steps=2, update_actor_interval=2, lam=0.6)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "imitator": mg.sd(impl._imitator)}
    orc = oupdate.PLAS(O, A, critics=init["q"], policy=init["pi"], imitator=init["imitator"], warmup_steps=2,
                       update_actor_interval=2, lam=0.6)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "imitator": mg.sd(impl._imitator),
             "targ_q": mg.sd(impl._targ_q_func), "targ_pi": mg.sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("imitator", orc.imitator), ("targ_q", orc.targ_q),
                 ("targ_pi", orc.targ_pi)):
        mg.assert_params_close(final[g], p, f"plas {g}")
    mg.pack_case("plas", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, v0=48, v1=48, warmup_steps=2,
                                   update_actor_interval=2, lam=0.6), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("plas")

    # ---- BEAR: two warm-up actor steps (MMD only), then SAC + MMD; Laplacian and Gaussian kernels
    from d3rlpy.algos import BEAR

    for name, kernel, seed in (("bear", "laplacian", 66), ("bear_gaussian", "gaussian", 67)):
        O, A, B, steps = 6, 3, 16, 4
        o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
        trs = mg.ref_transitions(o, a, r, t)
        torch.manual_seed(seed)
        enc, venc = VectorEncoderFactory([32, 32]), VectorEncoderFactory([48, 48])
        algo = BEAR(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=venc, batch_size=B,
                    n_steps=2, warmup_steps=2, n_target_samples=3, n_mmd_action_samples=4, mmd_kernel=kernel,
                    mmd_sigma=5.0, lam=0.6)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "imitator": mg.sd(impl._imitator)}
        orc = oupdate.BEAR(O, A, critics=init["q"], policy=init["pi"], imitator=init["imitator"], warmup_steps=2,
                           n_target_samples=3, n_mmd_action_samples=4, mmd_kernel=kernel, mmd_sigma=5.0, lam=0.6)
        batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
        metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
        final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "imitator": mg.sd(impl._imitator),
                 "targ_q": mg.sd(impl._targ_q_func), "targ_pi": mg.sd(impl._targ_policy),
                 "log_temp": mg.sd(impl._log_temp), "log_alpha": mg.sd(impl._log_alpha)}
        for g, p in (("q", orc.q), ("pi", orc.pi), ("imitator", orc.imitator), ("targ_q", orc.targ_q),
                     ("targ_pi", orc.targ_pi), ("log_temp", orc.log_temp), ("log_alpha", orc.log_alpha)):
            mg.assert_params_close(final[g], p, f"{name} {g}")
        mg.pack_case(name, out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, v0=48, v1=48, warmup_steps=2,
                                     n_target_samples=3, n_mmd_action_samples=4, gaussian=float(kernel == "gaussian"),
                                     mmd_sigma=5.0, lam=0.6), init,
                     [mg.batch_arrays(b) for b in batches], noises, metrics, final)
        cases.append(name)

    # ---- DiscreteBCQ (vector observations: the imitator owns its encoder), target copied every 2 steps
    from d3rlpy.algos import DiscreteBCQ

    O, A, B, steps = 6, 4, 16, 4
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A, discrete=True)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(68)
    algo = DiscreteBCQ(encoder_factory=VectorEncoderFactory([32, 32]), batch_size=B, n_steps=2, n_critics=2,
                       target_update_interval=2, action_flexibility=0.6, beta=0.3)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "imitator": mg.sd(impl._imitator)}
    orc = oupdate.DiscreteBCQ((O,), A, critics=init["q"], imitator=init["imitator"], target_update_interval=2,
                              action_flexibility=0.6, beta=0.3)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    assert all(len(n) == 0 for n in noises)
    final = {"q": mg.sd(impl._q_func), "imitator": mg.sd(impl._imitator), "targ_q": mg.sd(impl._targ_q_func)}
    for g, p in (("q", orc.q), ("imitator", orc.imitator), ("targ_q", orc.targ_q)):
        mg.assert_params_close(final[g], p, f"discrete_bcq {g}")
    xe = o[:20]
    greedy = algo.predict(xe)
    assert np.array_equal(greedy, orc.best_action(torch.tensor(xe)).numpy())
    mg.pack_case("discrete_bcq", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_critics=2,
                                           target_update_interval=2, action_flexibility=0.6, beta=0.3), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    out["discrete_bcq/eval_x"], out["discrete_bcq/predict"] = xe, greedy
    cases.append("discrete_bcq")

    # ---- DiscreteSAC: categorical policy, expectation-form soft target, Adam eps 1e-4, target copied every 2 steps
    from d3rlpy.algos import DiscreteSAC

    O, A, B, steps = 6, 4, 16, 4
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A, discrete=True)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(69)
    enc = VectorEncoderFactory([32, 32])
    algo = DiscreteSAC(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_steps=2,
                       target_update_interval=2)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.DiscreteSAC(O, A, critics=init["q"], policy=init["pi"], target_update_interval=2)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    assert all(len(n) == 0 for n in noises)
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "log_temp": mg.sd(impl._log_temp)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("log_temp", orc.log_temp)):
        mg.assert_params_close(final[g], p, f"discrete_sac {g}")
    mg.pack_case("discrete_sac", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, target_update_interval=2),
                 init, [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("discrete_sac")

    # ---- TD3PlusRelation (the fork's own algorithm): TD3+BC schedule with the batch-relational actor term
    from d3rlpy.algos import TD3PlusRelation

    O, A, B, steps = 6, 3, 16, 4
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(70)
    enc = VectorEncoderFactory([32, 32])
    algo = TD3PlusRelation(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_steps=2, scaler=None,
                           use_gpu=False)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.TD3PlusRelation(O, A, critics=init["q"], policy=init["pi"])
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "targ_pi": mg.sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
        mg.assert_params_close(final[g], p, f"td3_relation {g}")
    mg.pack_case("td3_relation", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("td3_relation")

    # ---- BC (deterministic regressor) and DiscreteBC
    from d3rlpy.algos import BC, DiscreteBC

    for name, cls, discrete, seed in (("bc", BC, False, 71), ("discrete_bc", DiscreteBC, True, 72)):
        O, A, B, steps = 6, 3, 16, 3
        o, a, r, t = mg.vector_dataset(rs, obs=O, act=A, discrete=discrete)
        trs = mg.ref_transitions(o, a, r, t)
        torch.manual_seed(seed)
        kw = dict(beta=0.3) if discrete else {}
        algo = cls(encoder_factory=VectorEncoderFactory([32, 32]), batch_size=B, **kw)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"imitator": mg.sd(impl._imitator)}
        orc = oupdate.BC(O, A, imitator=init["imitator"], discrete=discrete, beta=0.3)
        batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B)) for _ in range(steps)]
        metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
        final = {"imitator": mg.sd(impl._imitator)}
        mg.assert_params_close(final["imitator"], orc.imitator, name)
        xe = o[:20]
        pred = algo.predict(xe)
        got = orc.predict(torch.tensor(xe)).numpy()
        assert np.array_equal(pred, got) if discrete else np.allclose(pred, got, atol=1e-6)
        mg.pack_case(name, out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, discrete=float(discrete),
                                     beta=0.3), init, [mg.batch_arrays(b) for b in batches], noises, metrics, final)
        out[f"{name}/eval_x"], out[f"{name}/predict"] = xe, pred
        cases.append(name)

    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "update_awac.npz"), **out)
    print("update_awac.npz:", cases)


if __name__ == "__main__":
    main()
