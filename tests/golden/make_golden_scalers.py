"""Golden fixtures for the observation / action / reward scalers on the update path,
recorded from the LIVE unmodified reference (helpers of tests/golden/make_golden.py):

* `fit/*`        parameters each reference scaler fits on a small dataset (terminal and timed-out episodes mixed),
* `tr/*`         `transform` / `reverse_transform` of every scaler on a sampled minibatch (float32 tensors),
* case `td3bc_scaled`  three `TD3PlusBC.update` calls with MinMaxScaler + MinMaxActionScaler + StandardRewardScaler,
                       then `predict` / `predict_value` / `sample_action` on raw inputs,
* case `cql_scaled`    three `CQL.update` calls with StandardScaler + MinMaxActionScaler + ClipRewardScaler,
* case `dqn_scaled`    three `DoubleDQN.update` calls with MinMaxScaler + ReturnBasedRewardScaler,
each checked against the oracle restatement (oracle/scalers.py, oracle/update.py) before it is written.

    python tests/golden/make_golden_scalers.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through oracle/ref_import)

from oracle import scalers as osc  # noqa: E402

oupdate = mg.oupdate
d3 = mg.d3


def _close(a, b, what, tol=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))), (what, np.abs(a - b).max())


def main():
    from d3rlpy.algos import CQL, DoubleDQN, TD3PlusBC
    from d3rlpy.models.encoders import VectorEncoderFactory
    from d3rlpy.preprocessing import (ClipRewardScaler, MinMaxActionScaler, MinMaxRewardScaler, MinMaxScaler,
                                      MultiplyRewardScaler, ReturnBasedRewardScaler, StandardRewardScaler,
                                      StandardScaler)

    out, cases = {}, []
    rs = np.random.RandomState(21)

    # ------------------------------------------------------------------ fit + transform of every scaler
    O, A, n, ep = 5, 3, 200, 25
    o = (rs.randn(n, O) * np.array([1, 5, 0.1, 2, 10]) + np.array([0, 3, -1, 0.5, -20])).astype(np.float32)
    a = (rs.uniform(-1, 1, (n, A)) * np.array([2.0, 0.5, 10.0]) + np.array([0.5, 0.0, -3.0])).astype(np.float32)
    r = (rs.randn(n) * 3 + 1).astype(np.float32)
    ept = np.zeros(n, np.float32)
    ept[ep - 1::ep] = 1.0
    t = ept.copy()
    t[ep - 1::2 * ep] = 0.0   # every other episode ends by time-out: its last step is not a transition
    ds = d3.dataset.MDPDataset(o, a, r, t, episode_terminals=ept)
    trs = [tr for e in ds.episodes for tr in e.transitions]
    tr_obs = np.stack([np.asarray(tr.observation) for tr in trs])
    tr_act = np.stack([np.asarray(tr.action) for tr in trs])
    tr_rew = np.array([tr.reward for tr in trs], np.float32)
    tr_ep = np.concatenate([[i] * len(e.transitions) for i, e in enumerate(ds.episodes)])
    out.update({"data/observations": o, "data/actions": a, "data/rewards": r, "data/terminals": t,
                "data/episode_terminals": ept})
    idx = rs.randint(len(trs), size=32)
    batch = mg.batch_arrays(mg.ref_batch(trs, idx, n_steps=2))
    out["tr/indices"] = idx.astype(np.int64)
    for k, v in batch.items():
        out[f"tr/batch/{k}"] = v
    x = torch.tensor(batch["observations"])
    act = torch.tensor(batch["actions"])
    rew = torch.tensor(batch["rewards"])
    unit = torch.tensor(rs.uniform(-1, 1, (32, A)).astype(np.float32))   # policy outputs to map back
    out["tr/unit_actions"] = unit.numpy()
    rets = osc.episode_returns(tr_rew, tr_ep)

    sc = MinMaxScaler()
    sc.fit(trs)
    osc_mm = osc.MinMaxScaler().fit(tr_obs)
    assert np.array_equal(sc._minimum, osc_mm.minimum) and np.array_equal(sc._maximum, osc_mm.maximum)
    out.update({"fit/min_max/minimum": sc._minimum, "fit/min_max/maximum": sc._maximum,
                "tr/min_max": sc.transform(x).numpy()})
    assert np.array_equal(osc_mm(x).numpy(), out["tr/min_max"])

    sc = StandardScaler()
    sc.fit(trs)
    osc_st = osc.StandardScaler().fit(tr_obs)
    _close(osc_st.mean, sc._mean, "standard mean", 1e-12)
    _close(osc_st.std, sc._std, "standard std", 1e-12)
    out.update({"fit/standard/mean": sc._mean, "fit/standard/std": sc._std, "tr/standard": sc.transform(x).numpy()})
    _close(osc_st(x).numpy(), out["tr/standard"], "standard transform")

    asc = MinMaxActionScaler()
    asc.fit(trs)
    osc_a = osc.MinMaxActionScaler().fit(tr_act)
    assert np.array_equal(asc._minimum, osc_a.minimum) and np.array_equal(asc._maximum, osc_a.maximum)
    out.update({"fit/action_min_max/minimum": asc._minimum, "fit/action_min_max/maximum": asc._maximum,
                "tr/action_min_max": asc.transform(act).numpy(),
                "tr/action_min_max_reverse": asc.reverse_transform(unit).numpy()})
    assert np.array_equal(osc_a(act).numpy(), out["tr/action_min_max"])
    assert np.array_equal(osc_a.reverse(unit).numpy(), out["tr/action_min_max_reverse"])

    reward_cases = {
        "multiply": (MultiplyRewardScaler(multiplier=0.25), osc.MultiplyRewardScaler(0.25)),
        "clip": (ClipRewardScaler(-1.0, 1.5, multiplier=2.0), osc.ClipRewardScaler(-1.0, 1.5, 2.0)),
        "min_max": (MinMaxRewardScaler(multiplier=3.0), osc.MinMaxRewardScaler(multiplier=3.0)),
        "standard": (StandardRewardScaler(multiplier=0.5), osc.StandardRewardScaler(multiplier=0.5)),
        "return": (ReturnBasedRewardScaler(multiplier=1000.0), osc.ReturnBasedRewardScaler(multiplier=1000.0)),
    }
    for name, (ref, orc) in reward_cases.items():
        ref.fit(trs)
        orc.fit(tr_rew, rets)
        params = {k: v for k, v in ref.get_params().items() if v is not None}
        for k, v in params.items():
            out[f"fit/reward_{name}/{k}"] = np.float64(v)
            ov = {"minimum": "minimum", "maximum": "maximum", "mean": "mean", "std": "std", "eps": "eps",
                  "return_max": "return_max", "return_min": "return_min", "multiplier": "multiplier", "low": "low",
                  "high": "high"}[k]
            _close(getattr(orc, ov), v, f"{name} {k}", 1e-12)
        out[f"tr/reward_{name}"] = ref.transform(rew).numpy()
        _close(orc(rew).numpy(), out[f"tr/reward_{name}"], f"reward {name} transform", 1e-6)

    # ------------------------------------------------------------------ TD3+BC: min_max obs / min_max action / standard reward
    B, steps = 16, 3
    torch.manual_seed(31)
    enc = VectorEncoderFactory([32, 32])
    sc, asc, rsc = MinMaxScaler(), MinMaxActionScaler(), StandardRewardScaler(multiplier=2.0)
    for s_ in (sc, asc, rsc):
        s_.fit(trs)
    algo = TD3PlusBC(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, scaler=sc, action_scaler=asc,
                     reward_scaler=rsc, n_steps=2)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.TD3PlusBC(O, A, critics=init["q"], policy=init["pi"])
    o_sc = osc.MinMaxScaler(sc._minimum, sc._maximum)
    o_asc = osc.MinMaxActionScaler(asc._minimum, asc._maximum)
    o_rsc = osc.StandardRewardScaler(rsc._mean, rsc._std, rsc._eps, 2.0)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
    obatches = [oupdate.Batch(mg.batch_arrays(b), o_sc, o_rsc, o_asc) for b in batches]
    metrics, noises = mg.run_steps(algo, orc, batches, obatches)
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "targ_pi": mg.sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
        mg.assert_params_close(final[g], p, f"td3bc_scaled {g}")
    mg.pack_case("td3bc_scaled", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_steps=2,
                                           reward_mean=rsc._mean, reward_std=rsc._std, reward_eps=rsc._eps,
                                           reward_multiplier=2.0), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    out["td3bc_scaled/obs_minimum"], out["td3bc_scaled/obs_maximum"] = sc._minimum, sc._maximum
    out["td3bc_scaled/act_minimum"], out["td3bc_scaled/act_maximum"] = asc._minimum, asc._maximum
    xe, ae = o[:24], a[:24]
    out["td3bc_scaled/eval_x"], out["td3bc_scaled/eval_action"] = xe, ae
    out["td3bc_scaled/predict"] = algo.predict(xe)
    out["td3bc_scaled/predict_value"] = algo.predict_value(xe, ae)
    cases.append("td3bc_scaled")

    # ------------------------------------------------------------------ CQL: standard obs / min_max action / clip reward
    torch.manual_seed(32)
    enc = VectorEncoderFactory([32, 32])
    sc, asc, rsc = StandardScaler(), MinMaxActionScaler(), ClipRewardScaler(-1.0, 1.0, multiplier=0.5)
    for s_ in (sc, asc, rsc):
        s_.fit(trs)
    N = 4
    algo = CQL(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_action_samples=N, scaler=sc,
               action_scaler=asc, reward_scaler=rsc)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy)}
    orc = oupdate.CQL(O, A, critics=init["q"], policy=init["pi"], n_action_samples=N)
    o_sc = osc.StandardScaler(sc._mean, sc._std, sc._eps)
    o_asc = osc.MinMaxActionScaler(asc._minimum, asc._maximum)
    o_rsc = osc.ClipRewardScaler(-1.0, 1.0, 0.5)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B)) for _ in range(steps)]
    obatches = [oupdate.Batch(mg.batch_arrays(b), o_sc, o_rsc, o_asc) for b in batches]
    metrics, noises = mg.run_steps(algo, orc, batches, obatches)
    final = {"q": mg.sd(impl._q_func), "pi": mg.sd(impl._policy), "targ_q": mg.sd(impl._targ_q_func),
             "log_temp": mg.sd(impl._log_temp), "log_alpha": mg.sd(impl._log_alpha)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("log_temp", orc.log_temp),
                 ("log_alpha", orc.log_alpha)):
        mg.assert_params_close(final[g], p, f"cql_scaled {g}")
    mg.pack_case("cql_scaled", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_action_samples=N,
                                         reward_low=-1.0, reward_high=1.0, reward_multiplier=0.5), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    out["cql_scaled/obs_mean"], out["cql_scaled/obs_std"] = sc._mean, sc._std
    out["cql_scaled/act_minimum"], out["cql_scaled/act_maximum"] = asc._minimum, asc._maximum
    out["cql_scaled/eval_x"] = xe
    out["cql_scaled/predict"] = algo.predict(xe)
    cases.append("cql_scaled")

    # ------------------------------------------------------------------ DoubleDQN: min_max obs / return-based reward
    ad = rs.randint(0, 4, size=n).astype(np.int32)
    dsd = d3.dataset.MDPDataset(o, ad, r, t, episode_terminals=ept, discrete_action=True)
    trd = [tr for e in dsd.episodes for tr in e.transitions]
    torch.manual_seed(33)
    sc, rsc = MinMaxScaler(), ReturnBasedRewardScaler(multiplier=100.0)
    sc.fit(trd)
    rsc.fit(trd)
    algo = DoubleDQN(encoder_factory=VectorEncoderFactory([32, 32]), batch_size=B, scaler=sc, reward_scaler=rsc,
                     target_update_interval=2)
    algo.create_impl((O,), 4)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func)}
    orc = oupdate.DiscreteCQL((O,), 4, critics=init["q"], double=True, conservative=False, target_update_interval=2)
    o_sc = osc.MinMaxScaler(sc._minimum, sc._maximum)
    o_rsc = osc.ReturnBasedRewardScaler(rsc._return_max, rsc._return_min, 100.0)
    batches = [mg.ref_batch(trd, rs.randint(len(trd), size=B)) for _ in range(steps)]
    obatches = [oupdate.Batch(mg.batch_arrays(b), o_sc, o_rsc) for b in batches]
    metrics, noises = mg.run_steps(algo, orc, batches, obatches)
    final = {"q": mg.sd(impl._q_func), "targ_q": mg.sd(impl._targ_q_func)}
    for g, p in (("q", orc.q), ("targ_q", orc.targ_q)):
        mg.assert_params_close(final[g], p, f"dqn_scaled {g}")
    mg.pack_case("dqn_scaled", out, dict(obs=O, act=4, batch=B, steps=steps, h0=32, h1=32, return_max=rsc._return_max,
                                         return_min=rsc._return_min, reward_multiplier=100.0), init,
                 [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    out["dqn_scaled/obs_minimum"], out["dqn_scaled/obs_maximum"] = sc._minimum, sc._maximum
    cases.append("dqn_scaled")

    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "scalers.npz"), **out)
    print("scalers.npz:", cases, f"{os.path.getsize(os.path.join(HERE, 'scalers.npz')) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
