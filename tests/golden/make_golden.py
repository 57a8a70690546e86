"""Generate the golden fixtures in this directory from the LIVE reference.

Runs only in the build container (needs /root/reference + oracle/_ref).  For each
case it (1) drives the unmodified reference (`d3rlpy` 1.1.0, use_gpu=False) for a
few `algo.update(batch)` calls while recording every random draw, (2) replays the
same weights / minibatches / noise through the oracle restatement
(`oracle/update.py`, `oracle/sampler.py`) and asserts agreement, and (3) writes the
inputs and the REFERENCE's outputs to `tests/golden/*.npz`.

    python tests/golden/make_golden.py

The fixtures are what pins the oracle on the GPU box, where the reference is absent.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_import  # noqa: E402

d3 = ref_import.load()
from oracle import sampler as osampler  # noqa: E402
from oracle import update as oupdate  # noqa: E402


# ----------------------------------------------------------------------------- noise capture
class NoiseTap:
    """Records torch.randn / Tensor.uniform_ / Normal.rsample draws made by the reference
    (patch the name inside torch.distributions.normal)."""

    def __enter__(self):
        import torch.distributions.normal as tdn

        self.log = []
        self._randn, self._uniform, self._std = torch.randn, torch.Tensor.uniform_, tdn._standard_normal
        tap = self

        def randn(*a, **k):
            t = tap._randn(*a, **k)
            tap.log.append(t.detach().clone())
            return t

        def uniform_(self_t, *a, **k):
            t = tap._uniform(self_t, *a, **k)
            tap.log.append(t.detach().clone())
            return t

        def std_normal(shape, dtype, device):
            t = tap._std(shape, dtype, device)
            tap.log.append(t.detach().clone())
            return t

        torch.randn, torch.Tensor.uniform_, tdn._standard_normal = randn, uniform_, std_normal
        self._tdn = tdn
        return self

    def __exit__(self, *exc):
        torch.randn, torch.Tensor.uniform_ = self._randn, self._uniform
        self._tdn._standard_normal = self._std


def sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def flat(prefix, d):
    return {f"{prefix}/{k}": v.detach().cpu().numpy() for k, v in d.items()}


# ----------------------------------------------------------------------------- datasets
def vector_dataset(rs, n=400, obs=5, act=3, ep=50, discrete=False):
    o = rs.randn(n, obs).astype(np.float32)
    a = rs.randint(0, act, size=n).astype(np.int32) if discrete else rs.uniform(-1, 1, (n, act)).astype(np.float32)
    r = rs.randn(n).astype(np.float32)
    t = np.zeros(n, np.float32)
    t[ep - 1::ep] = 1.0
    return o, a, r, t


def image_dataset(rs, n=120, hw=42, act=4, ep=40):
    o = rs.randint(0, 256, size=(n, 1, hw, hw)).astype(np.uint8)
    a = rs.randint(0, act, size=n).astype(np.int32)
    r = (rs.rand(n) < 0.1).astype(np.float32)
    t = np.zeros(n, np.float32)
    t[ep - 1::ep] = 1.0
    return o, a, r, t


def ref_transitions(o, a, r, t, episode_terminals=None):
    ds = d3.dataset.MDPDataset(o, a, r, t, episode_terminals=episode_terminals)
    return [tr for e in ds.episodes for tr in e.transitions]


def ref_batch(trs, idx, n_frames=1, n_steps=1, gamma=0.99):
    return d3.dataset.TransitionMiniBatch([trs[i] for i in idx], n_frames=n_frames, n_steps=n_steps, gamma=gamma)


def batch_arrays(b):
    return dict(observations=np.array(b.observations), actions=np.array(b.actions), rewards=np.array(b.rewards),
                next_observations=np.array(b.next_observations), terminals=np.array(b.terminals),
                n_steps=np.array(b.n_steps))


# ----------------------------------------------------------------------------- sampler goldens
def make_sampler():
    rs = np.random.RandomState(7)
    out = {}
    cases = []
    for kind in ("vector", "image"):
        for discrete in (False, True):
            for term_eps in (True, False):
                if kind == "vector":
                    o, a, r, t = vector_dataset(rs, n=90, obs=6, act=3, ep=30, discrete=discrete)
                else:
                    o, a, r, t = image_dataset(rs, n=60, hw=12, act=4, ep=20)
                    if not discrete:
                        a = rs.uniform(-1, 1, (60, 2)).astype(np.float32)
                ept = t.copy()
                if not term_eps:  # episodes that end by timeout: last step is dropped (dataset.pyx:91-92)
                    t = np.zeros_like(t)
                    t[ept.nonzero()[0][0]] = 1.0  # keep one true terminal for mixing
                trs = ref_transitions(o, a, r, t, ept)
                replay = osampler.FlatReplay(o, a, r, t, ept)
                assert len(replay) == len(trs), (len(replay), len(trs))
                for n_frames in (1, 4):
                    for n_steps in (1, 3):
                        idx = rs.randint(len(trs), size=24)
                        # force the edges in: episode starts and ends
                        idx[:4] = [0, len(trs) - 1, int(replay.ep_last[0]), int(replay.ep_last[0]) + 1]
                        rb = batch_arrays(ref_batch(trs, idx, n_frames, n_steps, 0.99))
                        ob = osampler.gather(replay, idx, n_frames, n_steps, 0.99)
                        for k in rb:
                            if k == "rewards":
                                assert np.allclose(rb[k], ob[k], rtol=1e-6, atol=1e-7), k
                            else:
                                assert rb[k].dtype == ob[k].dtype and np.array_equal(rb[k], ob[k]), (k, kind, n_frames, n_steps)
                        name = f"{kind}_{'disc' if discrete else 'cont'}_{'term' if term_eps else 'trunc'}_f{n_frames}_s{n_steps}"
                        cases.append(name)
                        out.update({f"{name}/data/observations": o, f"{name}/data/actions": a, f"{name}/data/rewards": r,
                                    f"{name}/data/terminals": t, f"{name}/data/episode_terminals": ept,
                                    f"{name}/indices": idx.astype(np.int64),
                                    f"{name}/cfg": np.array([n_frames, n_steps], np.int64)})
                        out.update({f"{name}/ref/{k}": v for k, v in rb.items()})
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler.npz:", len(cases), "cases")


# ----------------------------------------------------------------------------- update goldens
def run_steps(algo, oracle, batches, obatches, n_noise_expected=None):
    """Runs reference + oracle; returns (ref metrics list, noise per step)."""
    ref_metrics, noises = [], []
    for b, ob in zip(batches, obatches):
        with NoiseTap() as tap:
            m = algo.update(b)
        ref_metrics.append({k: float(v) for k, v in m.items()})
        noises.append(tap.log)
        om = oracle.update(ob, oupdate.Noise(injected=tap.log))
        assert set(om) == set(m), (set(om), set(m))
        for k in m:
            assert abs(om[k] - float(m[k])) <= 1e-5 * max(1.0, abs(float(m[k]))), (k, om[k], float(m[k]))
    return ref_metrics, noises


def assert_params_close(ref_sd, oracle_params, what, tol=2e-6):
    for k, v in ref_sd.items():
        d = (oracle_params[k].detach() - v).abs().max().item()
        s = v.abs().max().item()
        assert d <= tol * max(1.0, s), (what, k, d, s)


def pack_case(name, out, cfg, init, batches_np, noises, metrics, final):
    out[f"{name}/cfg_keys"] = np.array(list(cfg.keys()))
    out[f"{name}/cfg_vals"] = np.array([float(v) for v in cfg.values()], np.float64)
    for grp, d in init.items():
        out.update(flat(f"{name}/init/{grp}", d))
    for grp, d in final.items():
        out.update(flat(f"{name}/final/{grp}", d))
    for s, b in enumerate(batches_np):
        for k, v in b.items():
            out[f"{name}/batch{s}/{k}"] = v
    for s, ns in enumerate(noises):
        for j, t in enumerate(ns):
            out[f"{name}/noise{s}/{j}"] = t.numpy()
    keys = sorted({k for m in metrics for k in m})
    out[f"{name}/metric_keys"] = np.array(keys)
    out[f"{name}/metrics"] = np.array([[m.get(k, np.nan) for k in keys] for m in metrics], np.float64)


def make_updates():
    from d3rlpy.algos import BCQ, CQL, DiscreteCQL, TD3PlusBC
    from d3rlpy.models.encoders import PixelEncoderFactory, VectorEncoderFactory

    out = {}
    cases = []
    rs = np.random.RandomState(3)
    steps = 3

    # ---- TD3+BC (c1-shaped, small): standard scaler, n_steps 1
    O, A, B = 5, 3, 16
    o, a, r, t = vector_dataset(rs, obs=O, act=A)
    trs = ref_transitions(o, a, r, t)
    mean, std = o.mean(0), o.std(0)
    torch.manual_seed(0)
    enc = VectorEncoderFactory([32, 32])
    algo = TD3PlusBC(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B,
                     scaler=d3.preprocessing.StandardScaler(mean=mean, std=std))
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": sd(impl._q_func), "pi": sd(impl._policy)}
    orc = oupdate.TD3PlusBC(O, A, critics=init["q"], policy=init["pi"])
    idxs = [rs.randint(len(trs), size=B) for _ in range(steps)]
    batches = [ref_batch(trs, i) for i in idxs]
    scaler = oupdate.standard_scaler(mean, std)
    metrics, noises = run_steps(algo, orc, batches, [oupdate.Batch(batch_arrays(b), scaler) for b in batches])
    final = {"q": sd(impl._q_func), "pi": sd(impl._policy), "targ_q": sd(impl._targ_q_func), "targ_pi": sd(impl._targ_policy)}
    assert_params_close(final["q"], orc.q, "td3bc q")
    assert_params_close(final["pi"], orc.pi, "td3bc pi")
    assert_params_close(final["targ_q"], orc.targ_q, "td3bc targ_q")
    assert_params_close(final["targ_pi"], orc.targ_pi, "td3bc targ_pi")
    pack_case("td3bc", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32), init,
              [batch_arrays(b) for b in batches], noises, metrics, final)
    out["td3bc/scaler_mean"], out["td3bc/scaler_std"] = mean, std
    cases.append("td3bc")

    # ---- CQL (c2-shaped, small) with n_steps=3 batches to exercise gamma**n
    O, A, B, N = 6, 3, 16, 4
    o, a, r, t = vector_dataset(rs, obs=O, act=A)
    trs = ref_transitions(o, a, r, t)
    for variant, kw in (("cql", {}), ("cql_softq", {"soft_q_backup": True})):
        torch.manual_seed(1)
        enc = VectorEncoderFactory([32, 32, 32])
        algo = CQL(actor_encoder_factory=enc, critic_encoder_factory=enc, batch_size=B, n_action_samples=N,
                   n_steps=3, **kw)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"q": sd(impl._q_func), "pi": sd(impl._policy)}
        orc = oupdate.CQL(O, A, critics=init["q"], policy=init["pi"], n_action_samples=N, **kw)
        idxs = [rs.randint(len(trs), size=B) for _ in range(steps)]
        batches = [ref_batch(trs, i, n_steps=3) for i in idxs]
        metrics, noises = run_steps(algo, orc, batches, [oupdate.Batch(batch_arrays(b)) for b in batches])
        final = {"q": sd(impl._q_func), "pi": sd(impl._policy), "targ_q": sd(impl._targ_q_func),
                 "targ_pi": sd(impl._targ_policy), "log_temp": sd(impl._log_temp), "log_alpha": sd(impl._log_alpha)}
        for g, p in (("q", orc.q), ("pi", orc.pi), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi),
                     ("log_temp", orc.log_temp), ("log_alpha", orc.log_alpha)):
            assert_params_close(final[g], p, f"{variant} {g}")
        pack_case(variant, out, dict(obs=O, act=A, batch=B, steps=steps, n=N, h0=32, h1=32, h2=32,
                                     soft_q_backup=int(bool(kw))), init,
                  [batch_arrays(b) for b in batches], noises, metrics, final)
        cases.append(variant)

    # ---- BCQ (c3-shaped, small)
    O, A, B, N = 6, 3, 8, 5
    torch.manual_seed(2)
    enc, venc = VectorEncoderFactory([40, 24]), VectorEncoderFactory([48, 48])
    algo = BCQ(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=venc,
               batch_size=B, n_action_samples=N)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": sd(impl._q_func), "pi": sd(impl._policy), "imitator": sd(impl._imitator)}
    orc = oupdate.BCQ(O, A, critics=init["q"], policy=init["pi"], imitator=init["imitator"], n_action_samples=N)
    idxs = [rs.randint(len(trs), size=B) for _ in range(steps)]
    batches = [ref_batch(trs, i) for i in idxs]
    metrics, noises = run_steps(algo, orc, batches, [oupdate.Batch(batch_arrays(b)) for b in batches])
    final = {"q": sd(impl._q_func), "pi": sd(impl._policy), "imitator": sd(impl._imitator),
             "targ_q": sd(impl._targ_q_func), "targ_pi": sd(impl._targ_policy)}
    for g, p in (("q", orc.q), ("pi", orc.pi), ("imitator", orc.imitator), ("targ_q", orc.targ_q), ("targ_pi", orc.targ_pi)):
        assert_params_close(final[g], p, f"bcq {g}")
    pack_case("bcq", out, dict(obs=O, act=A, batch=B, steps=steps, n=N, h0=40, h1=24, v0=48, v1=48), init,
              [batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("bcq")

    # ---- DiscreteCQL, vector observations, 2 critics
    O, A, B = 6, 4, 16
    o, a, r, t = vector_dataset(rs, obs=O, act=A, discrete=True)
    trs = ref_transitions(o, a, r, t)
    torch.manual_seed(3)
    algo = DiscreteCQL(encoder_factory=VectorEncoderFactory([32, 32]), batch_size=B, n_critics=2,
                       target_update_interval=2)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": sd(impl._q_func)}
    orc = oupdate.DiscreteCQL((O,), A, critics=init["q"], target_update_interval=2)
    idxs = [rs.randint(len(trs), size=B) for _ in range(steps)]
    batches = [ref_batch(trs, i) for i in idxs]
    metrics, noises = run_steps(algo, orc, batches, [oupdate.Batch(batch_arrays(b)) for b in batches])
    final = {"q": sd(impl._q_func), "targ_q": sd(impl._targ_q_func)}
    assert_params_close(final["q"], orc.q, "dcql q")
    assert_params_close(final["targ_q"], orc.targ_q, "dcql targ")
    pack_case("dcql_vec", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_critics=2, interval=2), init,
              [batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("dcql_vec")

    # ---- DiscreteCQL, pixels (c4-shaped, small 42x42 frames, n_frames=4, pixel scaler)
    HW, A, B = 42, 4, 8
    o, a, r, t = image_dataset(rs, hw=HW, act=A)
    trs = ref_transitions(o, a, r, t)
    torch.manual_seed(4)
    algo = DiscreteCQL(encoder_factory=PixelEncoderFactory(feature_size=64), batch_size=B, n_frames=4,
                       scaler="pixel", target_update_interval=8000)
    algo.create_impl((4, HW, HW), A)
    impl = algo._impl
    init = {"q": sd(impl._q_func)}
    orc = oupdate.DiscreteCQL((4, HW, HW), A, critics=init["q"])
    idxs = [rs.randint(len(trs), size=B) for _ in range(steps)]
    batches = [ref_batch(trs, i, n_frames=4) for i in idxs]
    metrics, noises = run_steps(algo, orc, batches,
                                [oupdate.Batch(batch_arrays(b), oupdate.pixel_scaler()) for b in batches])
    final = {"q": sd(impl._q_func), "targ_q": sd(impl._targ_q_func)}
    assert_params_close(final["q"], orc.q, "dcql_pix q")
    assert_params_close(final["targ_q"], orc.targ_q, "dcql_pix targ")
    pack_case("dcql_pix", out, dict(hw=HW, act=A, batch=B, steps=steps, n_frames=4, feature=64), init,
              [batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("dcql_pix")

    out["cases"] = np.array(cases)
    # fixtures are fp32; keep them small
    path = os.path.join(HERE, "update.npz")
    np.savez_compressed(path, **out)
    print("update.npz:", cases, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_sampler()
    make_updates()
