"""Golden fixtures for the quantile-regression Q head: DiscreteCQL / DQN with
``QRQFunctionFactory`` recorded from the LIVE unmodified reference exactly like tests/golden/make_golden.py does for the
mean Q function (same helpers: reference-vs-oracle agreement check, case packing), plus plain DQN / DoubleDQN cases
with the mean Q function.

    python tests/golden/make_golden_qr.py        (build container only: needs /root/reference)
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
    from d3rlpy.algos import DQN, NFQ, DiscreteCQL, DoubleDQN
    from d3rlpy.models.encoders import PixelEncoderFactory, VectorEncoderFactory
    from d3rlpy.models.q_functions import QRQFunctionFactory

    out, cases = {}, []
    rs = np.random.RandomState(21)
    steps = 3

    # ---- DiscreteCQL + QR, vector observations, 2 critics (member picked by its mean), n_steps = 3, hard sync at 0, 2
    O, A, B, NQ = 6, 4, 16, 8
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A, discrete=True)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(7)
    algo = DiscreteCQL(encoder_factory=VectorEncoderFactory([32, 32]), q_func_factory=QRQFunctionFactory(n_quantiles=NQ),
                       batch_size=B, n_critics=2, n_steps=3, target_update_interval=2)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func)}
    orc = oupdate.DiscreteCQL((O,), A, critics=init["q"], target_update_interval=2, n_quantiles=NQ)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=3) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    final = {"q": mg.sd(impl._q_func), "targ_q": mg.sd(impl._targ_q_func)}
    mg.assert_params_close(final["q"], orc.q, "qr_dcql_vec q")
    mg.assert_params_close(final["targ_q"], orc.targ_q, "qr_dcql_vec targ")
    mg.pack_case("qr_dcql_vec", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_critics=2, interval=2,
                                          n_quantiles=NQ), init, [mg.batch_arrays(b) for b in batches], noises,
                 metrics, final)
    cases.append("qr_dcql_vec")

    # ---- DQN + QR (DQNImpl.compute_target: greedy action of the TARGET network; no conservative term), default
    # n_quantiles = 32, 2 critics
    O, A, B, NQ = 5, 3, 16, 32
    o, a, r, t = mg.vector_dataset(rs, obs=O, act=A, discrete=True)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(8)
    algo = DQN(encoder_factory=VectorEncoderFactory([32, 32]), q_func_factory="qr", batch_size=B, n_critics=2,
               target_update_interval=3)
    algo.create_impl((O,), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func)}
    orc = oupdate.DiscreteCQL((O,), A, critics=init["q"], target_update_interval=3, n_quantiles=NQ, double=False,
                              conservative=False)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B)) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
    final = {"q": mg.sd(impl._q_func), "targ_q": mg.sd(impl._targ_q_func)}
    mg.assert_params_close(final["q"], orc.q, "qr_dqn_vec q")
    mg.assert_params_close(final["targ_q"], orc.targ_q, "qr_dqn_vec targ")
    mg.pack_case("qr_dqn_vec", out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_critics=2, interval=3,
                                         n_quantiles=NQ), init, [mg.batch_arrays(b) for b in batches], noises,
                 metrics, final)
    cases.append("qr_dqn_vec")

    # ---- DiscreteCQL + QR, pixels (c4-shaped, 42x42 frames, n_frames = 4, pixel scaler) — the configuration of
    # reproductions/offline/discrete_cql.py:27-28 with fewer quantiles
    HW, A, B, NQ = 42, 4, 8, 24
    o, a, r, t = mg.image_dataset(rs, hw=HW, act=A)
    trs = mg.ref_transitions(o, a, r, t)
    torch.manual_seed(9)
    algo = DiscreteCQL(encoder_factory=PixelEncoderFactory(feature_size=64),
                       q_func_factory=QRQFunctionFactory(n_quantiles=NQ), batch_size=B, n_frames=4, scaler="pixel")
    algo.create_impl((4, HW, HW), A)
    impl = algo._impl
    init = {"q": mg.sd(impl._q_func)}
    orc = oupdate.DiscreteCQL((4, HW, HW), A, critics=init["q"], n_quantiles=NQ)
    batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_frames=4) for _ in range(steps)]
    metrics, noises = mg.run_steps(algo, orc, batches,
                                   [oupdate.Batch(mg.batch_arrays(b), oupdate.pixel_scaler()) for b in batches])
    final = {"q": mg.sd(impl._q_func), "targ_q": mg.sd(impl._targ_q_func)}
    mg.assert_params_close(final["q"], orc.q, "qr_dcql_pix q")
    mg.assert_params_close(final["targ_q"], orc.targ_q, "qr_dcql_pix targ")
    mg.pack_case("qr_dcql_pix", out, dict(hw=HW, act=A, batch=B, steps=steps, n_frames=4, feature=64, n_quantiles=NQ),
                 init, [mg.batch_arrays(b) for b in batches], noises, metrics, final)
    cases.append("qr_dcql_pix")

    # ---- plain DQN and DoubleDQN with the mean Q function (dqn_impl.py:97-171 without the conservative term): the
    # update.npz fixtures only hold DiscreteCQL, so the two target rules are pinned here
    for name, cls, double, seed in (("dqn_vec", DQN, False, 10), ("ddqn_vec", DoubleDQN, True, 11),
                                    ("nfq_vec", NFQ, False, 12)):
        O, A, B = 6, 5, 16
        o, a, r, t = mg.vector_dataset(rs, obs=O, act=A, discrete=True)
        trs = mg.ref_transitions(o, a, r, t)
        torch.manual_seed(seed)
        interval = 1 if name == "nfq_vec" else 2   # NFQ: target copied after every update (nfq.py:127-131)
        kw = {} if name == "nfq_vec" else {"target_update_interval": interval}
        algo = cls(encoder_factory=VectorEncoderFactory([32, 32]), batch_size=B, n_critics=2, n_steps=2, **kw)
        algo.create_impl((O,), A)
        impl = algo._impl
        init = {"q": mg.sd(impl._q_func)}
        orc = oupdate.DiscreteCQL((O,), A, critics=init["q"], target_update_interval=interval, double=double,
                                  conservative=False)
        batches = [mg.ref_batch(trs, rs.randint(len(trs), size=B), n_steps=2) for _ in range(steps)]
        metrics, noises = mg.run_steps(algo, orc, batches, [oupdate.Batch(mg.batch_arrays(b)) for b in batches])
        final = {"q": mg.sd(impl._q_func), "targ_q": mg.sd(impl._targ_q_func)}
        mg.assert_params_close(final["q"], orc.q, f"{name} q")
        mg.assert_params_close(final["targ_q"], orc.targ_q, f"{name} targ")
        mg.pack_case(name, out, dict(obs=O, act=A, batch=B, steps=steps, h0=32, h1=32, n_critics=2, interval=interval,
                                     n_quantiles=0), init, [mg.batch_arrays(b) for b in batches], noises, metrics,
                     final)
        cases.append(name)

    out["cases"] = np.array(cases)
    path = os.path.join(HERE, "update_qr.npz")
    np.savez_compressed(path, **out)
    print("update_qr.npz:", cases, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    torch.set_num_threads(1)
    main()
