"""Golden fixtures for the callable Q-function API (SURVEY.md section 8b "signatures to keep"): the LIVE unmodified
reference's EnsembleContinuousQFunction / EnsembleDiscreteQFunction (d3rlpy/models/torch/q_functions/
ensemble_q_function.py:69-184) as built by CQL / DoubleDQN `create_impl`, evaluated on recorded inputs:
`__call__` under all five reductions, `compute_target` (min / mix with lam), `compute_error` with a float and with a
per-row gamma, `q_funcs[i]` single-member calls.

    python tests/golden/make_golden_qfunc.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through oracle/ref_import)


def main():
    from d3rlpy.algos import CQL, DoubleDQN
    from d3rlpy.models.encoders import VectorEncoderFactory

    out = {}
    rs = np.random.RandomState(21)
    # ---- continuous: 3 critics, 2x32 encoders
    O, A, n, E = 7, 3, 37, 3
    torch.manual_seed(3)
    enc = VectorEncoderFactory([32, 32])
    algo = CQL(actor_encoder_factory=enc, critic_encoder_factory=enc, n_critics=E)
    algo.create_impl((O,), A)
    qf = algo._impl._q_func
    for k, v in mg.sd(qf).items():
        out[f"cont/q/{k}"] = v.numpy()
    x = torch.tensor(rs.randn(n, O).astype(np.float32))
    a = torch.tensor(rs.uniform(-1, 1, (n, A)).astype(np.float32))
    r = torch.tensor(rs.randn(n, 1).astype(np.float32))
    tgt = torch.tensor(rs.randn(n, 1).astype(np.float32))
    term = torch.tensor((rs.rand(n, 1) < 0.2).astype(np.float32))
    g_rows = torch.tensor((0.99 ** rs.randint(1, 4, (n, 1))).astype(np.float32))
    out.update({"cont/x": x.numpy(), "cont/a": a.numpy(), "cont/r": r.numpy(), "cont/target": tgt.numpy(),
                "cont/term": term.numpy(), "cont/gamma_rows": g_rows.numpy(), "cont/cfg": np.array([O, A, n, E])})
    with torch.no_grad():
        for red in ("min", "max", "mean", "none", "mix"):
            out[f"cont/call/{red}"] = qf(x, a, red).numpy()
        out["cont/target/min"] = qf.compute_target(x, a).numpy()
        out["cont/target/mix"] = qf.compute_target(x, a, "mix", 0.6).numpy()
        out["cont/error/float"] = qf.compute_error(x, a, r, tgt, term, 0.99).numpy()
        out["cont/error/rows"] = qf.compute_error(x, a, r, tgt, term, g_rows).numpy()
        for e in range(E):
            out[f"cont/member{e}"] = qf.q_funcs[e](x, a).numpy()
    # ---- discrete: 2 critics
    O, A, n, E = 6, 4, 29, 2
    torch.manual_seed(4)
    algo = DoubleDQN(encoder_factory=VectorEncoderFactory([32, 32]), n_critics=E)
    algo.create_impl((O,), A)
    qf = algo._impl._q_func
    for k, v in mg.sd(qf).items():
        out[f"disc/q/{k}"] = v.numpy()
    x = torch.tensor(rs.randn(n, O).astype(np.float32))
    a = torch.tensor(rs.randint(0, A, n).astype(np.int64))
    r = torch.tensor(rs.randn(n, 1).astype(np.float32))
    tgt = torch.tensor(rs.randn(n, 1).astype(np.float32))
    term = torch.tensor((rs.rand(n, 1) < 0.2).astype(np.float32))
    out.update({"disc/x": x.numpy(), "disc/a": a.numpy(), "disc/r": r.numpy(), "disc/target": tgt.numpy(),
                "disc/term": term.numpy(), "disc/cfg": np.array([O, A, n, E])})
    with torch.no_grad():
        for red in ("min", "max", "mean", "none", "mix"):
            out[f"disc/call/{red}"] = qf(x, red).numpy()
        out["disc/target/all_min"] = qf.compute_target(x).numpy()
        out["disc/target/picked_min"] = qf.compute_target(x, a).numpy()
        out["disc/target/picked_mix"] = qf.compute_target(x, a, "mix", 0.6).numpy()
        out["disc/error/float"] = qf.compute_error(x, a, r, tgt, term, 0.99).numpy()
    path = os.path.join(HERE, "qfunc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if "/call/" in k or "error" in k})


if __name__ == "__main__":
    main()
