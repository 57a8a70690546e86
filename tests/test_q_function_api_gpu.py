"""GPU: the callable Q-function API (`impl.q_function(x, action, reduction)`, `compute_target`, `compute_error`,
`q_funcs[i]`; d3rlpy_b200/q_functions.py) against outputs of the UNMODIFIED reference's EnsembleContinuousQFunction /
EnsembleDiscreteQFunction on copied weights (tests/golden/qfunc.npz, written by tests/golden/make_golden_qfunc.py;
the reference's own tests for these modules — tests/models/torch/q_functions/test_ensemble_q_function.py:24-41,
124-149,208-231 — check shapes and the reductions against a member loop, which the fixture subsumes).
Tolerance: 1e-5 relative in fp32 mode on both dense-layer engines, 1e-2 in bf16 mode."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qfunc.npz"))


def _sd(prefix):
    return OrderedDict((k[len(prefix):], torch.tensor(Z[k])) for k in Z.files if k.startswith(prefix))


def _close(got, ref, rel, what):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = float(np.abs(got - ref).max())
    assert err <= rel * max(1.0, float(np.abs(ref).max())), f"{what}: err {err:.3e}"


@pytest.mark.parametrize("precision,engine", [("fp32", 1), ("fp32", 0), ("bf16", 1)])
def test_continuous_ensemble_q_function_api(precision, engine):
    from d3rlpy_b200._lib import lib
    from d3rlpy_b200.algos import CQL

    lib().set_fp32_engine(engine)
    try:
        O, A, n, E = (int(v) for v in Z["cont/cfg"])
        rel = 1e-5 if precision == "fp32" else 1e-2
        algo = CQL(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], n_critics=E, precision=precision)
        algo.create_impl((O,), A)
        impl = algo.impl
        impl.q_function.load_state_dict(_sd("cont/q/"))
        qf = impl.q_function
        x, a = torch.tensor(Z["cont/x"]), torch.tensor(Z["cont/a"])
        for red in ("min", "max", "mean", "none", "mix"):
            _close(qf(x, a, red), Z[f"cont/call/{red}"], rel, f"call {red}")
        _close(qf(x, a), Z["cont/call/mean"], rel, "default reduction is mean")
        _close(qf.compute_target(x, a), Z["cont/target/min"], rel, "compute_target default (min)")
        _close(qf.compute_target(x, a, "mix", 0.6), Z["cont/target/mix"], rel, "compute_target mix lam 0.6")
        r, tgt, term = torch.tensor(Z["cont/r"]), torch.tensor(Z["cont/target"]), torch.tensor(Z["cont/term"])
        _close(qf.compute_error(x, a, r, tgt, term, 0.99), Z["cont/error/float"], rel, "compute_error float gamma")
        _close(qf.compute_error(x, a, r, tgt, term, torch.tensor(Z["cont/gamma_rows"])), Z["cont/error/rows"], rel,
               "compute_error per-row gamma")
        assert len(qf.q_funcs) == E
        for e, member in enumerate(qf.q_funcs):
            _close(member(x, a), Z[f"cont/member{e}"], rel, f"q_funcs[{e}]")
            keys = list(member.state_dict().keys())
            assert keys == ["_encoder._fcs.0.weight", "_encoder._fcs.0.bias", "_encoder._fcs.1.weight",
                            "_encoder._fcs.1.bias", "_fc.weight", "_fc.bias"], keys
        # the target network is an independent copy under the same API
        impl.targ_q_function.load_state_dict(_sd("cont/q/"))
        _close(impl.targ_q_function(x, a, "min"), Z["cont/call/min"], rel, "target network")
        with pytest.raises(ValueError):
            qf(x, a, "median")
        with pytest.raises(AssertionError):
            qf.compute_error(x, a, r, tgt.reshape(-1), term, 0.99)   # `assert target.ndim == 2`
        # inputs may live on the device already
        _close(qf(x.cuda(), a.cuda(), "min"), Z["cont/call/min"], rel, "device inputs")
    finally:
        lib().set_fp32_engine(1)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_discrete_ensemble_q_function_api(precision):
    from d3rlpy_b200.algos import DoubleDQN

    O, A, n, E = (int(v) for v in Z["disc/cfg"])
    rel = 1e-5 if precision == "fp32" else 1e-2
    algo = DoubleDQN(encoder_factory=[32, 32], n_critics=E, precision=precision)
    algo.create_impl((O,), A)
    impl = algo.impl
    impl.q_function.load_state_dict(_sd("disc/q/"))
    qf = impl.q_function
    x, a = torch.tensor(Z["disc/x"]), torch.tensor(Z["disc/a"])
    for red in ("min", "max", "mean", "none", "mix"):
        _close(qf(x, red), Z[f"disc/call/{red}"], rel, f"call {red}")
    _close(qf.compute_target(x), Z["disc/target/all_min"], rel, "compute_target(x)")
    _close(qf.compute_target(x, a), Z["disc/target/picked_min"], rel, "compute_target(x, action)")
    _close(qf.compute_target(x, a, "mix", 0.6), Z["disc/target/picked_mix"], rel, "compute_target mix")
    r, tgt, term = torch.tensor(Z["disc/r"]), torch.tensor(Z["disc/target"]), torch.tensor(Z["disc/term"])
    _close(qf.compute_error(x, a, r, tgt, term, 0.99), Z["disc/error/float"], rel, "compute_error (Huber)")
