"""Worker for the data-parallel equivalence tests (launched with torch.distributed.run, one process per
rank).  Mode "gpu": W ranks each run the CUDA CQL / TD3+BC update on their row shard of one global
minibatch (NCCL all-reduce of gradients and loss sums) and rank 0 checks metrics and post-step parameters
against the single-process CPU oracle on the full batch.  Mode "cpu": the same sharding helpers drive the
ORACLE over gloo — the host-side partitioning / noise slicing / loss scaling logic without a GPU."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from d3rlpy_b200 import parallel  # noqa: E402
from oracle import update as ou  # noqa: E402

CQL_KINDS = {"temp": "B*", "alpha_t": "NB*", "alpha_tp1": "NB*", "alpha_rand": "BN*", "critic_t": "NB*",
             "critic_tp1": "NB*", "critic_rand": "BN*", "actor": "B*"}
CQL_ORDER = ["temp", "alpha_t", "alpha_tp1", "alpha_rand", "critic_t", "critic_tp1", "critic_rand", "actor"]


def synthetic(rs, B, O, A):
    return dict(observations=rs.randn(B, O).astype(np.float32), actions=rs.uniform(-1, 1, (B, A)).astype(np.float32),
                rewards=rs.randn(B, 1).astype(np.float32), next_observations=rs.randn(B, O).astype(np.float32),
                terminals=(rs.rand(B, 1) < 0.1).astype(np.float32), n_steps=np.ones((B, 1), np.float32))


def shard_batch(arrays, world, rank):
    B = arrays["rewards"].shape[0]
    lo, hi = parallel.shard_rows(B, world, rank)
    return {k: v[lo:hi] for k, v in arrays.items()}


def close(a, b, rel=2e-5):
    return abs(a - b) <= rel * max(1.0, abs(b)) + 1e-6


def run_gpu(world, rank):
    from d3rlpy_b200.algos import CQL, TD3PlusBC

    torch.cuda.set_device(rank)
    O, A, B, N, H = 9, 3, 64, 4, [64, 64]
    # ---- CQL
    orc = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=11)
    algo = CQL(actor_encoder_factory=H, critic_encoder_factory=H, batch_size=B // world, n_action_samples=N,
               use_gpu=rank, world_size=world, rank=rank)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    rs = np.random.RandomState(0)
    ok = True
    for s in range(3):
        arrays = synthetic(rs, B, O, A)
        noise = ou.Noise(seed=20 + s)
        ref = orc.update(ou.Batch(arrays), noise)
        local = [parallel.shard_noise(t, CQL_KINDS[k], B, N, world, rank) for k, t in zip(CQL_ORDER, noise.log)]
        impl.inject_noise(local, B // world)
        m = algo.update(SimpleNamespace(**shard_batch(arrays, world, rank)))
        for k, v in ref.items():
            if not close(float(m[k]), v):
                ok = False
                print(f"[rank {rank}] cql step {s} {k}: {float(m[k])} vs {v}", flush=True)
    for name, view, refp in (("q", impl.q_function, orc.q), ("pi", impl.policy, orc.pi),
                             ("targ_q", impl.targ_q_function, orc.targ_q)):
        for k, v in refp.items():
            g = view.state_dict()[k].cpu()
            err = float((g - v.detach()).abs().max())
            if err > 2e-5 * max(1.0, float(v.abs().max())):
                ok = False
                print(f"[rank {rank}] cql {name}/{k} err {err}", flush=True)
    # ---- CQL, tensor-core mode: the fused program (one row-assembly kernel, fused forward / backward launches) with
    # the loss tails split around the all-reduces; tolerance 1e-2 (bf16 operands)
    O2, A2, B2, N2, H2 = 17, 6, 128, 6, [64, 64, 64]
    orc = ou.CQL(O2, A2, hidden=H2, n_action_samples=N2, seed=13)
    algo = CQL(actor_encoder_factory=H2, critic_encoder_factory=H2, batch_size=B2 // world, n_action_samples=N2,
               use_gpu=rank, world_size=world, rank=rank, precision="bf16")
    algo.create_impl((O2,), A2)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    for s in range(3):
        arrays = synthetic(rs, B2, O2, A2)
        noise = ou.Noise(seed=60 + s)
        ref = orc.update(ou.Batch(arrays), noise)
        local = [parallel.shard_noise(t, CQL_KINDS[k], B2, N2, world, rank) for k, t in zip(CQL_ORDER, noise.log)]
        impl.inject_noise(local, B2 // world)
        m = algo.update(SimpleNamespace(**shard_batch(arrays, world, rank)))
        for k, v in ref.items():
            if not close(float(m[k]), v, 1e-2):
                ok = False
                print(f"[rank {rank}] cql bf16 step {s} {k}: {float(m[k])} vs {v}", flush=True)
    for k, v in orc.q.items():
        g = impl.q_function.state_dict()[k].cpu()
        if float((g - v.detach()).abs().max()) > 1e-2 * max(1.0, float(v.abs().max())):
            ok = False
            print(f"[rank {rank}] cql bf16 q/{k} mismatch", flush=True)
    # ---- TD3+BC (the actor's lambda = alpha / mean|Q| needs the GLOBAL mean)
    orc = ou.TD3PlusBC(O, A, hidden=H, seed=12)
    algo = TD3PlusBC(actor_encoder_factory=H, critic_encoder_factory=H, batch_size=B // world, scaler=None,
                     use_gpu=rank, world_size=world, rank=rank)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    for s in range(2):
        arrays = synthetic(rs, B, O, A)
        noise = ou.Noise(seed=40 + s)
        ref = orc.update(ou.Batch(arrays), noise)
        impl.inject_noise([parallel.shard_noise(noise.log[0], "B*", B, 1, world, rank)], B // world)
        m = algo.update(SimpleNamespace(**shard_batch(arrays, world, rank)))
        for k, v in ref.items():
            if not close(float(m[k]), v):
                ok = False
                print(f"[rank {rank}] td3bc step {s} {k}: {float(m[k])} vs {v}", flush=True)
    for k, v in orc.pi.items():
        g = impl.policy.state_dict()[k].cpu()
        if float((g - v.detach()).abs().max()) > 2e-5 * max(1.0, float(v.abs().max())):
            ok = False
            print(f"[rank {rank}] td3bc pi/{k} mismatch", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{rank}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    return bool(flag.item() == 1.0)


def run_cpu(world, rank):
    """Sharded ORACLE over gloo: per-rank losses scaled by 1/W, gradients summed, must equal the
    full-batch oracle gradients (every loss is a batch mean)."""
    O, A, B, N, H = 7, 3, 32, 3, [16, 16]
    full = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=5)
    mine = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=5)
    rs = np.random.RandomState(3)
    arrays = synthetic(rs, B, O, A)
    noise = ou.Noise(seed=9)
    # full-batch critic gradient
    bf = ou.Batch(arrays)
    q_tpn = full.compute_target(bf, noise)
    n0 = len(noise.log)
    loss = full.compute_critic_loss(bf, q_tpn, noise)
    draws = noise.log[n0:]
    for p in full.q.values():
        p.grad = None
    loss.backward()
    ref = {k: p.grad.clone() for k, p in full.q.items()}
    # my shard
    lo, hi = parallel.shard_rows(B, world, rank)
    bl = ou.Batch(shard_batch(arrays, world, rank))
    kinds = ["NB*", "NB*", "BN*"]
    local = ou.Noise(injected=[parallel.shard_noise(t, k, B, N, world, rank) for t, k in zip(draws, kinds)])
    l = mine.compute_critic_loss(bl, q_tpn[lo:hi], local) / world
    for p in mine.q.values():
        p.grad = None
    l.backward()
    ok = True
    for k, p in mine.q.items():
        g = p.grad.clone()
        dist.all_reduce(g)
        if not torch.allclose(g, ref[k], rtol=1e-4, atol=1e-6):
            ok = False
            print(f"[rank {rank}] {k} max err {float((g - ref[k]).abs().max())}", flush=True)
    tl = l.detach().clone()
    dist.all_reduce(tl)
    ok = ok and close(float(tl), float(loss.detach()), 1e-5)
    return ok


if __name__ == "__main__":
    mode = sys.argv[1]
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    dist.init_process_group("gloo" if mode == "cpu" else "nccl",
                            **({"device_id": torch.device("cuda", rank)} if mode == "gpu" else {}))
    torch.set_num_threads(2)
    ok = run_cpu(world, rank) if mode == "cpu" else run_gpu(world, rank)
    if rank == 0:
        print("DP_EQUIVALENCE_OK" if ok else "DP_EQUIVALENCE_FAILED", flush=True)
    sys.stdout.flush()
    # hard exit: NCCL communicators captured in CUDA graphs are torn down with the process
    os._exit(0 if ok else 1)
