#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the offline-RL update path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c5] [--impl ours|reference] [--precision bf16|fp32]

Metric (BASELINE.json): CQL / TD3+BC gradient updates/sec at batch 256.  The headline line is config c2 (CQL, obs 17,
act 6, B 256, N 10, 2 critics, 3x256 MLPs, synthetic halfcheetah-shaped data); one "step" = one `algo.update(batch)`
(temp -> alpha -> critic -> actor -> target sync) on one minibatch.

  value : steps/s with the replay buffer and the step's indices already in HBM — per step one gather kernel launch + one
          CUDA-graph launch, timed with CUDA events on the launching stream, L2 flushed between timed steps (and, for
          N > 1, a barrier after the flush so that every rank's timed region starts together).
  e2e   : the same metric through the reference-facing call `algo.update(TransitionMiniBatch-like numpy batch)`: pinned
          H2D of the six arrays + graph + pinned D2H of the metrics, every step.
  fp32_parity_mode : the same c2 workload in `precision="fp32"` (3xTF32 tensor-core GEMMs, fp32 everything else): the mode
          that meets the 1e-5 parity tolerance; its own value / e2e / dtype.
  extra.c1         : config c1 (TD3+BC, obs 11, act 3, B 256, 2 critics, 256x256), value and e2e, both precisions.
  extra.c5_strong  : config c5 (CQL, obs 111, act 8, B 8192, 10 critics) with the minibatch SHARDED over the N ranks
          (strong scaling, batch-8192 updates/s) — the north-star scaling configuration, at every N.
  hbm              : achieved GB/s of the HBM-bound kernels (gather, Adam + Polyak + shadow refresh, soft_sync) at the
          sizes of this configuration and at a bandwidth-regime size, measured live.
  dp_check (N > 1) : sharded update == single-GPU update on the same global batch (metrics, parameters), pass / fail.
  roofline / cpu_baseline : see DESIGN.md §Measurement.

`--impl reference` times the CPU restatement of the reference update (oracle/update.py — plain PyTorch fp32 + autograd
+ torch.optim.Adam, i.e. what the reference executes with use_gpu=False) on the host cores, on the same config/metric.
"""
import argparse
import json
import os
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: obs, act, batch, n_action_samples, n_critics, hidden, algo
    "c1": dict(algo="td3bc", obs=11, act=3, batch=256, n=0, critics=2, hidden=[256, 256],
               desc="TD3+BC hopper-shaped (obs 11, act 3), batch 256, 2 critics, 256x256 MLP"),
    "c2": dict(algo="cql", obs=17, act=6, batch=256, n=10, critics=2, hidden=[256, 256, 256],
               desc="CQL halfcheetah-shaped (obs 17, act 6), batch 256, n_action_samples 10, 2 critics, 3x256 MLP"),
    "c5": dict(algo="cql", obs=111, act=8, batch=8192, n=10, critics=10, hidden=[256, 256, 256],
               desc="CQL ant-shaped (obs 111, act 8), batch 8192, n_action_samples 10, 10 critics, 3x256 MLP"),
}
METRIC = "CQL gradient updates/sec at batch 256"
# keys of `config`, identical on both arms (the driver compares the two lines' configs)
CONFIG_KEYS = ("workload", "per_gpu_batch", "precision", "global_batch", "units", "parallelism", "l2", "timing")
CQL_NOISE_KINDS = {"temp": "B*", "alpha_t": "NB*", "alpha_tp1": "NB*", "alpha_rand": "BN*", "soft": "B*",
                   "critic_t": "NB*", "critic_tp1": "NB*", "critic_rand": "BN*", "actor": "B*", "target": "B*"}


def req_gemm_flops(w) -> float:
    """ALGORITHMIC (strictly required) GEMM FLOPs per update — SURVEY.md §8(d) 'req', DESIGN.md."""
    O, A, B, N, E, H = w["obs"], w["act"], w["batch"], w["n"], w["critics"], w["hidden"]

    def mlp(in_dim, out):
        dims = [in_dim] + H
        return sum(a * b for a, b in zip(dims[:-1], dims[1:])) + H[-1] * out

    if w["algo"] == "cql":
        c, p = mlp(O + A, 1), mlp(O, 2 * A)
        R = B * (1 + 3 * N)
        f = 2 * R * c * E            # alpha step: forward only
        f += 3 * 2 * R * c * E       # critic step: forward + dgrad + wgrad
        f += 2 * B * c * E           # target critics
        f += 2 * 2 * B * c * E       # actor step: critic forward + dgrad
        f += 2 * 2 * B * p           # policy forward on [obs; next_obs]
        f += 2 * 2 * B * p           # policy backward (dgrad + wgrad) on B rows
        return float(f)
    c, p = mlp(O + A, 1), mlp(O, A)
    f = 2 * B * c * E + 3 * 2 * B * c * E + 2 * B * p          # target (policy+critics), critic fwd+bwd
    f += 0.5 * (3 * 2 * B * p + 2 * 2 * B * c)                 # actor every 2nd step (member 0 only)
    return float(f)


def make_dataset(w, steps_total=1_000_000, seed=0):
    """BASELINE.md synthetic data: N(0,1) observations, U(-1,1) actions, N(0,1) rewards, episodes of 1000."""
    rs = np.random.RandomState(seed)
    S = steps_total
    obs = rs.randn(S, w["obs"]).astype(np.float32)
    act = rs.uniform(-1, 1, (S, w["act"])).astype(np.float32)
    rew = rs.randn(S).astype(np.float32)
    term = np.zeros(S, np.float32)
    term[999::1000] = 1.0
    return obs, act, rew, term


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (NVML)."""

    def __init__(self, index=0, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self._stop_evt = index, period, [], set(), threading.Event()
        self.max_mhz = None
        self.ready = threading.Event()  # NVML initialised: the timed region (tens of ms) starts only after this

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            self.ready.set()
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------- reference arm
def build_oracle(w, seed=0):
    from oracle import update as ou

    if w["algo"] == "cql":
        return ou.CQL(w["obs"], w["act"], hidden=w["hidden"], n_critics=w["critics"], n_action_samples=w["n"], seed=seed)
    return ou.TD3PlusBC(w["obs"], w["act"], hidden=w["hidden"], n_critics=w["critics"], seed=seed)


def time_oracle(w, batches, threads, steps, warmup, budget_s=None, warm_budget_s=30.0):
    """(updates/s, timed updates done, seconds, warm-up updates done); both loops stop early at their time budget."""
    import torch

    from oracle import update as ou

    torch.set_num_threads(threads)
    orc = build_oracle(w)
    noise = ou.Noise(seed=0)
    t0 = time.perf_counter()
    warmed = 0
    for i in range(warmup):
        orc.update(ou.Batch(batches[i % len(batches)]), noise)
        warmed += 1
        if time.perf_counter() - t0 > warm_budget_s:
            break
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        orc.update(ou.Batch(batches[i % len(batches)]), noise)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt, warmed


def host_batches(w, n_batches, obs, act, rew, term):
    """Minibatches as the reference's sampler would build them (numpy arrays), via the oracle sampler."""
    from oracle import sampler as osampler

    small = 20_000  # the oracle's flat metadata build is a Python loop; a slice of the data is enough
    replay = osampler.FlatReplay(obs[:small], act[:small], rew[:small], term[:small])
    rs = np.random.RandomState(1)
    return [osampler.gather(replay, rs.randint(len(replay), size=w["batch"])) for _ in range(n_batches)]


def thread_sweep(cores):
    return sorted({1, max(1, cores // 2), cores})


def run_reference(args, w):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    obs, act, rew, term = make_dataset(w, steps_total=50_000)
    batches = host_batches(w, 4, obs, act, rew, term)
    cores = os.cpu_count() or 1
    # the reference arm gets its best thread count: short probes at {1, cores/2, cores}, full run with the fastest
    probes = {th: time_oracle(w, batches, th, 4, 1, budget_s=15.0)[0] for th in thread_sweep(cores)}
    threads = max(probes, key=probes.get)
    rate, done, dt, warmed = time_oracle(w, batches, threads, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": done, "warmup": warmed, "ms_per_step": 1e3 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same keys as the GPU arm's `config` (same workload, this arm's own values)
        "config": {"workload": w["desc"], "per_gpu_batch": w["batch"],
                   "precision": "fp32 (plain PyTorch CPU kernels of the oracle port)", "global_batch": w["batch"],
                   "units": "updates of 256-transition minibatches per second",
                   "parallelism": f"host CPU, {threads} torch threads (rank 0 only)", "l2": "n/a (CPU)",
                   "timing": "time.perf_counter around the timed updates"},
        "cpu_baseline": {"value": rate, "unit": "updates/s", "cores": threads, "host_cores": cores, "kind": "port",
                         "sample": f"{done} full updates of the oracle port (plain PyTorch fp32 CPU) in {dt:.1f}s; torch "
                                   f"threads swept over {thread_sweep(cores)} of {cores} host cores "
                                   f"({ {k: round(v, 2) for k, v in probes.items()} } updates/s), best = {threads}"},
        "e2e": {"value": rate, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "torch_threads": torch.get_num_threads(),
    }
    assert tuple(line["config"]) == CONFIG_KEYS
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- our arm
def build_algo(w, world_size=1, rank=0, precision="bf16", batch=None, **extra):
    from d3rlpy_b200.algos import CQL, TD3PlusBC

    kw = dict(world_size=world_size, rank=rank) if world_size > 1 else {}
    kw["precision"] = precision
    kw.update(extra)
    B = batch or w["batch"]
    if w["algo"] == "cql":
        algo = CQL(actor_encoder_factory=w["hidden"], critic_encoder_factory=w["hidden"], batch_size=B,
                   n_action_samples=w["n"], n_critics=w["critics"], use_gpu=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    else:
        algo = TD3PlusBC(actor_encoder_factory=w["hidden"], critic_encoder_factory=w["hidden"], batch_size=B,
                         n_critics=w["critics"], scaler=None, use_gpu=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    algo.create_impl((w["obs"],), w["act"])
    return algo


def kernel_profile(algo, batch_np, n_iter=5):
    """Eager (no graph) instrumented passes: CUDA events on the launching stream around every C-ABI
    launch; returns per-family totals for one update (averaged over n_iter)."""
    import torch

    impl = algo.impl
    L = impl._lib
    saved = impl.use_graph
    impl.use_graph = False
    algo.update(batch_np)  # warm
    L.start_profile(impl._stream_obj)
    for _ in range(n_iter):
        # a 3 ms busy-wait kernel first: the host enqueues the whole update behind it, so the per-launch event
        # durations below are device execution times, not host launch latency
        L.raw("d3b_spin")(3_000_000, impl._stream)
        algo.update(batch_np)
    recs = L.stop_profile()
    impl.use_graph = saved
    torch.cuda.synchronize()
    fam = {}
    per_launch = {}
    for name, flops, ms in recs:
        f = fam.setdefault(name, [0, 0.0, 0.0])
        f[0] += 1
        f[1] += ms
        f[2] += flops
        if flops > 0:
            d = per_launch.setdefault((name, flops), [0, 0.0])
            d[0] += 1
            d[1] += ms
    fams = {k: {"launches_per_update": v[0] / n_iter, "us_per_update": 1e3 * v[1] / n_iter,
                "gflop_per_update": v[2] / n_iter / 1e9} for k, v in fam.items()}
    # the single launch with the most algorithmic FLOPs (the dominant kernel of the update)
    dominant = None
    if per_launch:
        (name, flops), (n, ms) = max(per_launch.items(), key=lambda kv: kv[0][1])
        dominant = {"name": name, "gflop_per_launch": flops / 1e9, "us_per_launch": 1e3 * ms / n, "launches_timed": n}
    return fams, dominant


def other_configs_e2e(precision, n=60, warm=10):
    """BASELINE configs[2] (BCQ, 750x750 VAE + 400x300 nets, n_action_samples 100, batch 256) and configs[3]
    (DiscreteCQL, NatureDQN encoder on 4x84x84 uint8 frames, batch 32): `algo.update(numpy batch)` wall clock per update
    (pinned H2D + one CUDA-graph replay + metric D2H), both modes."""
    import torch

    from d3rlpy_b200.algos import BCQ, DiscreteCQL

    rs = np.random.RandomState(0)
    out = {"how": "wall clock over algo.update(host numpy batch) calls, 4 rotating batches; us per update"}
    vec = [SimpleNamespace(observations=rs.randn(256, 17).astype(np.float32),
                           actions=rs.uniform(-1, 1, (256, 6)).astype(np.float32),
                           rewards=rs.randn(256, 1).astype(np.float32),
                           next_observations=rs.randn(256, 17).astype(np.float32),
                           terminals=(rs.rand(256, 1) < 0.01).astype(np.float32),
                           n_steps=np.ones((256, 1), np.float32)) for _ in range(4)]
    pix = [SimpleNamespace(observations=rs.randint(0, 256, (32, 4, 84, 84)).astype(np.uint8),
                           actions=rs.randint(0, 4, 32).astype(np.int32),
                           rewards=(rs.rand(32, 1) < 0.1).astype(np.float32),
                           next_observations=rs.randint(0, 256, (32, 4, 84, 84)).astype(np.uint8),
                           terminals=(rs.rand(32, 1) < 0.01).astype(np.float32),
                           n_steps=np.ones((32, 1), np.float32)) for _ in range(4)]
    for prec in dict.fromkeys([precision, "fp32"]):
        for name, make, batches, shape, act in (
                ("c3_bcq_b256", lambda: BCQ(actor_encoder_factory=[400, 300], critic_encoder_factory=[400, 300],
                                            imitator_encoder_factory=[750, 750], batch_size=256, n_action_samples=100,
                                            precision=prec), vec, (17,), 6),
                ("c4_discrete_cql_pixels_b32", lambda: DiscreteCQL(batch_size=32, n_frames=4, scaler="pixel",
                                                                   precision=prec), pix, (4, 84, 84), 4)):
            algo = make()
            algo.create_impl(shape, act)
            for i in range(warm):
                algo.update(batches[i % 4])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n):
                algo.update(batches[i % 4])
            torch.cuda.synchronize()
            us = (time.perf_counter() - t0) / n * 1e6
            out[f"{name}_{prec}"] = {"us_per_update": us, "updates_per_s": 1e6 / us}
            del algo
            torch.cuda.empty_cache()
    return out


class Runner:
    """One workload on this rank: replay + indices in HBM, device-timed steps, end-to-end steps."""

    def __init__(self, w, world, rank, local, precision, strong, data=None, **algo_kw):
        import torch

        from d3rlpy_b200.dataset import MDPDataset

        self.w, self.world, self.rank, self.precision, self.strong = w, world, rank, precision, strong
        self.dev = torch.device("cuda", local)
        assert not strong or w["batch"] % world == 0
        self.B = w["batch"] // world if strong else w["batch"]
        self.unit_scale = 1 if strong else world
        self.algo = build_algo(w, world, rank, precision, batch=self.B, **algo_kw)
        self.impl = self.algo.impl
        self.data = data or make_dataset(w)
        obs, act, rew, term = self.data
        self.replay = MDPDataset(obs, act, rew, term).device_replay(self.dev)
        self.db = self.impl.device_batch(self.B)
        self.holder = SimpleNamespace(_device_batch=self.db)

    def indices(self, n):
        import torch

        B, world, rank = self.B, self.world, self.rank
        if self.strong:  # every rank draws the same global index vector and takes its own row shard
            rs = np.random.RandomState(1)
            idx = rs.randint(len(self.replay), size=(n, B * world))[:, rank * B:(rank + 1) * B]
        else:
            rs = np.random.RandomState(1 + rank)
            idx = rs.randint(len(self.replay), size=(n, B))
        return torch.from_numpy(np.ascontiguousarray(idx.astype(np.int64))).to(self.dev)

    def gather(self, idx_row, stream=None):
        r, db, w = self.replay, self.db, self.w
        self.impl._lib.gather_vector(r.obs.data_ptr(), w["obs"], r.actions.data_ptr(), w["act"], 0, r.rewards.data_ptr(),
                                     r.meta.data_ptr(), idx_row.data_ptr(), self.B, 1, 0.99, db.ptr("obs"), db.ptr("act"),
                                     db.ptr("rew"), db.ptr("next_obs"), db.ptr("term"), db.ptr("nsteps"), None, None, 0.0,
                                     stream if stream is not None else self.impl._stream)

    def step_device(self, idx_row):
        self.gather(idx_row)
        self.algo._update_async(self.holder)
        self.algo._grad_step += 1

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def timed(self, K, W, flush_buf):
        """W warm-up + K timed steps; returns (max-over-ranks total ms, per-step ms of this rank, launches)."""
        import torch
        import torch.distributed as dist

        impl, L = self.impl, self.impl._lib
        idx = self.indices(K + W)
        for i in range(W):
            self.step_device(idx[i])
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        n0 = L.launch_count()
        with torch.cuda.stream(impl._stream_obj):
            for i in range(K):
                if flush_buf is not None:
                    flush_buf.fill_(float(i))
                    if self.world > 1:
                        dist.barrier()   # every rank's timed region starts after everybody's flush (no rank skew in `value`)
                evs[i][0].record(impl._stream_obj)
                self.step_device(idx[W + i])
                evs[i][1].record(impl._stream_obj)
        self.barrier()
        eager = L.launch_count() - n0
        step_ms = np.array([a.elapsed_time(b) for a, b in evs])
        total_ms = float(step_ms.sum())
        if self.world > 1:
            t = torch.tensor([total_ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        nodes = max(impl._graph_nodes.values()) if impl._graph_nodes else 0
        return total_ms, step_ms, int(eager + K * nodes), nodes

    def e2e(self, n_steps):
        """`algo.update(numpy batch)`: pinned H2D + graph + pinned D2H every step, wall clock, max over ranks."""
        import torch
        import torch.distributed as dist

        obs, act, rew, term = self.data
        hb = host_batches(dict(self.w, batch=self.B), 8, obs, act, rew, term)
        hbs = [SimpleNamespace(**b) for b in hb]
        for i in range(3):
            self.algo.update(hbs[i % len(hbs)])
        self.barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            m = self.algo.update(hbs[i % len(hbs)])
        torch.cuda.synchronize(self.dev)
        dt = time.perf_counter() - t0
        if self.world > 1:
            t = torch.tensor([dt], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert all(np.isfinite(float(v)) for v in m.values()), m
        return {"value": self.unit_scale * n_steps / dt, "unit": "updates/s", "h2d_bytes_per_step": self.impl._batch.h2d_bytes,
                "d2h_bytes_per_step": 4 * 64, "steps": n_steps, "ms_per_step": 1e3 * dt / n_steps}, hb, hbs

    def plane(self):
        """Which exchange carried the gradients of the data-parallel update."""
        if self.world == 1:
            return "single"
        px = getattr(self.impl, "_px", None)
        return f"dp{self.world}:" + ("peer (CUDA-IPC NVLink loads fused into the Adam kernel, no NCCL call on the update path)"
                                     if px is not None else "nccl (ncclAllReduce per optimizer group)")


def summarize(r, total_ms, step_ms, K):
    return {"value": r.unit_scale * K / (total_ms * 1e-3), "unit": "updates/s", "ms_per_step": total_ms / K,
            "step_ms_p10_p50_p90": [float(np.percentile(step_ms, p)) for p in (10, 50, 90)]}


def graph_time_us(fn, n=20, reps=5):
    """us per call of `fn(stream)` with n calls captured in one CUDA graph (device time, no host launch gaps)."""
    import torch

    for _ in range(2):
        fn(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(n):
            fn(st)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * n) * 1e3


def hbm_block(r, peaks):
    """Achieved GB/s of the HBM-bound kernels, measured live (n launches per CUDA graph, CUDA events): at the sizes of
    THIS configuration (a 256-row batch / a 0.4 M-parameter arena are single partial waves: latency-, not
    bandwidth-bound) and at a bandwidth-regime size of the same kernels."""
    import torch

    impl, w, L = r.impl, r.w, r.impl._lib
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
           "how": "20 launches per CUDA graph, 5 replays, CUDA events; bytes = algorithmic (read + write)"}
    O, A = w["obs"], w["act"]

    def add(name, nbytes, us, note):
        out[name] = {"bytes": int(nbytes), "us": us, "gbs": nbytes / us / 1e3, "frac": nbytes / us / 1e3 / peak, "note": note}

    idx = r.indices(2)
    row_bytes = 4 * (2 * O + A + 3)
    us = graph_time_us(lambda st: r.gather(idx[0], st))
    add("gather_config", 2 * r.B * row_bytes, us, f"gather_vector, {r.B} rows x {row_bytes} B (read + write)")
    # bandwidth regime: 64 Ki rows
    big = 65536
    from d3rlpy_b200.algos.torch.base import DeviceBatch

    dbig = DeviceBatch(big, O, A, r.dev)
    ibig = torch.from_numpy(np.random.RandomState(3).randint(len(r.replay), size=big).astype(np.int64)).to(r.dev)
    rp = r.replay
    us = graph_time_us(lambda st: L.gather_vector(rp.obs.data_ptr(), O, rp.actions.data_ptr(), A, 0, rp.rewards.data_ptr(),
                                                   rp.meta.data_ptr(), ibig.data_ptr(), big, 1, 0.99, dbig.ptr("obs"),
                                                   dbig.ptr("act"), dbig.ptr("rew"), dbig.ptr("next_obs"), dbig.ptr("term"),
                                                   dbig.ptr("nsteps"), None, None, 0.0, st))
    add("gather_64k_rows", 2 * big * row_bytes, us, "same kernel, 65 536 rows")
    for name, net in (("critic", impl._q_func), ("policy", impl._policy)):
        a = net.arena
        shadow = 2 * (net.shadow.numel() + (net.shadow_target.numel() if net.shadow_target is not None else 0)) \
            if r.precision == "bf16" else 0
        nbytes = a.size * (28 + 8) + shadow   # p, g, m, v read + p, m, v written (28 B) + target read / write (8 B) + bf16 shadows
        us = graph_time_us(lambda st, net=net: net.adam(3e-4, st, tau=0.005))
        add(f"adam_polyak_{name}", nbytes, us, f"fused Adam + Polyak + grad zeroing (+ bf16 shadow refresh), {a.size} parameters")
        us = graph_time_us(lambda st, a=a: L.soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, 0.005, st))
        add(f"soft_sync_{name}", a.size * 12, us, f"standalone soft_sync, {a.size} parameters")
    n = 64 * 1024 * 1024
    p, g, m, v, t = (torch.zeros(n, device=r.dev) for _ in range(5))
    step = torch.ones(1, dtype=torch.int32, device=r.dev)
    us = graph_time_us(lambda st: L.adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), t.data_ptr(), n,
                                               step.data_ptr(), 3e-4, 0.9, 0.999, 1e-8, 0.005, 1, st), n=5, reps=3)
    add("adam_polyak_64M", n * 36, us, "same kernel family (adam_step + Polyak), 64 Mi parameters")
    return out


def dp_check(w, world, rank, local, precision):
    """Sharded update == single-GPU update on the same global batch: every rank runs the W-way data-parallel update on
    its row shard and, on its own GPU, the plain update on the whole W x B batch from identical weights and injected
    noise; metrics and post-step parameters must agree (summation order is the only difference)."""
    import torch

    from d3rlpy_b200 import parallel

    B = w["batch"]
    Bg = B * world
    dp = build_algo(w, world, rank, precision, batch=B)
    one = build_algo(w, 1, 0, precision, batch=Bg)
    for a, b in ((one.impl.q_function, dp.impl.q_function), (one.impl.targ_q_function, dp.impl.targ_q_function),
                 (one.impl.policy, dp.impl.policy), (one.impl.targ_policy, dp.impl.targ_policy)):
        a.load_state_dict(b.state_dict())
    init = {k: v.clone() for k, v in dp.impl.q_function.state_dict().items()}
    rs = np.random.RandomState(123)
    worst_m, ok = 0.0, True
    for s in range(2):
        arrays = dict(observations=rs.randn(Bg, w["obs"]).astype(np.float32),
                      actions=rs.uniform(-1, 1, (Bg, w["act"])).astype(np.float32), rewards=rs.randn(Bg, 1).astype(np.float32),
                      next_observations=rs.randn(Bg, w["obs"]).astype(np.float32),
                      terminals=(rs.rand(Bg, 1) < 0.05).astype(np.float32), n_steps=np.ones((Bg, 1), np.float32))
        gen = torch.Generator().manual_seed(1000 + s)
        layout = one.impl.noise_layout(Bg)
        full, shard = [], []
        for name, (kind, shape) in layout.items():
            t = torch.randn(*shape, generator=gen) if kind == "normal" else torch.rand(*shape, generator=gen) * 2 - 1
            full.append(t)
            shard.append(parallel.shard_noise(t, CQL_NOISE_KINDS[name], Bg, w["n"], world, rank))
        one.impl.inject_noise(full, Bg)
        dp.impl.inject_noise(shard, B)
        m1 = one.update(SimpleNamespace(**arrays))
        lo, hi = parallel.shard_rows(Bg, world, rank)
        mw = dp.update(SimpleNamespace(**{k: v[lo:hi] for k, v in arrays.items()}))
        for k in m1:
            worst_m = max(worst_m, abs(float(mw[k]) - float(m1[k])) / max(1.0, abs(float(m1[k]))))
    num = den = 0.0
    sd1, sdw = one.impl.q_function.state_dict(), dp.impl.q_function.state_dict()
    for k in sd1:
        num += float(((sdw[k] - init[k]) - (sd1[k] - init[k])).double().pow(2).sum())
        den += float((sd1[k] - init[k]).double().pow(2).sum())
    rel = (num / max(den, 1e-300)) ** 0.5
    tol_m, tol_p = (1e-4, 2e-2) if precision == "bf16" else (2e-5, 2e-2)
    ok = worst_m <= tol_m and rel <= tol_p
    return {"pass": bool(ok), "world": world, "global_batch": Bg, "steps": 2, "max_metric_rel_diff": worst_m,
            "critic_update_rel_l2_diff": rel, "tolerances": {"metric_rel": tol_m, "update_rel_l2": tol_p},
            "what": "W-way sharded update vs the same update on the whole global batch on one GPU (identical weights, "
                    "injected noise); only the summation order differs"}


def trace(msg):
    if os.environ.get("BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    K, W = args.steps, args.warmup
    strong = args.workload == "c5"
    flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass

    # ---- headline workload
    r = Runner(w, world, rank, local, args.precision, strong)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(5.0)
    total_ms, step_ms, launches, graph_nodes = r.timed(K, W, flush_buf)
    clocks = sampler.stop()
    head = summarize(r, total_ms, step_ms, K)
    trace(f"headline timed: {head['ms_per_step']:.4f} ms/step")
    # the same device-timed region WITHOUT the L2 flush (and, for N > 1, without the per-step barrier behind it): the
    # figure `e2e` is comparable with -- its public-API loop does not flush either, so e2e <= value_l2_warm always,
    # while `value` (cold L2 every step) can fall below `e2e` when the cold misses cost more than the host path
    kw = max(20, K // 5)
    tw, sw, _, _ = r.timed(kw, 3, None)
    warm = summarize(r, tw, sw, kw)
    warm["steps"] = kw
    trace(f"warm-L2 timed: {warm['ms_per_step']:.4f} ms/step")
    e2e, hb, hbs = r.e2e(max(10, min(K, 200)))
    plane = r.plane()
    trace("e2e done")

    extra = {}
    fp32_mode = None
    roof = cpu_baseline = hbm = variant = check = None
    if not args.headline_only:
        # ---- the parity mode of the same workload: precision="fp32" on the 3xTF32 tensor-core engine
        if args.precision == "bf16" and world == 1 and not strong:
            r32 = Runner(w, 1, 0, local, "fp32", False, data=r.data)
            k32 = max(20, K // 3)
            t32, s32, l32, n32 = r32.timed(k32, max(3, W // 2), flush_buf)
            fp32_mode = summarize(r32, t32, s32, k32)
            fp32_mode.update({"dtype": "tf32x3 (fp32 operands split hi + lo, three tcgen05 kind::tf32 MMAs per step, fp32 "
                                       "accumulate and fp32 everything else)", "parity": "1e-5 (tests/test_update_gpu.py)",
                              "steps": k32, "graph_nodes_per_update": n32, "e2e": r32.e2e(max(10, min(k32, 100)))[0]})
            del r32
        # ---- c1 (TD3+BC) next to the CQL headline: BASELINE.json's metric names both
        if world == 1 and args.workload == "c2":
            c1 = {}
            for prec in ("bf16", "fp32"):
                rc = Runner(WORKLOADS["c1"], 1, 0, local, prec, False)
                kc = max(20, K // 2)
                tc, sc, lc, nc = rc.timed(kc, max(3, W // 2), flush_buf)
                c1[prec] = summarize(rc, tc, sc, kc)
                c1[prec].update({"steps": kc, "graph_nodes_per_update": nc, "e2e": rc.e2e(max(10, min(kc, 100)))[0]})
                del rc
            extra["c1"] = {"workload": WORKLOADS["c1"]["desc"], "metric": "TD3+BC gradient updates/sec at batch 256", **c1}
        # ---- c3 (BCQ) and c4 (DiscreteCQL on pixels): end to end through the public API (host batch in, metrics out)
        if world == 1 and args.workload == "c2":
            extra["c3_c4_e2e"] = other_configs_e2e(args.precision)
        # ---- c5 strong scaling (the north-star scaling configuration) at every N
        if args.workload == "c2" and WORKLOADS["c5"]["batch"] % world == 0:
            r5 = Runner(WORKLOADS["c5"], world, rank, local, args.precision, True)
            k5 = max(10, min(K // 6, 50))
            t5, s5, l5, n5 = r5.timed(k5, 3, flush_buf)
            extra["c5_strong"] = summarize(r5, t5, s5, k5)
            extra["c5_strong"].update({"workload": WORKLOADS["c5"]["desc"], "scaling": "strong", "n_gpus": world,
                                       "per_gpu_batch": r5.B, "steps": k5, "parallelism": r5.plane(),
                                       "unit": "batch-8192 updates/s (minibatch sharded over the ranks)"})
            del r5
            trace("c5 done")
        if world > 1 and w["algo"] == "cql":
            check = dp_check(w, world, rank, local, args.precision)
            trace("dp_check done")
        if world > 1 and args.profile_dp:
            # per-launch device times of the data-parallel update (every rank runs the instrumented pass: the update is
            # collective); kernels that wait for a peer include the wait
            prof_dp, _ = kernel_profile(r.algo, hbs[0], n_iter=10)
            extra["dp_families"] = prof_dp

    if rank == 0 and world == 1 and not args.headline_only:
        # ---- roofline of the dominant kernel family (dense layers), measured live with CUDA events
        prof, dom = kernel_profile(r.algo, hbs[0])
        tc = [v for k, v in prof.items() if k.startswith("linear_") or k.startswith("umma_gemm") or k.startswith("mlp_")]
        tc_us = sum(v["us_per_update"] for v in tc)
        all_us = sum(v["us_per_update"] for v in prof.values())
        # the dominant kernel is timed alone (events around the single launch) -> burst peak
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        flops = req_gemm_flops(w)
        achieved = dom["gflop_per_launch"] / (dom["us_per_launch"] * 1e-6) / 1e3 if dom else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_dominant.json")))["dram_bytes_per_launch"]
        except Exception:  # noqa: BLE001
            pass
        roof = {"bound": "tensor",
                "kernel": (f"{dom['name']} (largest launch: fused critic trunk+head, tcgen05.mma + TMA + TMEM, "
                           "all members, the 7 936 importance-sampling rows of the critic step; the alpha step's rows are an identical "
                           "launch on a concurrent graph branch)")
                if args.precision == "bf16" and dom else "tc32_gemm_kernel (3xTF32 tcgen05)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks
                else "fallback 1.59 PFLOP/s",
                "traffic": traffic,
                "traffic_source": "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of this launch from "
                                  "the committed ncu capture profiles/r2_ncu_dominant.json",
                "how": "algorithmic FLOPs of the launch / mean CUDA-event duration around that launch on the update "
                       "stream, eager instrumented pass of the same update (a busy-wait kernel keeps the stream ahead "
                       "of the host); the graph replay itself cannot be split by events",
                "dominant_launch": dom,
                "tensor_core_aggregate": {"algorithmic_gflop_per_update": flops / 1e9, "us_per_update": tc_us,
                                          "tflops": flops / (tc_us * 1e-6) / 1e12 if tc_us else 0.0,
                                          "share_of_kernel_time": tc_us / max(all_us, 1e-9)},
                "families": prof}
        hbm = hbm_block(r, peaks)
        # ---- CPU baseline: the oracle port on this host's cores, bounded sample
        cores = os.cpu_count() or 1
        best, swept = None, {}
        for th in ([] if strong else thread_sweep(cores)):
            rate, done, dts, _ = time_oracle(w, hb, th, 100000, 1, budget_s=6.0)
            swept[th] = round(rate, 2)
            if best is None or rate > best[0]:
                best = (rate, th, done, dts)
        cpu_baseline = None if best is None else {
            "value": best[0], "unit": "updates/s", "cores": best[1], "host_cores": cores, "kind": "port",
            "sample": f"{best[2]} full updates (same config, batch 256) of oracle/update.py in {best[3]:.1f}s; torch threads "
                      f"swept over {thread_sweep(cores)} of {cores} host cores: {swept} updates/s"}
        # ---- second line of SURVEY 8d: the reproduction script's variant (reproductions/offline/cql.py sets
        # alpha_learning_rate=0.0, so update_alpha and its importance-sampling pass are not executed)
        if w["algo"] == "cql" and not strong:
            r0 = Runner(w, 1, 0, local, args.precision, False, data=r.data, alpha_learning_rate=0.0)
            k0 = max(20, K // 3)
            t0, s0, _, _ = r0.timed(k0, max(3, W // 2), flush_buf)
            variant = {"alpha_learning_rate": 0.0, "updates_per_s": k0 / (t0 * 1e-3), "ms_per_step": t0 / k0,
                       "note": "reproductions/offline/cql.py variant: no update_alpha step"}

    trace("assembling the line")
    if rank == 0:
        line = {
            "metric": METRIC if not strong else "CQL gradient updates/sec at batch 8192 (c5, sharded)",
            "value": head["value"], "unit": "updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "tf32x3", "data": "synthetic",
            "config": {"workload": w["desc"], "per_gpu_batch": r.B,
                       "precision": "bf16 operands / fp32 accumulate, fp32 master weights and optimizer (losses within 1e-2; "
                                    "see fp32_parity_mode for the mode that meets 1e-5)"
                       if args.precision == "bf16" else "fp32 (3xTF32 tensor-core GEMMs, fp32 accumulate)",
                       "global_batch": r.B * world,
                       "units": "batch-8192 updates per second (minibatch sharded over ranks)" if strong
                       else "updates of 256-transition minibatches per second, summed over ranks",
                       "parallelism": plane,
                       "l2": ("flushed (256 MiB write) between timed steps" + (", barrier after the flush" if world > 1 else ""))
                       if not args.no_flush else "not flushed",
                       "timing": "CUDA events per step on the launching stream, max over ranks"},
            "step_ms_p10_p50_p90": head["step_ms_p10_p50_p90"], "variant_alpha_lr0": variant,
            "value_l2_warm": warm,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "graph_nodes_per_update": graph_nodes,
            "fp32_parity_mode": fp32_mode, "extra": extra, "dp_check": check,
            "roofline": roof, "hbm": hbm, "cpu_baseline": cpu_baseline,
        }
        assert tuple(line["config"]) == CONFIG_KEYS
        print(json.dumps(line), flush=True)
        trace("line printed")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    trace("exit")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the extra workloads / roofline / baselines")
    ap.add_argument("--profile-dp", action="store_true", help="N > 1: add per-launch device times of the sharded update")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: tcgen05 tensor-core GEMMs (bf16 operands, fp32 accumulate; losses within 1e-2); "
                         "fp32: 3xTF32 tcgen05 GEMMs, fp32 everything else (parity 1e-5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
