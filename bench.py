#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the offline-RL update path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c5] [--impl ours|reference]

Metric (BASELINE.json): CQL gradient updates/sec at batch 256 (config c2: obs 17, act 6, B 256,
N 10, 2 critics, 3x256 MLPs, synthetic halfcheetah-shaped data).  One "step" = one
`algo.update(batch)` (temp -> alpha -> critic -> actor -> target sync) on one minibatch.

  value : steps/s with the replay buffer and the step's indices already in HBM — per step one gather
          kernel launch + one CUDA-graph launch, timed with CUDA events on the launching stream,
          L2 flushed between timed steps.
  e2e   : the same metric through the reference-facing call `algo.update(TransitionMiniBatch-like
          numpy batch)`: pinned H2D of the six arrays + graph + pinned D2H of the metrics, every step.
  roofline / cpu_baseline : see DESIGN.md §Measurement.

`--impl reference` times the CPU restatement of the reference update (oracle/update.py — plain
PyTorch fp32 + autograd + torch.optim.Adam, i.e. what the reference executes with use_gpu=False) on
the host cores with all threads, on the same config/metric.
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


def time_oracle(w, batches, threads, steps, warmup, budget_s=None):
    import torch

    from oracle import update as ou

    torch.set_num_threads(threads)
    orc = build_oracle(w)
    noise = ou.Noise(seed=0)
    for i in range(warmup):
        orc.update(ou.Batch(batches[i % len(batches)]), noise)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        orc.update(ou.Batch(batches[i % len(batches)]), noise)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def host_batches(w, n_batches, obs, act, rew, term):
    """Minibatches as the reference's sampler would build them (numpy arrays), via the oracle sampler."""
    from oracle import sampler as osampler

    small = 20_000  # the oracle's flat metadata build is a Python loop; a slice of the data is enough
    replay = osampler.FlatReplay(obs[:small], act[:small], rew[:small], term[:small])
    rs = np.random.RandomState(1)
    return [osampler.gather(replay, rs.randint(len(replay), size=w["batch"])) for _ in range(n_batches)]


def run_reference(args, w):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    obs, act, rew, term = make_dataset(w, steps_total=50_000)
    batches = host_batches(w, 4, obs, act, rew, term)
    cores = os.cpu_count() or 1
    # the reference arm gets its best thread count: short probe at {cores/2, cores}, full run with the faster one
    probes = {th: time_oracle(w, batches, th, 4, 1, budget_s=20.0)[0] for th in sorted({max(1, cores // 2), cores})}
    threads = max(probes, key=probes.get)
    rate, done, dt = time_oracle(w, batches, threads, args.steps, min(args.warmup, 3), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": done, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "batch": w["batch"]},
        "cpu_baseline": {"value": rate, "unit": "updates/s", "cores": threads, "kind": "port",
                         "sample": f"{done} full updates of the oracle port (plain PyTorch fp32 CPU) in {dt:.1f}s, "
                                   f"torch threads={threads} of {cores} host cores"},
        "e2e": {"value": rate, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "torch_threads": torch.get_num_threads(),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- our arm
def build_algo(w, world_size=1, rank=0, precision="bf16", **extra):
    from d3rlpy_b200.algos import CQL, TD3PlusBC

    kw = dict(world_size=world_size, rank=rank) if world_size > 1 else {}
    kw["precision"] = precision
    kw.update(extra)
    if w["algo"] == "cql":
        algo = CQL(actor_encoder_factory=w["hidden"], critic_encoder_factory=w["hidden"], batch_size=w["batch"],
                   n_action_samples=w["n"], n_critics=w["critics"], use_gpu=int(os.environ.get("LOCAL_RANK", "0")), **kw)
    else:
        algo = TD3PlusBC(actor_encoder_factory=w["hidden"], critic_encoder_factory=w["hidden"], batch_size=w["batch"],
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


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    dev = torch.device("cuda", local)
    algo = build_algo(w, world, rank, args.precision)
    impl = algo.impl
    # c2 (default): weak scaling — every rank trains on its own 256 rows of a global batch 256*W (gradients
    # all-reduced), units = 256-row minibatches/s summed over ranks.  c5: STRONG scaling — the batch-8192 update
    # is sharded over the ranks, units = batch-8192 updates/s.
    strong = args.workload == "c5"
    assert not strong or w["batch"] % world == 0
    B = w["batch"] // world if strong else w["batch"]
    unit_scale = 1 if strong else world
    obs, act, rew, term = make_dataset(w)
    ds = MDPDataset(obs, act, rew, term)
    replay = ds.device_replay(dev)
    K, W = args.steps, args.warmup
    if strong:  # every rank draws the same global index vector and takes its own row shard
        rs = np.random.RandomState(1)
        idx_all = rs.randint(len(replay), size=(K + W, B * world))[:, rank * B:(rank + 1) * B].astype(np.int64)
    else:
        rs = np.random.RandomState(1 + rank)
        idx_all = rs.randint(len(replay), size=(K + W, B)).astype(np.int64)
    idx_all = np.ascontiguousarray(idx_all)
    idx_dev = torch.from_numpy(idx_all).to(dev)
    db = impl.device_batch(B)
    L = impl._lib
    st = impl._stream
    holder = SimpleNamespace(_device_batch=db)

    def gather(i):
        L.gather_vector(replay.obs.data_ptr(), w["obs"], replay.actions.data_ptr(), w["act"], 0,
                        replay.rewards.data_ptr(), replay.meta.data_ptr(), idx_dev[i].data_ptr(), B, 1, 0.99,
                        db.ptr("obs"), db.ptr("act"), db.ptr("rew"), db.ptr("next_obs"), db.ptr("term"),
                        db.ptr("nsteps"), None, None, 0.0, st)

    def step_device(i):
        gather(i)
        impl.update_fused_async(holder) if w["algo"] == "cql" else impl.update_fused_async(holder, algo.grad_step % 2 == 0)
        algo._grad_step += 1

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(W):
        step_device(i)
    barrier()
    # ---- device-resident timing: CUDA events per step on the launching stream, L2 flushed in between
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(5.0)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    n0 = L.launch_count()
    with torch.cuda.stream(impl._stream_obj):
        for i in range(K):
            if not args.no_flush:
                flush_buf.fill_(float(i))
            evs[i][0].record(impl._stream_obj)
            step_device(W + i)
            evs[i][1].record(impl._stream_obj)
    barrier()
    clocks = sampler.stop()
    eager_launches = L.launch_count() - n0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    graph_nodes = max(impl._graph_nodes.values()) if impl._graph_nodes else 0
    launches = int(eager_launches + K * graph_nodes)

    # ---- end-to-end: host numpy batch -> pinned H2D -> graph -> pinned D2H, every step
    e2e = None
    cpu_baseline = None
    roof = None
    if True:
        hb = host_batches(dict(w, batch=B), 8, obs, act, rew, term)
        hbs = [SimpleNamespace(**b) for b in hb]
        n_e2e = max(10, min(K, 200))
        for i in range(3):
            algo.update(hbs[i % len(hbs)])
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            m = algo.update(hbs[i % len(hbs)])
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": unit_scale * n_e2e / dt, "unit": "updates/s", "h2d_bytes_per_step": impl._batch.h2d_bytes,
               "d2h_bytes_per_step": 4 * 64, "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e}
        assert all(np.isfinite(float(v)) for v in m.values()), m

    if rank == 0 and world == 1:
        # ---- roofline of the dominant kernel family (dense layers), measured live with CUDA events
        prof, dom = kernel_profile(algo, hbs[0])
        tc = [v for k, v in prof.items() if k.startswith("linear_") or k.startswith("umma_gemm") or k.startswith("mlp_")]
        tc_us = sum(v["us_per_update"] for v in tc)
        all_us = sum(v["us_per_update"] for v in prof.values())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        # the dominant kernel is timed alone (events around the single launch) -> burst peak
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        flops = req_gemm_flops(w)
        achieved = dom["gflop_per_launch"] / (dom["us_per_launch"] * 1e-6) / 1e3 if dom else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_dominant.json")))["dram_bytes_per_launch"]
        except Exception:  # noqa: BLE001
            pass
        roof = {"bound": "tensor",
                "kernel": (f"{dom['name']} (largest launch: fused critic trunk+head, tcgen05.mma + TMA + TMEM, "
                           "all members and all importance-sampling rows of the alpha and critic steps)")
                if args.precision == "bf16" and dom else "gemm_f32_kernel",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks
                else "fallback 1.59 PFLOP/s",
                "traffic": traffic,
                "how": "algorithmic FLOPs of the launch / mean CUDA-event duration around that launch on the update "
                       "stream, eager instrumented pass of the same update (a busy-wait kernel keeps the stream ahead "
                       "of the host); the graph replay itself cannot be split by events",
                "dominant_launch": dom,
                "tensor_core_aggregate": {"algorithmic_gflop_per_update": flops / 1e9, "us_per_update": tc_us,
                                          "tflops": flops / (tc_us * 1e-6) / 1e12 if tc_us else 0.0,
                                          "share_of_kernel_time": tc_us / max(all_us, 1e-9)},
                "families": prof}
        # ---- CPU baseline: the oracle port on this host's cores, bounded sample
        cores = os.cpu_count() or 1
        best = None
        for th in ([] if strong else sorted({max(1, cores // 2), cores})):
            rate, done, dts = time_oracle(w, hb, th, 100000, 1, budget_s=8.0)
            if best is None or rate > best[0]:
                best = (rate, th, done, dts)
        cpu_baseline = None if best is None else {"value": best[0], "unit": "updates/s", "cores": best[1], "kind": "port",
                        "sample": f"{best[2]} full updates (same config, batch 256) of oracle/update.py in {best[3]:.1f}s; "
                                  f"threads swept over {{{max(1, cores // 2)},{cores}}} of {cores} host cores"}

    # ---- second line of SURVEY 8d: the reproduction script's variant (reproductions/offline/cql.py sets
    # alpha_learning_rate=0.0, so update_alpha and its importance-sampling pass are not executed)
    variant = None
    if world == 1 and w["algo"] == "cql" and not strong:
        algo0 = build_algo(w, 1, 0, args.precision, alpha_learning_rate=0.0)
        impl0 = algo0.impl
        holder0 = SimpleNamespace(_device_batch=db)   # same gathered minibatch buffers
        for i in range(W):
            gather(i)
            impl0.update_fused_async(holder0)
        impl0.sync()
        ev0 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        with torch.cuda.stream(impl0._stream_obj):
            for i in range(K):
                if not args.no_flush:
                    flush_buf.fill_(float(i))
                impl0._stream_obj.wait_stream(impl._stream_obj)
                ev0[i][0].record(impl0._stream_obj)
                L.gather_vector(replay.obs.data_ptr(), w["obs"], replay.actions.data_ptr(), w["act"], 0,
                                replay.rewards.data_ptr(), replay.meta.data_ptr(), idx_dev[W + i].data_ptr(), B, 1, 0.99,
                                db.ptr("obs"), db.ptr("act"), db.ptr("rew"), db.ptr("next_obs"), db.ptr("term"),
                                db.ptr("nsteps"), None, None, 0.0, impl0._stream)
                impl0.update_fused_async(holder0)
                ev0[i][1].record(impl0._stream_obj)
        torch.cuda.synchronize(dev)
        ms0 = float(np.sum([a.elapsed_time(b) for a, b in ev0]))
        variant = {"alpha_learning_rate": 0.0, "updates_per_s": K / (ms0 * 1e-3), "ms_per_step": ms0 / K,
                   "note": "reproductions/offline/cql.py variant: no update_alpha step"}

    if rank == 0:
        line = {
            "metric": METRIC if not strong else "CQL gradient updates/sec at batch 8192 (c5, sharded)", "value": unit_scale * K / (total_ms * 1e-3), "unit": "updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "per_gpu_batch": B,
                       "precision": "bf16 operands / fp32 accumulate, fp32 master weights and optimizer"
                       if args.precision == "bf16" else "fp32", "global_batch": B * world,
                       "units": "batch-8192 updates per second (minibatch sharded over ranks)" if strong else "updates of 256-transition minibatches per second, summed over ranks",
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "l2": "flushed (256 MiB write) between timed steps" if not args.no_flush else "not flushed",
                       "timing": "CUDA events per step on the launching stream, max over ranks",
                       "step_ms_p10_p50_p90": [float(np.percentile(step_ms, p)) for p in (10, 50, 90)],
                       "variant_alpha_lr0": variant},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "graph_nodes_per_update": graph_nodes,
            "roofline": roof, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: tcgen05 tensor-core GEMMs (bf16 operands, fp32 accumulate; parity 1e-2); "
                         "fp32: SIMT GEMMs (parity 1e-5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
