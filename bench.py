#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec of the batched search loop on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: ONE plan call that builds one search tree per root
state (BASELINE config 2: SimpleGridWorld vanilla UCT, 65,536 roots, n_iterations=1000, depth=20, c=5,
deterministic variant tprob=1.0 + FunctionPolicy(s->:up) rollouts, max_depth=50).

  value     whole-job simulations/s with roots and outputs resident in HBM (mctsb200_plan_device)
  e2e       the same through the public host-buffer call (mctsb200_plan via Context.plan): pinned host
            roots -> H2D -> search -> D2H of actions/best_Q/status inside the timed region
  roofline  algorithmic bytes (144 B per tree level + 112 B per inserted node) / the search
            kernel's CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the CPU oracle (a C++ restatement of the reference; Julia is not installed) on a bounded
            sample of the same workload, all host threads

N > 1 (torchrun, one rank per GPU): roots are sharded by rank with no data-path collective ("weak":
every rank plans its own 65,536-root batch, global root indices offset by rank); torch.distributed is only
used for the barrier and the max-over-ranks of the elapsed time.

`--impl reference` times the reference's CPU path (the oracle port, all host threads) on the same
config / metric / unit, each step a bounded sample of the workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (mdp kwargs, solver kind, solver kwargs, n_roots)
    "config2-uct-gridworld-deterministic": dict(mdp="gw", mdp_kw=dict(tprob=1.0), kind="uct", policy="up",
                                                solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=5.0), n_roots=65536),
    # the default SimpleGridWorld (tprob = 0.7) with the deterministic rollout policy: the other reading of BASELINE config 2
    "config2-uct-gridworld-default": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="uct", policy="up",
                                          solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=5.0), n_roots=65536),
    "config2-uct-gridworld-stochastic": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="uct", policy="random",
                                             solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=5.0), n_roots=65536),
    "config3-dpw-gridworld": dict(mdp="gw", mdp_kw=dict(tprob=0.7), kind="dpw", policy="random",
                                  solver_kw=dict(n_iterations=1000, depth=20, exploration_constant=10.0, k_state=4.0, alpha_state=1 / 8,
                                                 k_action=4.0, alpha_action=1 / 8), n_roots=16384),
    "config4-dpw-mountaincar": dict(mdp="mc", mdp_kw=dict(), kind="dpw", policy="random", solver_kw=dict(n_iterations=10000, depth=50),
                                    n_roots=1024),
    # (not a BASELINE config) vanilla UCT on MountainCar: the tile kernel's successor cache for single-successor hashed models
    "uct-mountaincar": dict(mdp="mc", mdp_kw=dict(), kind="uct", policy="random",
                            solver_kw=dict(n_iterations=10000, depth=50, exploration_constant=10.0), n_roots=1024),
    # config 4 with MountainCar registered at run time as a plugin (mctsb200_mdp_register: the NVRTC-compiled tile kernel), and
    # its noisy variant (continuous-state stochastic transitions: state widening really widens)
    "config4-dpw-mountaincar-plugin": dict(mdp="mc-plugin", mdp_kw=dict(), kind="dpw", policy="random",
                                           solver_kw=dict(n_iterations=10000, depth=50), n_roots=1024),
    "config4-dpw-noisy-mountaincar-plugin": dict(mdp="noisy-plugin", mdp_kw=dict(), kind="dpw", policy="random",
                                                 solver_kw=dict(n_iterations=10000, depth=50), n_roots=1024),
}


def make_workload(name, n_roots=None, lanes=0, n_iterations=None, same_root=None):
    import mcts_jl_b200 as m
    from helpers import gridworld_roots, mountaincar_roots

    w = WORKLOADS[name]
    n = n_roots or w["n_roots"]
    mdp = {"gw": m.SimpleGridWorld, "mc": m.MountainCar, "mc-plugin": m.MountainCarPlugin, "noisy-plugin": m.NoisyMountainCar}[w["mdp"]](**w["mdp_kw"])
    pol = {"up": m.ConstantPolicy("up"), "random": m.RandomSolver()}[w["policy"]]
    kw = dict(w["solver_kw"])
    if n_iterations:
        kw["n_iterations"] = n_iterations
    kw["estimate_value"] = m.RolloutEstimator(pol, max_depth=50)
    kw["lanes_per_tree"] = lanes
    solver = (m.MCTSSolver if w["kind"] == "uct" else m.DPWSolver)(rng=0, **kw)
    roots = gridworld_roots(n) if w["mdp"] == "gw" else mountaincar_roots(n)
    if same_root:  # experiment: every tree plans from the same root state
        roots = [tuple(int(v) for v in same_root.split(","))] * n
    return mdp, solver, mdp.encode_states(roots)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 20 ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_ev = index, [], set(), None, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_ev.wait(0.02)

    def stop(self):
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram bytes per launch of the search kernel from the committed ncu capture, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            v = json.load(f).get(workload)
            return int(v) if isinstance(v, (int, float)) else None
    except Exception:
        return None


def cpu_sample(workload, seconds_target=15.0, threads=None):
    """Times the oracle on a bounded sample of the workload (roots subsampled, iterations unchanged)."""
    import mcts_jl_b200 as m
    from oracle.oracle_binding import Oracle

    orc = Oracle()
    threads = threads or max(1, os.cpu_count() or 1)
    mdp, solver, roots = make_workload(workload)
    cfg, keep = m.build_cfg(solver, mdp)
    oid = mdp.mdp_id if getattr(mdp, "oracle_mdp_id", None) is None else mdp.oracle_mdp_id  # (plugins: the oracle's own id)
    orc.configure(oid, mdp.params())
    # calibrate on a small sample, then size the real one for ~seconds_target
    t0 = time.perf_counter()
    _, _, _, st = orc.plan(cfg, oid, roots[: 2 * threads], n_threads=threads, keep_trees=False)
    dt = max(time.perf_counter() - t0, 1e-6)
    per_root = dt / (2 * threads)
    n = int(min(len(roots), max(threads, seconds_target / per_root)))
    n = max(threads, n // threads * threads)
    t0 = time.perf_counter()
    _, _, _, st = orc.plan(cfg, oid, roots[:n], n_threads=threads, keep_trees=False)
    dt = time.perf_counter() - t0
    # the reference solver itself is single-threaded (one planner, one tree): the same port on ONE thread, a ~5 s sample
    n1 = int(max(1, min(len(roots), 5.0 / max(per_root * threads, 1e-9))))
    t1 = time.perf_counter()
    _, _, _, st1 = orc.plan(cfg, oid, roots[:n1], n_threads=1, keep_trees=False)
    dt1 = time.perf_counter() - t1
    single = {"value": st1.n_simulations / dt1, "unit": "simulations/s", "cores": 1,
              "sample": f"{n1} of the workload's roots x {cfg.n_iterations} iterations (oracle, 1 thread, {dt1:.1f} s)"}
    return {"value": st.n_simulations / dt, "unit": "simulations/s", "cores": threads, "kind": "port", "single_thread": single,
            "sample": f"{n} of the workload's roots x {cfg.n_iterations} iterations (oracle, {threads} threads, {dt:.1f} s)",
            "rollout_steps_per_s": st.n_rollout_steps / dt, "seconds": dt, "n_simulations": int(st.n_simulations),
            "n_iterations": int(cfg.n_iterations), "depth": int(cfg.depth)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = max(1, os.cpu_count() or 1)
    vals, last = [], None
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        last = cpu_sample(args.workload, seconds_target=per_step, threads=threads)
        if i >= args.warmup:
            vals.append(last)
    sims = sum(v["n_simulations"] for v in vals)
    secs = sum(v["seconds"] for v in vals)
    value = sims / secs
    line = {
        "impl": "reference", "metric": "MCTS simulations/sec", "value": value, "unit": "simulations/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(vals)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "n_iterations": int(last["n_iterations"]), "depth": int(last["depth"]),
                   "note": "reference CPU path = C++ oracle port of MCTS.jl (Julia not installed); each step is a bounded root sample"},
        "cpu_baseline": {"value": value, "unit": "simulations/s", "cores": threads, "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2-uct-gridworld-deterministic", choices=sorted(WORKLOADS))
    ap.add_argument("--roots", type=int, default=0, help="override roots per GPU (default: the workload's)")
    ap.add_argument("--iterations", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per tree (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--same-root", default="", help="experiment: x,y of a GridWorld root used for every tree")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import mcts_jl_b200 as m

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmctsb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mdp, solver, roots_np = make_workload(args.workload, args.roots or None, args.lanes, args.iterations or None, args.same_root or None)
    n_roots = roots_np.shape[0]
    base = rank * n_roots  # global root index of this rank's shard
    cfg, keep = m.build_cfg(solver, mdp)
    stream = torch.cuda.current_stream()
    ctx = m.Context(local_rank, stream=stream.cuda_stream)
    ctx.configure(mdp.mdp_id, mdp.params())

    tdt = torch.int64 if roots_np.dtype == np.int64 else torch.float64
    h_roots = torch.from_numpy(roots_np).pin_memory()
    d_roots = h_roots.to("cuda", non_blocking=False)
    d_act = torch.empty(n_roots, dtype=torch.int32, device="cuda")
    d_q = torch.empty(n_roots, dtype=torch.float64, device="cuda")
    d_st = torch.empty(n_roots, dtype=torch.int32, device="cuda")
    h_act = torch.empty(n_roots, dtype=torch.int32).pin_memory()
    h_q = torch.empty(n_roots, dtype=torch.float64).pin_memory()
    h_st = torch.empty(n_roots, dtype=torch.int32).pin_memory()
    state_bytes = roots_np.dtype.itemsize * roots_np.shape[1]

    def step_device():
        return ctx.plan_device(cfg, mdp.mdp_id, d_roots.data_ptr(), n_roots, state_bytes, d_act.data_ptr(), d_q.data_ptr(), d_st.data_ptr(), base)

    def step_e2e():
        return ctx.plan(cfg, mdp.mdp_id, h_roots.numpy(), base, want_status=True, out=(h_act.numpy(), h_q.numpy(), h_st.numpy()))[3]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank)
        if not os.environ.get("BENCH_NO_SAMPLER"):
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        stats = [fn() for _ in range(steps)]
        ev1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, stats, clocks

    ms_dev, st_dev, clocks = timed(step_device, args.steps, args.warmup)
    ms_e2e, st_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    sims_step = int(st_dev[-1].n_simulations)
    steps_roll = int(st_dev[-1].n_rollout_steps)
    value = world * sims_step * args.steps / (ms_dev / 1e3)
    e2e_value = world * int(st_e2e[-1].n_simulations) * args.steps / (ms_e2e / 1e3)
    kernel_ms = float(np.mean([s.kernel_ms for s in st_dev]))
    alg_bytes = int(st_dev[-1].algorithmic_bytes)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    launches = int(sum(s.n_kernel_launches for s in st_dev))

    line = {
        "metric": "MCTS simulations/sec", "value": value, "unit": "simulations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": args.workload, "roots_per_gpu": n_roots, "n_iterations": int(cfg.n_iterations), "depth": int(cfg.depth),
            "exploration_constant": float(cfg.exploration_constant), "rollout_max_depth": int(cfg.rollout_max_depth),
            "lanes_per_tree": int(st_dev[-1].lanes_per_tree), "sharding": f"roots x{world} (rank r plans global roots [r*{n_roots}, (r+1)*{n_roots}))",
            "l2": "node tables of all trees are larger than L2 and rebuilt from zero every step (no flush needed)",
        },
        "rollout_steps_per_s": world * steps_roll * args.steps / (ms_dev / 1e3),
        "simulations_per_step": sims_step, "tree_levels_per_step": int(st_dev[-1].n_tree_levels), "rollout_steps_per_step": steps_roll,
        "e2e": {"value": e2e_value, "unit": "simulations/s", "h2d_bytes_per_step": int(st_e2e[-1].h2d_bytes), "d2h_bytes_per_step": int(st_e2e[-1].d2h_bytes),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                     "kernel": ((("uct_fast_kernel" if WORKLOADS[args.workload]["kind"] == "uct" else "dpw_fast_kernel")
                                 if int(st_dev[-1].kernel_variant) == 1 else "search_kernel") + " (one launch per step)")},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_sample(args.workload)
        except Exception as ex:  # the baseline must never break the GPU line
            line["cpu_baseline"] = {"error": str(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
