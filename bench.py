#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec of the batched search loop on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: ONE plan call that builds one search tree per root state. The headline
workload is BASELINE config 2 on the DEFAULT SimpleGridWorld (tprob 0.7): vanilla UCT, 65,536 roots per GPU, n_iterations=1000,
depth=20, c=5, the deterministic rollout policy FunctionPolicy(s -> :up) (test/options.jl:29), rollout max_depth=50.

  value       whole-job simulations/s with roots and outputs resident in HBM (mctsb200_plan_device)
  e2e         the same through the public host-buffer call (mctsb200_plan via Context.plan): pinned host roots -> H2D -> search ->
              D2H of actions / best_Q / status inside the timed region
  roofline    SURVEY.md 8(d): algorithmic bytes (144 B per tree level + 112 B per inserted node, reference data widths) / the
              search kernel's CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs. `traffic` is the DRAM traffic of the same
              launch from the committed ncu capture and `dram_gbs` = traffic / duration: the tables are served by L1 / L2, so
              the kernel is bound by the latency of dependent row loads, not by HBM bandwidth -- `binding` says so.
  strong      (every N) the two BASELINE configurations that are DEFINED across GPUs, at a total size that does not change with N:
              config 4 (MountainCar DPW, 8,192 roots in all, sharded over the ranks, no collective) and config 5 (root-parallel
              UCT of one root: a fixed ensemble of 9,472 trees x 1,056 iterations = 10^7 simulations per decision, the ranks'
              integer partial sums merged by ONE ncclAllReduce inside mctsb200_plan_root_parallel, inside the timed region)
  workloads   (N = 1) every other workload with its own roofline record
  cpu_baseline   the CPU oracle (a C++ restatement of the reference; Julia is not installed) on a bounded sample of each
              workload, all host threads (thread count stated) and one thread (the reference solver is single-threaded)

N > 1 (torchrun, one rank per GPU): the headline is weak-scaled (every rank plans its own 65,536-root batch, global root indices
offset by rank, no data-path collective); torch.distributed carries the barrier, the max-over-ranks of the elapsed time and the
128-byte NCCL id of the library's own communicator.

`--impl reference` times the reference's CPU path (the oracle port, all host threads) on the same config / metric / unit, each
step a bounded sample of the workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC, UNIT = "MCTS simulations/sec", "simulations/s"


def workloads():
    from mcts_jl_b200 import workloads as W

    return W


def config_of(name, n_roots, world):
    """The workload's identity: the same dict in the product arm and in the reference arm."""
    W = workloads()
    w = W.WORKLOADS[name]
    kw = w["solver_kw"]
    return {
        "workload": name, "roots_per_gpu": int(n_roots), "n_iterations": int(kw["n_iterations"]), "depth": int(kw["depth"]),
        "exploration_constant": float(kw.get("exploration_constant", 1.0)), "rollout_max_depth": 50, "rollout_policy": w["policy"],
        "mdp": {"gw": "SimpleGridWorld", "mc": "MountainCar", "mc-plugin": "MountainCar (run-time plugin)", "noisy-plugin": "NoisyMountainCar (run-time plugin)", "box": "GaussianBox (continuous actions)"}[w["mdp"]],
        "mdp_params": {k: v for k, v in w["mdp_kw"].items()}, "solver": "MCTSSolver" if w["kind"] == "uct" else "DPWSolver",
        "sharding": f"roots x{world} (rank r plans global roots [r*{n_roots}, (r+1)*{n_roots}))",
        "l2": "node tables of all trees are larger than L2 and rebuilt from zero every step (no flush needed)",
    }


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
    """DRAM bytes per launch of the search kernel from the committed ncu capture of this workload (profiles/traffic.json), if one
    exists: (bytes, source)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        v = d.get(workload)
        if isinstance(v, dict):
            return int(v["bytes"]), v.get("source")
        return (int(v), d.get("_source")) if isinstance(v, (int, float)) else (None, None)
    except Exception:
        return None, None


def roofline_record(workload, kind, stats_list):
    """algorithmic bytes / kernel time against the measured HBM peak, with the measured DRAM traffic beside it."""
    st = stats_list[-1]
    kernel_ms = float(np.mean([s.kernel_ms for s in stats_list]))
    alg = int(st.algorithmic_bytes)
    peak, peak_src = measured_peak()
    achieved = alg / (kernel_ms / 1e3) / 1e9
    traffic, tsrc = ncu_traffic(workload)
    kname = ("uct_fast_kernel" if kind == "uct" else "dpw_fast_kernel") if int(st.kernel_variant) == 1 else "search_kernel"
    rec = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
           "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": kernel_ms, "kernel": kname + " (one launch per step)",
           "achieved_is": "ALGORITHMIC bytes (reference data widths, SURVEY.md 8d) per second -- a normalised work rate, not DRAM bandwidth"}
    if traffic:
        rec["traffic_source"] = tsrc
        rec["dram_gbs"] = traffic / (kernel_ms / 1e3) / 1e9
        rec["dram_frac"] = rec["dram_gbs"] / peak
        rec["binding"] = ("latency: the node tables are served from L1/L2 (DRAM traffic is %.1fx below the algorithmic bytes); what binds is the "
                          "dependent row load + selection chain of each tree, see profiles/" % (alg / max(traffic, 1)))
    return rec


def cpu_sample(workload, seconds_target=15.0, threads=None, single_seconds=5.0):
    """Times the oracle on a bounded sample of the workload (roots subsampled, iterations unchanged)."""
    import mcts_jl_b200 as m
    from oracle.oracle_binding import Oracle

    W = workloads()
    orc = Oracle()
    threads = threads or max(1, os.cpu_count() or 1)
    same_root = W.CONFIG5_ROOT if workload.startswith("config5") else None
    mdp, solver, roots = W.make_workload(workload, same_root=same_root)
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
    out = {"value": st.n_simulations / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{n} of the workload's roots x {cfg.n_iterations} iterations (oracle, {threads} threads, {dt:.1f} s)",
           "rollout_steps_per_s": st.n_rollout_steps / dt, "seconds": dt, "n_simulations": int(st.n_simulations),
           "n_iterations": int(cfg.n_iterations), "depth": int(cfg.depth)}
    if single_seconds > 0:
        # the reference solver itself is single-threaded (one planner, one tree): the same port on ONE thread
        n1 = int(max(1, min(len(roots), single_seconds / max(per_root * threads, 1e-9))))
        t1 = time.perf_counter()
        _, _, _, st1 = orc.plan(cfg, oid, roots[:n1], n_threads=1, keep_trees=False)
        dt1 = time.perf_counter() - t1
        out["single_thread"] = {"value": st1.n_simulations / dt1, "unit": UNIT, "cores": 1,
                                "sample": f"{n1} of the workload's roots x {cfg.n_iterations} iterations (oracle, 1 thread, {dt1:.1f} s)"}
    del keep
    return out


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    W = workloads()
    threads = max(1, os.cpu_count() or 1)
    vals, last = [], None
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        last = cpu_sample(args.workload, seconds_target=per_step, threads=threads, single_seconds=0)
        if i >= args.warmup:
            vals.append(last)
    sims = sum(v["n_simulations"] for v in vals)
    secs = sum(v["seconds"] for v in vals)
    value = sims / secs
    n_roots = args.roots or W.WORKLOADS[args.workload]["n_roots"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(vals)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload, n_roots, args.gpus),
        "note": "reference CPU path = C++ oracle port of MCTS.jl (Julia is not installed); each step is a bounded sample of the workload's roots "
                "(the metric is a rate and every root costs the same search), see cpu_baseline.sample",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    # rank 0 prints exactly ONE line on stdout, the JSON line: whatever a library writes to fd 1 meanwhile (NCCL's version banner
    # when the library's own communicator comes up, for one) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    W = workloads()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=W.HEADLINE, choices=sorted(W.WORKLOADS))
    ap.add_argument("--roots", type=int, default=0, help="override roots per GPU (default: the workload's)")
    ap.add_argument("--iterations", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per tree (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload: no strong-scaling records, no other workloads")
    ap.add_argument("--same-root", default="", help="experiment: x,y of a GridWorld root used for every tree")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist

    import mcts_jl_b200 as m
    from mcts_jl_b200 import distributed as D

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

    stream = torch.cuda.current_stream()
    ctx = m.Context(local_rank, stream=stream.cuda_stream)

    def timed(fn, steps, warmup, sample_clocks=False):
        """W untimed steps, then K steps between CUDA events on the launching stream, barrier + synchronize on both sides,
        max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank)
        if sample_clocks and not os.environ.get("BENCH_NO_SAMPLER"):
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        out = [fn() for _ in range(steps)]
        ev1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, clocks

    class Batch:
        """One workload's roots on this rank: pinned host copy, device copy, output buffers."""

        def __init__(self, name, n_roots=None, base=None, lanes=0, n_iterations=None, same_root=None, lo=None, hi=None):
            self.name = name
            self.mdp, self.solver, roots_np = W.make_workload(name, n_roots, lanes, n_iterations, same_root)
            if lo is not None:  # a shard of a fixed-size batch
                roots_np = np.ascontiguousarray(roots_np[lo:hi])
            self.n = roots_np.shape[0]
            self.base = base if base is not None else 0
            self.cfg, self._keep = m.build_cfg(self.solver, self.mdp)
            ctx.configure(self.mdp.mdp_id, self.mdp.params())
            self.state_bytes = roots_np.dtype.itemsize * roots_np.shape[1]
            n = max(self.n, 1)
            self.h_roots = torch.from_numpy(roots_np).pin_memory() if self.n else None
            self.d_roots = self.h_roots.to("cuda") if self.n else None
            self.d_act = torch.empty(n, dtype=torch.int32, device="cuda")
            self.d_q = torch.empty(n, dtype=torch.float64, device="cuda")
            self.d_st = torch.empty(n, dtype=torch.int32, device="cuda")
            self.h_act = torch.empty(n, dtype=torch.int32).pin_memory()
            self.h_q = torch.empty(n, dtype=torch.float64).pin_memory()
            self.h_st = torch.empty(n, dtype=torch.int32).pin_memory()

        def step_device(self):
            if self.n == 0:
                return None
            ctx.configure(self.mdp.mdp_id, self.mdp.params())
            return ctx.plan_device(self.cfg, self.mdp.mdp_id, self.d_roots.data_ptr(), self.n, self.state_bytes, self.d_act.data_ptr(),
                                   self.d_q.data_ptr(), self.d_st.data_ptr(), self.base)

        def step_e2e(self):
            if self.n == 0:
                return None
            ctx.configure(self.mdp.mdp_id, self.mdp.params())
            return ctx.plan(self.cfg, self.mdp.mdp_id, self.h_roots.numpy(), self.base, want_status=True,
                            out=(self.h_act.numpy()[: self.n], self.h_q.numpy()[: self.n], self.h_st.numpy()[: self.n]))[3]

    def measure(batch, steps, warmup, sample_clocks=False, e2e=True):
        """value / e2e / roofline of one workload at `world` GPUs (weak: every rank runs its own batch)."""
        ms_dev, st_dev, clocks = timed(batch.step_device, steps, warmup, sample_clocks)
        sims_step = int(st_dev[-1].n_simulations)
        rec = {
            "value": world * sims_step * steps / (ms_dev / 1e3), "unit": UNIT, "ms_per_step": ms_dev / steps,
            "rollout_steps_per_s": world * int(st_dev[-1].n_rollout_steps) * steps / (ms_dev / 1e3),
            "simulations_per_step": sims_step, "tree_levels_per_step": int(st_dev[-1].n_tree_levels),
            "rollout_steps_per_step": int(st_dev[-1].n_rollout_steps), "lanes_per_tree": int(st_dev[-1].lanes_per_tree),
            "gpu_launches": int(sum(s.n_kernel_launches for s in st_dev)),
            "roofline": roofline_record(batch.name, W.WORKLOADS[batch.name]["kind"], st_dev),
        }
        if e2e:
            ms_e2e, st_e2e, _ = timed(batch.step_e2e, steps, max(1, warmup // 2))
            rec["e2e"] = {"value": world * int(st_e2e[-1].n_simulations) * steps / (ms_e2e / 1e3), "unit": UNIT,
                          "h2d_bytes_per_step": int(st_e2e[-1].h2d_bytes), "d2h_bytes_per_step": int(st_e2e[-1].d2h_bytes),
                          "ms_per_step": ms_e2e / steps}
        return rec, clocks

    # ---- headline: weak-scaled batch of the chosen workload ---------------------------------------------------------------
    n_roots = args.roots or W.WORKLOADS[args.workload]["n_roots"]
    head = Batch(args.workload, n_roots, base=rank * n_roots, lanes=args.lanes, n_iterations=args.iterations or None,
                 same_root=args.same_root or None)
    rec, clocks = measure(head, args.steps, args.warmup, sample_clocks=True)
    cfg_d = config_of(args.workload, n_roots, world)
    if args.iterations:
        cfg_d["n_iterations"] = int(args.iterations)
    line = {
        "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg_d, "lanes_per_tree": rec["lanes_per_tree"],
        "rollout_steps_per_s": rec["rollout_steps_per_s"], "simulations_per_step": rec["simulations_per_step"],
        "tree_levels_per_step": rec["tree_levels_per_step"], "rollout_steps_per_step": rec["rollout_steps_per_step"],
        "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"], "clocks": clocks,
    }
    del head

    extras = not args.no_extras and args.workload == W.HEADLINE and not args.roots and not args.iterations
    if extras:
        # ---- strong scaling: the configurations that are defined across GPUs, total work fixed ----------------------------
        strong = {}
        k_steps, k_warm = max(1, min(args.steps, 2)), 1
        # config 4: 8 192 MountainCar roots in all, sharded in contiguous blocks; no collective
        lo, hi = D.shard_bounds(W.CONFIG4_TOTAL_ROOTS, rank, world)
        b4 = Batch("config4-dpw-mountaincar", W.CONFIG4_TOTAL_ROOTS, base=lo, lo=lo, hi=hi)
        ms4, st4, _ = timed(b4.step_device, k_steps, k_warm)
        sims4 = torch.tensor([float(st4[-1].n_simulations) if st4[-1] is not None else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sims4)
        ms4e, st4e, _ = timed(b4.step_e2e, k_steps, 0)
        strong["config4-dpw-mountaincar-8192-roots"] = {
            "scaling": "strong", "total_roots": W.CONFIG4_TOTAL_ROOTS, "roots_per_gpu": hi - lo, "n_iterations": 10000, "depth": 50,
            "ms_per_step": ms4 / k_steps, "value": float(sims4.item()) * k_steps / (ms4 / 1e3), "unit": UNIT, "steps": k_steps, "warmup": k_warm,
            "e2e": {"value": float(sims4.item()) * k_steps / (ms4e / 1e3), "unit": UNIT, "ms_per_step": ms4e / k_steps},
            "collective": None, "lanes_per_tree": int(st4[-1].lanes_per_tree) if st4[-1] is not None else None,
            "roofline": roofline_record("config4-dpw-mountaincar", "dpw", [s for s in st4 if s is not None]) if st4[-1] is not None else None,
        }
        del b4
        # config 5: ONE root, a fixed ensemble of 9 472 trees x 1 056 iterations (10^7 simulations per decision) split over the
        # ranks; the library merges the integer partial sums with ONE ncclAllReduce inside the call
        if world > 1:
            D.attach_comm(ctx, rank, world)
        mdp5, solver5, _ = W.make_workload("config5-root-parallel-uct-gridworld", 1)
        cfg5, keep5 = m.build_cfg(solver5, mdp5)
        ctx.configure(mdp5.mdp_id, mdp5.params())
        root5 = mdp5.encode_states([W.CONFIG5_ROOT])
        lo5, hi5 = D.shard_bounds(W.CONFIG5_TREES, rank, world)
        last = {}

        def step5():
            sn, snq, st = ctx.plan_root_parallel(cfg5, mdp5.mdp_id, root5, hi5 - lo5, tree_index_base=lo5)
            last["decision"] = ctx.root_parallel_decide(sn, snq, float(cfg5.init_Q))
            last["sum_n"] = sn
            return st

        ms5, st5, _ = timed(step5, max(2, k_steps), k_warm)
        n5 = max(2, k_steps)
        a5, bq5, q5 = last["decision"]
        strong["config5-root-parallel-uct-gridworld"] = {
            "scaling": "strong", "ensemble_trees": W.CONFIG5_TREES, "iterations_per_tree": W.CONFIG5_ITERATIONS,
            "simulations_per_decision": W.CONFIG5_TREES * W.CONFIG5_ITERATIONS, "trees_per_gpu": hi5 - lo5, "root": list(W.CONFIG5_ROOT),
            "ms_per_decision": ms5 / n5, "value": W.CONFIG5_TREES * W.CONFIG5_ITERATIONS * n5 / (ms5 / 1e3), "unit": UNIT, "steps": n5, "warmup": k_warm,
            "collective": "one ncclAllReduce(sum, int64, 21 words) per decision inside mctsb200_plan_root_parallel, inside the timed region" if world > 1 else None,
            "timed_region": "host root state -> H2D, search, partial sums, allreduce, D2H of the merged sums, decision",
            "action": int(a5), "best_q": float(bq5), "q": [float(x) for x in q5], "sum_n": [float(x) for x in last["sum_n"]],
            "note": "integer partial sums: action, Q and sum_n are identical at every N (bit for bit)",
        }
        del keep5
        line["strong"] = strong

        # ---- the other workloads, N = 1 -----------------------------------------------------------------------------------
        if world == 1:
            others = {}
            for name in ("config2-uct-gridworld-deterministic", "config2-uct-gridworld-stochastic", "config3-dpw-gridworld", "config4-dpw-mountaincar",
                         "box-dpw-continuous-actions"):
                b = Batch(name)
                r, _ = measure(b, max(1, min(args.steps, 3)), 1)
                r["config"] = config_of(name, b.n, 1)
                others[name] = r
                del b
            line["workloads"] = others

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_sample(args.workload)
            if extras:
                for name, dst in (("config4-dpw-mountaincar", line["strong"]["config4-dpw-mountaincar-8192-roots"]),
                                  ("config5-root-parallel-uct-gridworld", line["strong"]["config5-root-parallel-uct-gridworld"]),
                                  ("config2-uct-gridworld-deterministic", line["workloads"]["config2-uct-gridworld-deterministic"]),
                                  ("config2-uct-gridworld-stochastic", line["workloads"]["config2-uct-gridworld-stochastic"]),
                                  ("config3-dpw-gridworld", line["workloads"]["config3-dpw-gridworld"]),
                                  ("config4-dpw-mountaincar", line["workloads"]["config4-dpw-mountaincar"]),
                                  ("box-dpw-continuous-actions", line["workloads"]["box-dpw-continuous-actions"])):
                    dst["cpu_baseline"] = cpu_sample(name, seconds_target=6.0, single_seconds=2.0)
        except Exception as ex:  # the baseline must never break the GPU line
            line.setdefault("cpu_baseline", {"error": str(ex)})
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
