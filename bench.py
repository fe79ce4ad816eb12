#!/usr/bin/env python
"""bench.py - LM iterations/sec of the g2o hot path on B200 (see BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload venice|sphere2500] [--impl reference]

A "step" is one Levenberg-Marquardt iteration (OptimizationAlgorithmLevenberg::solve: errors + chi2, linearize +
accumulate, [Schur], sparse Cholesky, back-substitution, oplus update, re-evaluation, lambda control) over one
synthetic input graph.  The optimisation is restarted from the initial estimates every RESTART steps (like running
`g2o -i 10` repeatedly) so that every step does comparable work; the restart copy is outside the per-step timers.

  value   steps/s with all inputs resident in HBM (CUDA events on the solver's stream, max over ranks)
  e2e     steps/s through the C-ABI with HOST buffers: per step, H2D of the vertex estimates from pinned memory, one
          LM iteration, D2H of the updated estimates + chi2 (what a Level-3 g2o adapter does around solve())
  roofline, cpu_baseline: see DESIGN.md section "Measurement"

N > 1 (torchrun): venice shards its landmarks over the ranks (cameras replicated, NCCL all-reduce of the reduced
camera system per LM trial) - strong scaling of the same graph.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
RESTART = 10  # LM iterations per optimisation run (the reference protocol: g2o -i 10)
ND_LEVELS = {"venice": 5, "ba10k": 7, "sphere2500": 7, "sphere40k": 7}  # dissection depth of the extra measurement


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="venice", choices=["venice", "sphere2500", "venice_small", "ba10k", "sphere40k"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parallel-ordering", action="store_true", help="skip the extra nested-dissection measurement")
    ap.add_argument("--nd-levels", type=int, default=0,
                    help="ordering of the reduced system: 0 = block AMD (reference, default); k = nested dissection, 2^k parts")
    return ap.parse_args()


def make_problem(workload):
    from openslam_g2o_b200 import synth
    if workload == "venice":
        return synth.venice_like(), "Venice-shaped BA (types_sba): 871 cameras / 530304 points / ~2.0M P2MC edges, seed 871"
    if workload == "ba10k":  # BASELINE.json configs[3] shape (one GPU here; shards by landmark under torchrun)
        return synth.venice_like(10000, 2000000, seed=10000, fixed_obs=10), "synthetic BA: 10000 cameras / 2000000 points / 20.0M P2MC edges (k = 10), seed 10000"
    if workload == "sphere40k":  # reduced BASELINE.json configs[4]: same generator, 200 x 200 instead of 1000 x 1000
        return synth.sphere(200, 200, seed=40000), "SE3 pose graph: sphere generator 200 x 200 = 40000 poses / 159399 edges, seed 40000"
    if workload == "venice_small":
        return synth.venice_like(100, 20000, seed=7), "small BA: 100 cameras / 20000 points"
    return synth.sphere(), "sphere2500 SE3 pose graph: 2500 poses / 9799 edges (create_sphere.cpp defaults), seed 2500"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    """The reference's own CPU implementation of the path: the oracle port (Eigen-free restatement of
    SparseOptimizer/BlockSolver/LM) linked against the reference's vendored CSparse compiled in oracle/_ref.
    Single thread: the reference builds with OpenMP OFF (CMakeLists.txt:137) and CSparse is sequential."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import LM, Oracle
    from openslam_g2o_b200 import synth
    prob, desc = make_problem(args.workload)
    o = Oracle()
    synth.feed(prob, o)
    o.setup_cli(True)
    o.initialize_optimization()
    total = args.warmup + args.steps
    budget_s = 150.0
    t_all = time.time()
    n, st = o.optimize(LM, 1)  # iteration 0: structure + symbolic (not timed)
    times, done = [], 1
    L = o.L
    import ctypes as C
    from oracle_binding import OracleStats
    # continue the same optimisation one LM iteration at a time (iteration index > 0: no re-analysis)
    L.oracle_lm_iteration.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    while done < total + 1 and (time.time() - t_all) < budget_s:
        s = OracleStats()
        t0 = time.perf_counter()
        L.oracle_lm_iteration(o.g, done, C.byref(s))
        times.append(time.perf_counter() - t0)
        done += 1
    timed = times[min(args.warmup, max(len(times) - 1, 0)):] or times
    ms = 1e3 * float(np.mean(timed))
    value = 1e3 / ms
    sample = "%d LM iterations of the same input after %d warm-up (iteration 0 = structure+symbolic excluded)" % (len(timed), len(times) - len(timed))
    print(json.dumps({
        "impl": "reference", "metric": "LM iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": len(timed), "warmup": len(times) - len(timed), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "solver": "lm_fix6_3 (CSparse flavour, block AMD)"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "oracle restatement of g2o's LM/BlockSolver + the reference's vendored CSparse compiled from /root/reference (oracle/_ref)"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
def algorithmic_bytes(workload, dims, info, n_hs, n_edges):
    """per-launch algorithmic bytes of the candidate dominant kernels, read-once/write-once"""
    if workload.startswith("venice") or workload == "ba10k":
        nl = dims["numLandmarks"]
        n_hpl, n_seg, n_contrib = info["hpl_slots"], info["schur_segments"], info["schur_contributions"]
        return {
            # ba_linearize_points: per edge cam idx 4 + meas 16 + info 24 + slot 4 + flag 1, write Hpl 144;
            # per landmark est 32 + eptr 4 + order 4 + vertex 4, write Hll 72 + b 24
            "linearize": n_edges * (49 + 144) + nl * 140,
            # schur_range: read every Hpl block once 144 + Wu 80 per landmark + 6 B per contribution index + 12 B per
            # segment descriptor; write one 36(+6)-double partial sum per segment
            "schur": n_hpl * 144 + nl * 84 + n_contrib * 6 + n_seg * (12 + 42 * 8),
            # ba_backsub: read Hpl 144 + slot/pose idx 8 per edge, Dinv 80 + b 24 + eptr/order 8 per landmark, write x 24
            "backsub": n_hpl * 144 + n_edges * 8 + nl * 136,
        }
    # pose graph: pg_linearize reads ids 8 + Zinv 96 + info 168 + 2 poses 192, writes the 120-double staging record
    return {"linearize": n_edges * (8 + 96 + 168 + 192 + 960)}


KERNEL_OF = {"linearize": {"ba": "ba_linearize_points_kernel", "pg": "pg_linearize_kernel<SE3>"},
             "schur": {"ba": "schur_range_kernel"}, "backsub": {"ba": "ba_backsub_kernel"}}


def ncu_traffic(kernel):
    """dram bytes (read + write) per launch of `kernel` from the committed ncu --set full summary, or None"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(kernel.split("<")[0])
    except Exception:
        return None


def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prob, desc = make_problem(args.workload)
    sharded = world > 1 and prob["kind"] == "ba"
    opt = g.SparseOptimizer(device=local_rank, shard=rank if sharded else 0, num_shards=world if sharded else 1)
    opt.set_algorithm("lm_fix6_3")
    synth.feed(prob, opt)
    opt.setup_cli()
    opt.initialize_optimization()
    opt._ensure_uploaded()
    ctx = opt.context
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    if sharded:
        from openslam_g2o_b200.distributed import make_allreduce
        ctx.set_allreduce(make_allreduce(ctx, local_rank), rank, world)
    if args.nd_levels:
        ctx.set_ordering(args.nd_levels)
    t_struct = time.perf_counter()
    assert ctx.build_structure()
    t_struct = time.perf_counter() - t_struct  # iteration-0 cost, reported beside the steady-state metric
    dims = ctx.dims()
    kinds = [g.VERTEX_CAM, g.VERTEX_XYZ] if prob["kind"] == "ba" else [g.VERTEX_SE3]
    # initial estimates in pinned host memory (source of the e2e H2D copies, destination of the D2H reads)
    init, host = {}, {}
    nverts = {}
    for kd in kinds:
        n = _vertex_count(ctx, kd, prob, sharded, opt)
        nverts[kd] = n
        a = ctx.estimates(kd, n)
        init[kd] = torch.from_numpy(a.copy()).pin_memory()
        host[kd] = torch.empty_like(init[kd]).pin_memory()

    def restart():
        for kd in kinds:
            ctx.set_estimates(kd, init[kd].numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.synchronize()
        torch.cuda.synchronize()

    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda") if args.workload == "sphere2500" else None  # others exceed L2
    state = {"it": 0}

    def step(e2e):
        it = state["it"] % RESTART
        if it == 0:
            restart()
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        if e2e:
            for kd in kinds:
                ctx.set_estimates(kd, init[kd].numpy() if it == 0 else host[kd].numpy())
        rc, st = ctx.algorithm_solve(g.LEVENBERG, it)
        if e2e:
            for kd in kinds:
                ctx.get_estimates_into(kd, host[kd].numpy())
        ev1.record(stream)
        ev1.synchronize()
        state["it"] += 1
        return ev0.elapsed_time(ev1), st

    def timed_run(e2e, steps, warmup):
        state["it"] = 0
        for _ in range(warmup):
            step(e2e)
        # keep the restart phase of the timed region aligned with a fresh optimisation run
        barrier()
        t = 0.0
        trials = 0
        wall0 = time.perf_counter()
        for _ in range(steps):
            ms, st = step(e2e)
            t += ms
            trials += st.levenberg_iterations
        barrier()
        wall = time.perf_counter() - wall0
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), trials, wall

    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    ms_total, trials, wall = timed_run(False, args.steps, max(args.warmup, 3))
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed_run(True, args.steps, 3)
    value = args.steps / (ms_total * 1e-3)
    e2e_value = args.steps / (ms_e2e * 1e-3)
    h2d = sum(int(init[kd].numel()) * 8 for kd in kinds)
    d2h = h2d + 8

    # ---- roofline of the dominant kernel, CUDA events on the launching stream inside this process
    ctx.set_profiling(True)
    state["it"] = 0
    for _ in range(RESTART):
        step(False)
    phases = ctx.phase_times()
    ctx.set_profiling(False)
    info = ctx.factor_info()
    roof = None
    per_phase = {k: {"ms_total": 1e3 * v[0], "launch_groups": v[1]} for k, v in phases.items() if v[1] > 0}
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = peaks.get("hbm_gbs", 6650.0)
        n_hs = lib_blocks(ctx, 3) if prob["kind"] == "ba" else 0
        ab = algorithmic_bytes(args.workload, dims, info, n_hs, dims["numEdges"])
        cand = {k: phases[k] for k in ab if phases.get(k, (0, 0))[1] > 0}
        dom = max(cand, key=lambda k: cand[k][0])
        sec = cand[dom][0] / cand[dom][1]
        achieved = ab[dom] / sec / 1e9
        kname = KERNEL_OF[dom]["ba" if prob["kind"] == "ba" else "pg"]
        roof = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": ab[dom], "avg_launch_ms": sec * 1e3, "traffic": ncu_traffic(kname),
                "others": {KERNEL_OF[k]["ba" if prob["kind"] == "ba" else "pg"]:
                           {"frac": ab[k] / (cand[k][0] / cand[k][1]) / 1e9 / peak, "avg_launch_ms": 1e3 * cand[k][0] / cand[k][1]}
                           for k in cand if k != dom}}

    # ---- the same workload with the optional ordering for parallelism (nested dissection on top of AMD): reported
    # beside the headline, which keeps the reference's AMD ordering
    nd = None
    if not args.nd_levels and not args.no_parallel_ordering and args.workload in ND_LEVELS:
        ctx.set_ordering(ND_LEVELS[args.workload])
        assert ctx.build_structure()
        restart()
        ms_nd, _, _ = timed_run(False, args.steps, max(args.warmup, 3))
        ms_nd_e2e, _, _ = timed_run(True, args.steps, 3)
        info_nd = ctx.factor_info()
        nd = {"ordering": "nested dissection, 2^%d parts, AMD inside (b200_set_ordering)" % ND_LEVELS[args.workload],
              "value": args.steps / (ms_nd * 1e-3), "ms_per_step": ms_nd / args.steps,
              "e2e": args.steps / (ms_nd_e2e * 1e-3), "unit": "iterations/s",
              "levels": info_nd["levels"], "factor_doubles": info_nd["factor_doubles"]}

    if rank == 0 and roof is not None and phases.get("chol_factor_flow", (0, 0))[1] > 0:
        # the kernel that dominates the step by TIME is the sparse factorisation: neither HBM- nor FLOP-bound but
        # bound by the latency of its dependency chain - reported for completeness
        f_s = phases["chol_factor_flow"][0] / phases["chol_factor_flow"][1]
        roof["dominant_by_time"] = {"kernel": "chol_factor_flow_kernel", "avg_launch_ms": 1e3 * f_s,
                                    "share_of_step": f_s / (ms_total * 1e-3 / args.steps),
                                    "bound": "latency of the elimination-tree dependency chain (levels: %d)" % info["levels"],
                                    "fp64_flops_per_launch": info["factor_flops"],
                                    "achieved_tflops": info["factor_flops"] / f_s / 1e12,
                                    "fp64_peak_tflops_nominal": 40.0}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, prob)

    if rank == 0:
        print(json.dumps({
            "metric": "LM iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "solver": "lm_fix6_3_b200 (Schur + supernodal Cholesky)" if prob["kind"] == "ba" else "lm_fix6_3_b200 (supernodal Cholesky)",
                       "restart_every": RESTART, "lm_trials_in_timed_region": trials,
                       "ordering": "block AMD (reference)" if not args.nd_levels else "nested dissection, 2^%d parts, AMD inside" % args.nd_levels,
                       "parallelism": ("landmark-sharded x%d, cameras replicated, NCCL all-reduce of Hschur per trial" % world) if sharded else ("replicas only" if world > 1 else "single GPU"),
                       "l2": "flushed between steps (256 MiB memset, outside the step timers)" if flush is not None else "working set (Hpl + edge arrays, resp. the factor) larger than the 126 MB L2",
                       "timing": "sum of per-step CUDA-event intervals on the solver stream, max over ranks", "wall_s": wall},
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "kernel_groups_ms_per_%d_iterations" % RESTART: per_phase,
            "factor": info,
            "parallel_ordering": nd,
            "iteration0": {"build_structure_s": t_struct, "what": "block patterns, ordering, symbolic factorisation, Schur and "
                           "Cholesky plans + their upload (host, once per graph)",
                           "cpu_port_iteration0_s": (cpu or {}).get("iteration0_s")},
        }))
    if world > 1:
        dist.destroy_process_group()


def lib_blocks(ctx, which):
    import openslam_g2o_b200 as g
    return int(g.lib.b200_get_blocks(ctx.handle, which, None, None, None))


def _vertex_count(ctx, kind, prob, sharded, opt):
    import openslam_g2o_b200 as g
    if kind == g.VERTEX_CAM:
        return len(prob["cam_ids"])
    if kind == g.VERTEX_SE3:
        return len(prob["vertex_ids"])
    # landmarks handed to this rank
    d = ctx.dims()
    return d["numVertices"] - len(prob["cam_ids"])


def cpu_baseline(workload, prob):
    """the oracle on a bounded sample of the same workload, on this box's host cores (1 thread: reference default)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import ctypes as C
        from oracle_binding import LM, Oracle, OracleStats
    except Exception as e:  # oracle missing
        return {"value": None, "unit": "iterations/s", "cores": 1, "kind": "port", "sample": "unavailable: %s" % e}
    from openslam_g2o_b200 import synth
    o = Oracle()
    synth.feed(prob, o)
    o.setup_cli(True)
    o.initialize_optimization()
    t_it0 = time.perf_counter()
    o.optimize(LM, 1)
    t_it0 = time.perf_counter() - t_it0
    o.L.oracle_lm_iteration.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    times = []
    t_all = time.time()
    it = 1
    min_it = 2 if workload in ("ba10k", "sphere40k") else 4  # bounded sample: the big graphs take seconds per CPU iteration
    while (it <= min_it or time.time() - t_all < 10.0) and time.time() - t_all < 40.0 and it < 40:
        s = OracleStats()
        t0 = time.perf_counter()
        o.L.oracle_lm_iteration(o.g, it, C.byref(s))
        times.append(time.perf_counter() - t0)
        it += 1
    ms = 1e3 * float(np.mean(times))
    return {"value": 1e3 / ms, "unit": "iterations/s", "cores": 1, "kind": "port", "ms_per_iteration": ms, "iteration0_s": t_it0,
            "sample": "LM iterations 1..%d of the same input (iteration 0 = structure + symbolic analysis excluded)" % (it - 1),
            "host_cores_available": os.cpu_count(),
            "note": "oracle = restatement of g2o's LM + BlockSolver linked to the reference's vendored CSparse (oracle/_ref); "
                    "single thread like the reference's default build (OpenMP OFF)"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world)


if __name__ == "__main__":
    main()
