#!/usr/bin/env python
"""bench.py - LM iterations/sec of the g2o hot path on B200 (see BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload venice|sphere2500|manhattan|ba10k|sphere40k|sphere1m]
                  [--impl reference] [--no-extras]

A "step" is one Levenberg-Marquardt iteration (OptimizationAlgorithmLevenberg::solve: errors + chi2, linearize +
accumulate, [Schur], sparse Cholesky, back-substitution, oplus update, re-evaluation, lambda control) over one
synthetic input graph (manhattan: one Gauss-Newton iteration, BASELINE.json configs[0]).  The optimisation is restarted
from the initial estimates every RESTART steps (like running `g2o -i 10` repeatedly) so that every step does comparable
work; the restart copy is outside the per-step timers.  Both arms follow this protocol.

  value    steps/s with all inputs resident in HBM (CUDA events on the solver's stream, max over ranks)
  e2e      steps/s through the C-ABI with HOST buffers: per step, H2D of the vertex estimates from pinned memory, one
           LM iteration, D2H of the updated estimates + chi2 (the copies a Level-3 g2o adapter makes around solve();
           the adapter's per-vertex setEstimate() walk over g2o's heap objects is NOT in it - no g2o binary exists here)
  parity   the first RESTART iterations of this very process against the oracle on the same arrays (rank 0)
  roofline the kernel that dominates the step by time first, the HBM-bound kernels beside it
  cpu_baseline  the oracle on the box's host cores (1 thread: the reference's default build), same iterations

N > 1 (torchrun): bundle adjustment shards its landmarks over the ranks (cameras replicated, two native ncclAllReduce
per LM trial) - strong scaling of the same graph; pose graphs do not shard ("replicas only").  The default run adds
short measurements of the other BASELINE.json configurations under `configs` (at N > 1: config 4, the one the
landmark sharding exists for).
"""
import argparse
import importlib.util
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
WORKLOADS = ["venice", "sphere2500", "manhattan", "venice_small", "ba10k", "sphere40k", "sphere1m"]
GN, LM = 0, 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="venice", choices=WORKLOADS)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle (no parity, no cpu_baseline)")
    ap.add_argument("--no-parallel-ordering", action="store_true", help="skip the extra nested-dissection measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other configurations")
    ap.add_argument("--nd-levels", type=int, default=0,
                    help="ordering of the reduced system: 0 = block AMD (reference, default); k = nested dissection, 2^k parts")
    return ap.parse_args()


def load_synth():
    """the generators by file path: importing the package would map libg2o_b200.so, which the reference arm must not"""
    spec = importlib.util.spec_from_file_location("g2o_b200_synth", os.path.join(ROOT, "openslam_g2o_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def manhattan_problem():
    """BASELINE.json configs[0]: the in-tree manhattanOlson3500 graph as parsed into tests/golden (the GPU box has no
    /root/reference)"""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "manhattan3500.npz"), allow_pickle=False)
    return dict(kind="se2", vertex_kind=0, edge_kind=0, vertex_ids=fx["v_ids"], vertex_payload=fx["v_pay"],
                edge_v0=fx["e_a"], edge_v1=fx["e_b"], edge_payload=fx["e_pay"])


def feed(synth, prob, target):
    if prob["kind"] == "se2":
        target.add_vertices(0, prob["vertex_ids"], prob["vertex_payload"])
        target.add_edges(0, prob["edge_v0"], prob["edge_v1"], prob["edge_payload"])
    else:
        synth.feed(prob, target)


def make_problem(synth, workload):
    if workload == "venice":
        return synth.venice_like(), "Venice-shaped BA (types_sba): 871 cameras / 530304 points / ~2.0M P2MC edges, seed 871"
    if workload == "ba10k":  # BASELINE.json configs[3] shape (one GPU here; shards by landmark under torchrun)
        return synth.venice_like(10000, 2000000, seed=10000, fixed_obs=10), "synthetic BA: 10000 cameras / 2000000 points / 20.0M P2MC edges (k = 10), seed 10000"
    if workload == "sphere40k":  # reduced BASELINE.json configs[4]: same generator, 200 x 200 instead of 1000 x 1000
        return synth.sphere(200, 200, seed=40000), "SE3 pose graph: sphere generator 200 x 200 = 40000 poses / 159399 edges, seed 40000"
    if workload == "sphere1m":   # BASELINE.json configs[4] at full size
        return synth.sphere(1000, 1000, seed=1000000), "SE3 pose graph: sphere generator 1000 x 1000 = 1000000 poses / 3995999 edges, seed 1000000"
    if workload == "venice_small":
        return synth.venice_like(100, 20000, seed=7), "small BA: 100 cameras / 20000 points"
    if workload == "manhattan":
        return manhattan_problem(), "manhattanOlson3500 (data/2d): 3500 SE2 poses / 5598 edges, Gauss-Newton"
    return synth.sphere(), "sphere2500 SE3 pose graph: 2500 poses / 9799 edges (create_sphere.cpp defaults), seed 2500"


def algorithm_of(workload):
    return GN if workload == "manhattan" else LM


def oracle_iterations(workload):
    """bounded oracle sample per workload (iterations incl. iteration 0); 0 = the reference cannot run it"""
    return {"ba10k": 2, "sphere40k": 3, "sphere1m": 0}.get(workload, RESTART)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index
        self.mark = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def wait_first(self, timeout=5.0):
        """the poller needs about a second to deliver its first line: the timed region starts after it"""
        t0 = time.time()
        while self.proc and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.05)

    def begin_region(self):
        self.mark = len(self.samples)

    def stop(self):
        if self.proc:
            time.sleep(0.12)  # one more sample after the region
            self.proc.terminate()
        region = self.samples[max(self.mark - 1, 0):]
        sm = [float(s[0]) for s in region if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in region if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in region:
            if len(s) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ oracle (checker / CPU arm)
def oracle_run(synth, prob, workload, iters, restarts=1, budget_s=1e9):
    """`restarts` optimisation runs of `iters` iterations each on the oracle, one iteration per call, timed.
    Returns per-iteration chi2 / lambda / LM trials of the first run, the wall time of every iteration and the oracle."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle, OracleStats
    algo = algorithm_of(workload)
    out = dict(chi2=[], lam=[], lev=[], times=[], iteration0_s=[])
    t_all = time.time()
    o = None
    for r in range(restarts):
        o = Oracle()
        feed(synth, prob, o)
        o.setup_cli(True)
        o.initialize_optimization()
        L = o.L
        L.oracle_iteration.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        for it in range(iters):
            s = OracleStats()
            L.oracle_iteration(o.g, algo, it, C.byref(s))
            dt = s.time_iteration   # the solve() call alone, like G2OBatchStatistics::timeIteration
            if it == 0:
                out["iteration0_s"].append(dt)   # carries buildStructure + symbolic analysis
            else:
                out["times"].append(dt)
            if r == 0:
                out["chi2"].append(s.chi2); out["lam"].append(s.lambda_); out["lev"].append(s.levenberg_iterations)
        if time.time() - t_all > budget_s:
            break
    out["oracle"] = o
    return out


def run_reference(args, rank):
    """The reference's own CPU implementation of the path: the oracle port (Eigen-free restatement of
    SparseOptimizer/BlockSolver/LM) linked against the reference's vendored CSparse compiled in oracle/_ref.
    Single thread: the reference builds with OpenMP OFF (CMakeLists.txt:137) and CSparse is sequential.
    Same protocol as the B200 arm: optimisation runs of RESTART iterations from the initial estimates; iteration 0 of
    every run additionally carries buildStructure + the symbolic analysis (the B200 arm keeps its structure across
    restarts), so it is counted as warm-up and the timed steps are iterations 1..RESTART-1 of each run."""
    if rank != 0:
        return
    synth = load_synth()
    prob, desc = make_problem(synth, args.workload)
    want = args.warmup + args.steps
    iters = RESTART if oracle_iterations(args.workload) else 0
    if iters == 0:
        print(json.dumps({"impl": "reference", "unavailable": "the reference's CSparse path indexes the factor with 32-bit ints: nnz(L) of this graph exceeds 2^31"}))
        return
    per_run = iters - 1
    restarts = max(1, -(-want // per_run))
    res = oracle_run(synth, prob, args.workload, iters, restarts=restarts, budget_s=150.0)
    times = res["times"]
    timed = times[min(args.warmup, max(len(times) - 1, 0)):] or times
    ms = 1e3 * float(np.mean(timed))
    value = 1e3 / ms
    sample = "%d iterations (iterations 1..%d of %d successive %d-iteration runs from the initial estimates; iteration 0 of each " \
             "run = structure + symbolic analysis, %.2f s, excluded) after %d warm-up" % (
                 len(timed), per_run, len(res["iteration0_s"]), iters, float(np.mean(res["iteration0_s"])), len(times) - len(timed))
    print(json.dumps({
        "impl": "reference", "metric": "LM iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": len(timed), "warmup": len(times) - len(timed), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "solver": "%s_fix%s (CSparse flavour, block AMD)" % ("gn" if algorithm_of(args.workload) == GN else "lm", "3_2" if args.workload == "manhattan" else "6_3"),
                   "restart_every": RESTART},
        "chi2_first_run": res["chi2"],
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "note": "oracle restatement of g2o's LM/BlockSolver + the reference's vendored CSparse compiled from /root/reference (oracle/_ref); "
                                 "not the CHOLMOD flavour (SuiteSparse and Eigen are absent from this image and from the GPU box: profiles/r02_box_probe.txt)"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
def algorithmic_bytes(prob_kind, dims, info, n_hs, n_edges):
    """per-launch ALGORITHMIC bytes of the HBM-bound kernels, exactly SURVEY.md section 8d (read-once / write-once model,
    f64 = 8 B, int = 4 B); plan indices, padding and partial sums are traffic, not algorithm"""
    if prob_kind.startswith("ba"):
        nl, n_hpl = dims["numLandmarks"], info["hpl_slots"]
        return {
            # P2MC edge: ids 8 + measurement 16 read, Hpl block 144 written; per point: estimate 24 read, Hll 72 + b_l 24 written
            "linearize": n_edges * 168 + nl * 120,
            # Schur, per landmark with k observations: 144 k + Hll 72 + b_l 24 read, Dinv 72 written (that part belongs to
            # schur_landmark_inverse); schur_range reads every Hpl block once + the landmark's transformed 3x3 (+u): 80 B,
            # and Hschur is written once: 288 B per block
            "schur": n_hpl * 144 + nl * 80 + n_hs * 288,
            # back-substitution: Hpl 144 B per block + Dinv 72 + b_l 24 per point read, x_l 24 written
            "backsub": n_hpl * 144 + nl * 120,
        }
    if prob_kind == "se2":   # SE2 edge: ids 8 + meas 24 + information 48 + 2 poses 48 read; 72-byte off-diagonal block written
        return {"linearize": n_edges * (128 + 72) + dims["numPoses"] * 96}
    # SE3 edge: ids 8 + measurement 56 + information 168 + 2 poses 192 = 424 B read, 288 B block written; + 336 B per vertex
    return {"linearize": n_edges * 712 + dims["numPoses"] * 336}


KERNEL_OF = {"linearize": {"ba": "ba_linearize_packets_kernel", "pg": "pg_linearize_kernel"},
             "schur": {"ba": "schur_range_kernel"}, "backsub": {"ba": "ba_backsub_kernel"}}


def ncu_traffic(kernel):
    """dram bytes (read + write) per launch of `kernel` from the committed ncu --set full summary, or None"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(kernel.split("<")[0])
    except Exception:
        return None


def fp64_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
        return float(d["dgemm_8192_sustained_tflops"]), "profiles/fp64_peak.json: cuBLAS DGEMM 8192^3 measured on this pool's B200 (sustained; tcgen05 has no FP64 kind, the FP64 tensor and vector rates are equal)"
    except Exception:
        return 37.2, "fallback: 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz (no measured DGEMM figure found)"


class Job:
    """one workload on this rank: optimizer, context, pinned host mirrors of the estimates"""

    def __init__(self, args, workload, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import openslam_g2o_b200 as g
        from openslam_g2o_b200 import synth
        self.g, self.torch, self.dist, self.synth = g, torch, dist, synth
        self.workload, self.rank, self.world, self.local_rank = workload, rank, world, local_rank
        self.algo = algorithm_of(workload)
        self.prob, self.desc = make_problem(synth, workload)
        self.sharded = world > 1 and self.prob["kind"] == "ba"
        opt = g.SparseOptimizer(device=local_rank, shard=rank if self.sharded else 0, num_shards=world if self.sharded else 1)
        opt.set_algorithm("gn_fix3_2" if self.algo == GN else "lm_fix6_3")
        feed(synth, self.prob, opt)
        opt.setup_cli()
        opt.initialize_optimization()
        opt._ensure_uploaded()
        self.opt, self.ctx = opt, opt.context
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", local_rank))
        if self.sharded:
            from openslam_g2o_b200.distributed import init_native_comm
            init_native_comm(self.ctx, rank, world)
        if args.nd_levels:
            self.ctx.set_ordering(args.nd_levels)
        t = time.perf_counter()
        assert self.ctx.build_structure()
        self.t_struct = time.perf_counter() - t  # iteration-0 cost, reported beside the steady-state metric
        self.dims = self.ctx.dims()
        kind = self.prob["kind"]
        self.kinds = [g.VERTEX_CAM, g.VERTEX_XYZ] if kind == "ba" else [g.VERTEX_SE2] if kind == "se2" else [g.VERTEX_SE3]
        self.init, self.host = {}, {}
        for kd in self.kinds:
            n = self._vertex_count(kd)
            a = self.ctx.estimates(kd, n)
            self.init[kd] = torch.from_numpy(a.copy()).pin_memory()
            self.host[kd] = torch.empty_like(self.init[kd]).pin_memory()
        # pose graphs smaller than the 126 MB L2 are flushed between steps; the others exceed it
        small = workload in ("sphere2500", "manhattan", "venice_small")
        self.flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda") if small else None
        self.it = 0

    def _vertex_count(self, kind):
        g = self.g
        if kind == g.VERTEX_CAM:
            return len(self.prob["cam_ids"])
        if kind in (g.VERTEX_SE3, g.VERTEX_SE2):
            return len(self.prob["vertex_ids"])
        return self.dims["numVertices"] - len(self.prob["cam_ids"])  # landmarks handed to this rank

    def restart(self):
        for kd in self.kinds:
            self.ctx.set_estimates(kd, self.init[kd].numpy())

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.ctx.synchronize()
        self.torch.cuda.synchronize()

    def step(self, e2e):
        torch, ctx = self.torch, self.ctx
        it = self.it % RESTART
        if it == 0:
            self.restart()
        if self.flush is not None:
            self.flush.zero_()
            torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        if e2e:
            for kd in self.kinds:
                ctx.set_estimates(kd, self.init[kd].numpy() if it == 0 else self.host[kd].numpy())
        rc, st = ctx.algorithm_solve(self.algo, it)
        if e2e:
            for kd in self.kinds:
                ctx.get_estimates_into(kd, self.host[kd].numpy())
        ev1.record(self.stream)
        ev1.synchronize()
        self.it += 1
        return ev0.elapsed_time(ev1), st

    def timed_run(self, e2e, steps, warmup, sampler=None):
        torch = self.torch
        self.it = 0
        for _ in range(warmup):
            self.step(e2e)
        self.barrier()  # the warm-up count is a multiple of nothing in particular: the restart phase just continues
        if sampler:
            sampler.begin_region()
        t, trials = 0.0, 0
        wall0 = time.perf_counter()
        for _ in range(steps):
            ms, st = self.step(e2e)
            t += ms
            trials += max(st.levenberg_iterations, 1)
        self.barrier()
        wall = time.perf_counter() - wall0
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item()), trials, wall

    def first_run(self, iters):
        """iterations 0..iters-1 from the initial estimates: chi2 / lambda / LM trials + the final state (the parity
        sample; doubles as warm-up of the CUDA graphs)"""
        self.restart()
        chi, lam, lev = [], [], []
        for it in range(iters):
            rc, st = self.ctx.algorithm_solve(self.algo, it)
            if self.algo == GN:
                chi.append(self.ctx.compute_active_errors())
            else:
                chi.append(st.chi2)
            lam.append(st.lambda_); lev.append(st.levenberg_iterations)
        self.ctx.synchronize()
        return dict(chi2=chi, lam=lam, lev=lev)

    def close(self):
        self.flush = None
        self.opt.close()
        self.torch.cuda.empty_cache()


def parity_block(job, mine, ora):
    """GPU trajectory of this process vs the oracle's on the same arrays (tolerance of the path: 1e-6 relative)"""
    g, o = job.g, ora["oracle"]
    k = min(len(mine["chi2"]), len(ora["chi2"]))
    cg, co = np.array(mine["chi2"][:k]), np.array(ora["chi2"][:k])
    out = {"iterations": k, "chi2_rel": float(np.max(np.abs(cg - co) / np.abs(co))), "chi2_final": float(cg[-1]),
           "chi2_final_oracle": float(co[-1]), "tolerance": 1e-6}
    if job.algo == LM:
        out["lambda_rel"] = float(np.max(np.abs(np.array(mine["lam"][:k]) - np.array(ora["lam"][:k])) / np.abs(ora["lam"][:k])))
        out["lm_trials_equal"] = list(mine["lev"][:k]) == list(ora["lev"][:k])
    out["ordering_bit_exact"] = bool(np.array_equal(job.ctx.block_ordering(), o.block_perm()))
    # final state: every pose + a sample of the landmarks this rank owns
    job.opt.sync_estimates()
    ids, kinds, _, _ = o.vertices()
    pose_ids = [int(i) for i, kd in zip(ids, kinds) if kd != 3]
    pt_ids = [int(i) for i, kd in zip(ids, kinds) if kd == 3][::16]
    worst = 0.0
    for group, is_pt in ((pose_ids[:: max(1, len(pose_ids) // 20000)], False), (pt_ids, True)):
        if not group:
            continue
        eg = np.stack([np.pad(job.opt.vertex_estimate(i), (0, 12))[:12] for i in group])
        eo = np.stack([np.pad(o.vertex_estimate(i), (0, 12))[:12] for i in group])
        if is_pt and job.sharded:   # landmarks of other shards keep their initial host estimate on this rank
            init = {int(i): p for i, p in zip(job.prob["point_ids"], job.prob["point_payload"])}
            own = np.array([not np.array_equal(e[:3], init[i][:3]) for e, i in zip(eg, group)])
            eg, eo = eg[own], eo[own]
        if len(eg):
            worst = max(worst, float(np.abs(eg - eo).max() / max(np.abs(eo).max(), 1e-300)))
    out["state_rel"] = worst
    out["ok"] = bool(out["chi2_rel"] < 1e-6 and out["state_rel"] < 1e-6 and out["ordering_bit_exact"])
    return out


def measure(args, workload, rank, world, local_rank, steps, warmup, full):
    """one workload: parity sample, value, e2e, roofline; `full` adds the clock sampler, the nested-dissection extra and
    the long oracle sample"""
    job = Job(args, workload, rank, world, local_rank)
    ctx, prob = job.ctx, job.prob
    n_or = 0 if args.no_cpu_baseline else oracle_iterations(workload)
    if not full and workload == "ba10k" and world > 1:
        n_or = 0  # config 4 under torchrun: the chi2 trajectory is printed, the oracle sample is in the N = 1 line
    huge = workload == "sphere1m"   # seconds per iteration: the first run is the warm-up, a handful of timed steps
    k_first = n_or if n_or else (2 if huge else min(RESTART, 3 if workload in ("ba10k", "sphere40k") else RESTART))
    mine = job.first_run(k_first)
    parity = cpu = None
    if rank == 0 and n_or:
        try:
            ora = oracle_run(job.synth, prob, workload, n_or)
            parity = parity_block(job, mine, ora)
            t = ora["times"]
            cpu = {"value": 1.0 / float(np.mean(t)), "unit": "iterations/s", "cores": 1, "kind": "port",
                   "ms_per_iteration": 1e3 * float(np.mean(t)), "iteration0_s": ora["iteration0_s"][0],
                   "sample": "iterations 1..%d of the same input (iteration 0 = structure + symbolic analysis excluded)" % (n_or - 1),
                   "host_cores_available": os.cpu_count(),
                   "note": "oracle = restatement of g2o's LM + BlockSolver linked to the reference's vendored CSparse (oracle/_ref); "
                           "single thread like the reference's default build (OpenMP OFF)"}
            ora.pop("oracle")
        except Exception as e:  # oracle missing or failed: say so, keep the measurement
            parity = {"unavailable": repr(e)}
    elif rank == 0 and not args.no_cpu_baseline and oracle_iterations(workload) == 0:
        parity = {"unavailable": "the reference's CSparse/CHOLMOD-int path indexes nnz(L) with 32-bit ints; this graph's factor has more than 2^31 entries"}
    if world > 1:
        job.dist.barrier()

    sampler = ClockSampler(local_rank) if (full and rank == 0) else None
    if sampler:
        sampler.start()
        sampler.wait_first()
    launches0 = ctx.launch_count()
    warm = 0 if huge else max(warmup, 3)
    ms_total, trials, wall = job.timed_run(False, steps, warm, sampler)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    e2e_steps = 1 if huge else steps
    ms_e2e, _, _ = job.timed_run(True, e2e_steps, 0 if huge else 3)
    value = steps / (ms_total * 1e-3)
    e2e_value = e2e_steps / (ms_e2e * 1e-3)
    h2d = sum(int(job.init[kd].numel()) * 8 for kd in job.kinds)
    d2h = h2d + 8

    # ---- per-kernel-group CUDA events inside this process (profiling pass: individual launches, graphs off)
    ctx.set_profiling(True)
    job.it = 0
    nprof = 1 if huge else RESTART
    for _ in range(nprof):
        job.step(False)
    phases = ctx.phase_times()
    ctx.set_profiling(False)
    info = ctx.factor_info()
    roof = None
    per_phase = {k: {"ms_total": 1e3 * v[0], "launch_groups": v[1]} for k, v in phases.items() if v[1] > 0}
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = peaks.get("hbm_gbs", 6650.0)
        kind = "ba" if prob["kind"].startswith("ba") else "pg"
        n_hs = lib_blocks(ctx, 3) if kind == "ba" else 0
        ab = algorithmic_bytes(prob["kind"], job.dims, info, n_hs, job.dims["numEdges"])
        hbm = {}
        for k2 in ab:
            if phases.get(k2, (0, 0))[1] > 0:
                sec = phases[k2][0] / phases[k2][1]
                name = KERNEL_OF[k2][kind]
                hbm[name] = {"achieved": ab[k2] / sec / 1e9, "frac": ab[k2] / sec / 1e9 / peak, "avg_launch_ms": 1e3 * sec,
                             "algorithmic_bytes_per_launch": ab[k2], "traffic": ncu_traffic(name)}
        step_s = ms_total * 1e-3 / steps
        fkeys = [k2 for k2 in ("chol_chain", "chol_factor_flow") if phases.get(k2, (0, 0))[1] > 0]
        f_s = sum(phases[k2][0] / phases[k2][1] for k2 in fkeys)
        tpeak, tsrc = fp64_peak()
        factor_entry = None
        if f_s > 0:
            factor_entry = {"bound": "tensor", "kernel": "+".join("%s_kernel" % k2 for k2 in fkeys),
                            "achieved": info["factor_flops"] / f_s / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                            "frac": info["factor_flops"] / f_s / 1e12 / tpeak, "traffic": ncu_traffic("chol_factor_flow_kernel"),
                            "avg_launch_ms": 1e3 * f_s, "share_of_step": f_s / step_s, "flops_per_launch": info["factor_flops"],
                            "peak_source": tsrc,
                            "note": ("FP64 tensor path: update products of the supernodal factorisation on mma.sync.m8n8k4.f64 (DMMA), 96 x 72 "
                                     "destination tiles, cp.async operand pipeline (levels: %d)" % info["levels"]) if info.get("wide_tiles") else
                                    ("FP64 pipe (DFMA) of the sparse supernodal factorisation; on these graphs the kernel is bound by the "
                                     "latency of the elimination-tree dependency chain (levels: %d), not by the pipe" % info["levels"])}
        best_hbm = max(hbm, key=lambda n: hbm[n]["avg_launch_ms"]) if hbm else None
        if factor_entry and (best_hbm is None or factor_entry["avg_launch_ms"] >= hbm[best_hbm]["avg_launch_ms"]):
            roof = dict(factor_entry)
        elif best_hbm:
            e = hbm[best_hbm]
            roof = {"bound": "hbm", "kernel": best_hbm, "achieved": e["achieved"], "peak": peak, "unit": "GB/s", "frac": e["frac"],
                    "traffic": e["traffic"], "avg_launch_ms": e["avg_launch_ms"], "share_of_step": e["avg_launch_ms"] * 1e-3 / step_s,
                    "algorithmic_bytes_per_launch": e["algorithmic_bytes_per_launch"], "fp64_kernel": factor_entry}
        if roof is not None:
            roof["hbm_peak"] = peak
            roof["hbm_peak_source"] = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"
            roof["hbm_kernels"] = hbm

    # ---- the same workload with the optional ordering for parallelism (nested dissection on top of AMD): reported
    # beside the headline, which keeps the reference's AMD ordering
    nd = None
    if full and not args.nd_levels and not args.no_parallel_ordering and workload in ND_LEVELS:
        ctx.set_ordering(ND_LEVELS[workload])
        assert ctx.build_structure()
        job.restart()
        ms_nd, _, _ = job.timed_run(False, steps, max(warmup, 3))
        ms_nd_e2e, _, _ = job.timed_run(True, steps, 3)
        info_nd = ctx.factor_info()
        nd = {"ordering": "nested dissection, 2^%d parts, AMD inside (b200_set_ordering)" % ND_LEVELS[workload],
              "value": steps / (ms_nd * 1e-3), "ms_per_step": ms_nd / steps,
              "e2e": steps / (ms_nd_e2e * 1e-3), "unit": "iterations/s",
              "levels": info_nd["levels"], "factor_doubles": info_nd["factor_doubles"]}

    res = None
    if rank == 0:
        ba = prob["kind"].startswith("ba")
        res = {
            "metric": "LM iterations/sec" if job.algo == LM else "GN iterations/sec", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": steps, "warmup": warm if not huge else k_first, "ms_per_step": ms_total / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic" if workload != "manhattan" else "in-tree dataset (tests/golden)",
            "config": {"workload": job.desc,
                       "solver": ("lm_fix6_3_b200 (Schur + supernodal Cholesky)" if ba else "gn_fix3_2_b200 (supernodal Cholesky)" if job.algo == GN else "lm_fix6_3_b200 (supernodal Cholesky)"),
                       "restart_every": RESTART, "lm_trials_in_timed_region": trials,
                       "ordering": "block AMD (reference)" if not args.nd_levels else "nested dissection, 2^%d parts, AMD inside" % args.nd_levels,
                       "parallelism": ("landmark-sharded x%d, cameras replicated, 2 native ncclAllReduce per LM trial ([Hschur|bschur|chi2], 2 scalars) in the trial CUDA graph" % world) if job.sharded else ("replicas only" if world > 1 else "single GPU"),
                       "l2": "flushed between steps (256 MiB memset, outside the step timers)" if job.flush is not None else "working set (Hpl + edge arrays, resp. the factor) larger than the 126 MB L2",
                       "timing": "sum of per-step CUDA-event intervals on the solver stream, max over ranks", "wall_s": wall},
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "chi2_first_run": mine["chi2"],
            "roofline": roof, "cpu_baseline": cpu,
            "kernel_groups_ms_per_%d_iterations" % nprof: per_phase,
            "factor": info,
            "parallel_ordering": nd,
            "iteration0": {"build_structure_s": job.t_struct, "what": "block patterns, ordering, symbolic factorisation, Schur and "
                           "Cholesky plans + their upload (host, once per graph)",
                           "cpu_port_iteration0_s": (cpu or {}).get("iteration0_s")},
        }
    job.close()
    return res


def lib_blocks(ctx, which):
    import openslam_g2o_b200 as g
    return int(g.lib.b200_get_blocks(ctx.handle, which, None, None, None))


def slim(r):
    """short form of a measurement for the `configs` section"""
    if r is None:
        return None
    keep = ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "parity", "chi2_first_run", "cpu_baseline", "factor")
    out = {k: r.get(k) for k in keep}
    out["workload"] = r["config"]["workload"]
    out["parallelism"] = r["config"]["parallelism"]
    if r.get("roofline"):
        rf = r["roofline"]
        out["roofline"] = {k: rf.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "avg_launch_ms", "share_of_step")}
        out["roofline"]["hbm_kernels"] = {n: {"frac": e["frac"], "avg_launch_ms": e["avg_launch_ms"]} for n, e in rf.get("hbm_kernels", {}).items()}
    return out


def widening_block(device):
    """SURVEY 8(f) rows measured beside the headline (N = 1, short; wall clock around synchronised calls): marginal
    covariances through the supernodal sparse inverse (all diagonal blocks of sphere2500 in one sweep), and the two
    landmark-SLAM families (`lm_var`) - LM iterations/s and chi2 parity against the oracle on the same arrays."""
    import openslam_g2o_b200 as g
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle
    synth = load_synth()
    out = {}
    # ---- marginals
    opt = g.SparseOptimizer(device=device)
    opt.set_algorithm("lm_fix6_3")
    synth.feed(synth.sphere(), opt)
    opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
    ctx = opt.context
    assert ctx.build_structure()
    ctx.compute_active_errors(); ctx.build_system()
    nb = ctx.dims()["numPoses"]
    diag = [(i, i) for i in range(nb)]
    ctx.compute_marginals(diag[:4])  # plan + buffers of the sweep
    l0 = ctx.launch_count()
    t = time.perf_counter(); m_all = ctx.compute_marginals(diag); t_all = time.perf_counter() - t
    launches = ctx.launch_count() - l0
    t = time.perf_counter(); ctx.compute_marginals(diag[:8]); t_8 = time.perf_counter() - t
    t = time.perf_counter(); ctx.compute_marginals([(0, nb - 1)]); t_off = time.perf_counter() - t
    out["marginals_sphere2500"] = {"blocks": nb, "all_diagonal_blocks_ms": t_all * 1e3, "launches": launches, "eight_blocks_ms": t_8 * 1e3,
                                   "one_block_outside_the_factor_pattern_ms": t_off * 1e3, "trace_sample": float(np.trace(m_all[nb // 2])),
                                   "method": "one factorisation + supernodal sparse-inverse sweep (csrc/sparse_inverse.cuh); unit solves only outside the pattern of L"}
    opt.close()
    # ---- landmark SLAM
    for key, prob in (("slam2d_lm_var", synth.landmark_slam_2d(2000, 3000, seed=40, growth=0.05)), ("slam3d_lm_var", synth.landmark_slam_3d(1500, 1200, seed=41))):
        opt = g.SparseOptimizer(device=device)
        opt.set_algorithm("lm_var")
        o = Oracle()
        for tgt in (opt, o):
            synth.feed(prob, tgt)
        opt.setup_cli(); o.setup_cli(False); o.set_block_ordering(True)
        opt.initialize_optimization(); o.initialize_optimization()
        iters = 10
        opt.optimize(2)   # structure phase + CUDA graphs; restart from the initial state below
        opt.close()
        opt = g.SparseOptimizer(device=device)
        opt.set_algorithm("lm_var")
        synth.feed(prob, opt)
        opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
        assert opt.context.build_structure()
        opt.context.synchronize()
        t = time.perf_counter(); n = opt.optimize(iters); opt.context.synchronize(); t_g = time.perf_counter() - t
        t = time.perf_counter(); no, st = o.optimize(LM, iters); t_o = time.perf_counter() - t
        cg = np.array([s.chi2 for s in opt.batch_statistics]); co = np.array([s.chi2 for s in st[:no]])
        k = min(len(cg), len(co))
        out[key] = {"poses": len(prob["pose_ids"]), "landmarks": len(prob["lm_ids"]), "edges": len(prob["odo_v0"]) + len(prob["obs_v0"]),
                    "iterations": int(n), "lm_iterations_per_s": n / t_g, "oracle_lm_iterations_per_s": no / t_o,
                    "chi2_rel": float(np.max(np.abs(cg[:k] - co[:k]) / np.abs(co[:k]))), "chi2_final": float(cg[-1]),
                    "ordering_bit_exact": bool(np.array_equal(opt.context.block_ordering(), o.block_perm())),
                    "note": "wall clock incl. the per-trial host round trips; first iteration excludes structure + symbolic analysis (GPU) but includes it for the oracle"}
        opt.close()
    # ---- LinearSolverPCG as the linear solver of the headline graph (lm_pcg6_3): inexact solves, so a flavour of its own -
    #      reported beside the Cholesky headline, never instead of it
    try:
        prob = synth.venice_like()
        res = {}
        for name in ("lm_pcg6_3", "lm_fix6_3"):
            opt = g.SparseOptimizer(device=device)
            opt.set_algorithm(name)
            synth.feed(prob, opt)
            opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
            assert opt.context.build_structure()
            opt.optimize(1)                      # warm-up (first launches, lazy module load); a fresh optimizer is timed
            opt.close()
            opt = g.SparseOptimizer(device=device)
            opt.set_algorithm(name)
            synth.feed(prob, opt)
            opt.setup_cli(); opt.initialize_optimization(); opt._ensure_uploaded()
            assert opt.context.build_structure()
            opt.context.synchronize()
            t = time.perf_counter(); n = opt.optimize(10); opt.context.synchronize(); dt = time.perf_counter() - t
            res[name] = {"iterations": int(n), "lm_iterations_per_s_wall": n / dt, "chi2": [float(s_.chi2) for s_ in opt.batch_statistics],
                         "cg_iterations_last_solve": opt.context.linear_solver_iterations()}
            opt.close()
        out["venice_pcg_vs_cholesky"] = res
    except Exception as e:
        out["venice_pcg_vs_cholesky"] = {"error": repr(e)}
    return out


def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    main_res = measure(args, args.workload, rank, world, local_rank, args.steps, args.warmup, full=True)
    extras = {}
    if args.workload == "venice" and not args.no_extras and not args.nd_levels:
        # the other BASELINE.json configurations, short runs.  N > 1: only config 4 shards; pose graphs stay single-GPU
        plan = [("config4_ba10k", "ba10k", 20)] if world > 1 else \
               [("config1_manhattan_gn", "manhattan", 40), ("config2_sphere2500", "sphere2500", 40),
                ("config4_ba10k", "ba10k", 20), ("config5_sphere1m", "sphere1m", 2)]
        for key, wl, st in plan:
            try:
                extras[key] = slim(measure(args, wl, rank, world, local_rank, st, 3, full=False))
            except Exception as e:  # the headline line must survive a failing extra
                extras[key] = {"error": repr(e)}
                if world > 1:
                    raise
    if rank == 0 and world == 1 and args.workload == "venice" and not args.no_extras and not args.nd_levels:
        try:
            main_res["widening"] = widening_block(local_rank)
        except Exception as e:
            main_res["widening"] = {"error": repr(e)}
    if rank == 0:
        main_res["configs"] = extras or None
        print(json.dumps(main_res))
    if world > 1:
        dist.destroy_process_group()


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
