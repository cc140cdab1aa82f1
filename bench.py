#!/usr/bin/env python3
"""bench.py — SbFDDP OCP-iterations/sec on synthetic batches of the reference's named problems (BASELINE.json).

A "step" is one complete batched SbFDDP solve (every OCP of the batch to its own stopping decision) of
  hexacopter370_flying_arm_3 / displacement.yaml, dt 20 ms, Euler, squash, T = 400, B = 4096 OCPs per GPU,
  x0_b = YAML initial state + 0.05*U(-1,1) noise (seed 2024+b), zero initial guess, maxiter 100  (SURVEY.md §8d config 2).
`value`  = OCP-iterations executed by all ranks / device time, inputs already resident in HBM.
`e2e`    = the same through the C ABI with host buffers: x0 copied H2D, xs/us/cost/iters copied D2H inside the timed region.
`--impl reference` times the CPU restatement (oracle/) on all host threads on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SbFDDP OCP-iterations/sec"
UNIT = "OCP-iterations/s"
WORKLOAD = "hexacopter370_flying_arm_3_displacement"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="OCPs per GPU")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--cpu-sample", type=int, default=0, help="OCPs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mpc", action="store_true", help="skip the single-instance MPC-step latency leg")
    ap.add_argument("--mpc-steps", type=int, default=300)
    ap.add_argument("--no-config4", action="store_true", help="skip the strong-scaling leg (16384 hextilt_flying_arm_5 OCPs over all ranks)")
    ap.add_argument("--config4-batch", type=int, default=16384, help="OCPs of the strong-scaling leg, TOTAL over all ranks")
    ap.add_argument("--no-config5", action="store_true", help="skip the iris_px4 rail / weighted MPC horizon sweep (N = 1 only)")
    ap.add_argument("--no-divergent", action="store_true", help="skip the divergent-batch leg (hexacopter370 hover, N = 1 only)")
    return ap.parse_args()


def load_problem(name):
    host = importlib.import_module("eagle-mpc_b200.host")
    wl = importlib.import_module("eagle-mpc_b200.workloads")
    yaml, dt, seed0 = wl.CONFIGS[name]
    tr = host.Trajectory(yaml)
    fp = tr.createProblem(dt)
    return tr, fp, wl, seed0, dt


def oracle_binding():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob  # test infrastructure: only used for the CPU baseline legs
    return ob


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def mpc_latency(n_steps, with_cpu):
    """BASELINE.json config 3: carrot MPC closed loop on hexacopter370_flying_arm_3 (knots 30, dt 30 ms, iters 2,
    plant RK4 @ 2 ms), single instance; p50/p95 of (updateProblem + solve) per step.  The reference trajectory is solved
    with the B200 path itself (B = 1, maxiter 400)."""
    host = importlib.import_module("eagle-mpc_b200.host")
    capi = importlib.import_module("eagle-mpc_b200.capi")
    mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
    tr = host.Trajectory("hexacopter370_flying_arm_3/trajectories/displacement.yaml")
    fp = tr.createProblem(20)
    s1 = capi.BatchSolver(fp, 1)
    p = capi.default_params(); p.maxiter = 400
    s1.set_params(p); s1.set_x0(fp.x0); s1.set_candidate(None, None, False); s1.solve()
    xs, us = s1.xs()[0], s1.us()[0]
    s1.close()
    mpc_yaml = "hexacopter370_flying_arm_3/mpc/mpc.yaml"
    mpc = mpcmod.CarrotMpc(tr, xs, 20, mpc_yaml, create_solver=True)
    mpcmod.closed_loop(mpc, xs, us, xs[0], 20)  # warm-up
    lat, _, _, _ = mpcmod.closed_loop(mpc, xs, us, xs[0], n_steps)
    out = {"config": "carrot MPC, hexacopter370_flying_arm_3/mpc/mpc.yaml (knots 30, dt 30 ms, iters 2), RK4 plant 2 ms, B=1",
           "steps": n_steps, "gpu_p50_ms": float(1e3 * np.median(lat)), "gpu_p95_ms": float(1e3 * np.percentile(lat, 95))}
    if with_cpu:
        ob = oracle_binding()
        mpc_o = mpcmod.CarrotMpc(tr, xs, 20, mpc_yaml, create_solver=False)
        lat_o, _, _, _ = ob.oracle_closed_loop(mpc_o, xs, us, xs[0], n_steps)
        out["cpu_p50_ms"] = float(1e3 * np.median(lat_o)); out["cpu_p95_ms"] = float(1e3 * np.percentile(lat_o, 95))
        out["cpu_kind"] = "port (oracle/), 1 thread"
    return out


def source_hash():
    """sha256 over the CUDA sources of the library: stamps ncu-derived numbers so that they cannot go stale silently"""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "eagle-mpc_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()[:16]


def config4_leg(args, torch, dist, world, rank, local_rank, barrier):
    """BASELINE.json config 4: hextilt_flying_arm_5 push_slide (T = 100), 16384 OCPs in TOTAL, contiguous index ranges per
    rank (strong scaling).  No collective on the solver path.  The e2e leg is the data plane of SURVEY 8(e): rank 0 owns the
    host buffers; every step it uploads all initial states, scatters them over NCCL, every rank solves its shard, and the
    solutions (xs, us, cost, iterations) are gathered over NCCL to rank 0 and copied to its pinned host memory."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    sharding = importlib.import_module("eagle-mpc_b200.sharding")
    name = "hextilt_flying_arm_5_push_slide"
    tr, fp, wl, seed0, dt = load_problem(name)
    total = args.config4_batch
    b0, b1 = sharding.shard_range(total, rank, world)
    nb = b1 - b0
    nx, nu, T = fp.nx, fp.nu, fp.T
    dev = torch.device("cuda", local_rank)
    x0_mine = wl.noisy_x0(fp.x0, nb, seed0, first=b0)
    solver = capi.BatchSolver(fp, nb, device=local_rank)
    solver.set_x0(x0_mine); solver.set_candidate(None, None, False)
    solver.enable_kernel_timing(False)

    def resident():
        solver.reset(); solver.solve()
        return solver.total_iterations(), solver.solve_stats()[0]

    for _ in range(args.warmup):
        resident()
    barrier()
    t0 = time.perf_counter()
    it_res, dev_ms = 0, 0.0
    for _ in range(args.steps):
        i, ms = resident(); it_res += i; dev_ms += ms
    barrier()
    wall = time.perf_counter() - t0
    # ---- e2e with the NCCL data plane ----
    x0_pin = torch.from_numpy(wl.noisy_x0(fp.x0, total, seed0)).pin_memory() if rank == 0 else None
    if rank == 0:
        xs_pin = torch.empty((total, T + 1, nx), dtype=torch.float64).pin_memory()
        us_pin = torch.empty((total, T, nu), dtype=torch.float64).pin_memory()
        cost_pin = torch.empty((total,), dtype=torch.float64).pin_memory()
        it_pin = torch.empty((total,), dtype=torch.int32).pin_memory()
    xs_d = torch.empty((nb, T + 1, nx), dtype=torch.float64, device=dev)
    us_d = torch.empty((nb, T, nu), dtype=torch.float64, device=dev)
    nccl_bytes = [0]

    def e2e():
        x0_all = x0_pin.to(dev, non_blocking=True) if rank == 0 else None
        x0_d = sharding.scatter_rows(x0_all, total, (nx,), dist, device=dev, world=world, rank=rank)
        torch.cuda.current_stream().synchronize()
        solver.set_x0_ptr(x0_d.data_ptr())            # device pointer: the C ABI copies with cudaMemcpyDefault
        solver.set_candidate(None, None, False)
        solver.solve()
        solver.get_into("xs", xs_d.data_ptr()); solver.get_into("us", us_d.data_ptr())
        cost_d = torch.from_numpy(solver.cost()).to(dev); it_d = torch.from_numpy(solver.iters()).to(dev)
        outs = [sharding.gather_rows(t_, total, dist, world=world, rank=rank) for t_ in (xs_d, us_d, cost_d, it_d)]
        if rank == 0:
            for dst, src in zip((xs_pin, us_pin, cost_pin, it_pin), outs):
                dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        if world > 1:
            mine = nb * (nx + (T + 1) * nx + T * nu + 1) * 8 + nb * 4
            nccl_bytes[0] += 0 if rank == 0 else mine
        return solver.total_iterations()

    e2e()
    barrier()
    t1 = time.perf_counter()
    it_e2e = 0
    for _ in range(args.steps):
        it_e2e += e2e()
    barrier()
    wall_e2e = time.perf_counter() - t1
    if rank == 0:  # the gathered batch is the batch: initial states in place, finite costs
        assert torch.equal(xs_pin[:, 0], x0_pin) and bool(torch.isfinite(cost_pin).all()) and int(it_pin.min()) >= 1
    stats = torch.tensor([wall, wall_e2e, dev_ms], dtype=torch.float64, device=dev)
    work = torch.tensor([it_res, it_e2e, nccl_bytes[0]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    wall, wall_e2e, dev_ms = stats.tolist()
    it_res, it_e2e, nccl_b = work.tolist()
    solver.close()
    if rank != 0:
        return None
    ndx = fp.ndx
    D = 2 * ndx * ndx + 2 * ndx * nu + nu * nu + ndx + nu
    peaks, _k = measured_peaks()
    ceiling = float(peaks.get("hbm_gbs", 6650.0)) * 1e9 / (T * wl.algorithmic_bytes_per_node(nx, ndx, nu))
    return {"workload": name, "T": T, "dt_ms": dt, "batch_total": total, "batch_per_gpu": nb, "scaling": "strong",
            "value": it_res / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / args.steps, "device_ms_per_step": dev_ms / args.steps,
            "iterations_per_step": it_res / args.steps,
            "frac_of_hbm_ceiling": it_res / wall / world / ceiling,
            "e2e": {"value": it_e2e / wall_e2e, "unit": UNIT, "ms_per_step": 1e3 * wall_e2e / args.steps,
                    "h2d_bytes_per_step": total * nx * 8, "d2h_bytes_per_step": total * ((T + 1) * nx + T * nu + 1) * 8 + total * 4,
                    "nccl_bytes_per_step": nccl_b / args.steps,
                    "data_plane": "rank 0 pinned host -> H2D -> NCCL scatter of x0 -> solve -> NCCL gather of xs/us/cost/iters -> D2H on rank 0",
                    "outputs": "xs, us, cost, iters"}}


def config5_leg(args):
    """BASELINE.json config 5: iris_px4 Rail and Weighted MPC, `knots` in {50, 100, 200, 400}, B = 1024 instances of one MPC
    problem warm-started from the trajectory slice (x0 = trajectory state + the benchmark's noise, seeds 9000+b), mpc.yaml
    iters (2) SbFDDP iterations each; reference trajectory = iris_px4 displacement solved by the B200 path (maxiter 400)."""
    import tempfile
    host = importlib.import_module("eagle-mpc_b200.host")
    capi = importlib.import_module("eagle-mpc_b200.capi")
    mpcmod = importlib.import_module("eagle-mpc_b200.mpc")
    wl = importlib.import_module("eagle-mpc_b200.workloads")
    traj = "iris_px4/trajectories/displacement.yaml"
    tr = host.Trajectory(traj); fp = tr.createProblem(20)
    s1 = capi.BatchSolver(fp, 1)
    p = capi.default_params(); p.maxiter = 400
    s1.set_params(p); s1.set_x0(fp.x0); s1.set_candidate(None, None, False); s1.solve()
    xs, us = s1.xs()[0], s1.us()[0]
    s1.close()
    tmpdir = tempfile.mkdtemp()
    B, t0_ms, rows = 1024, 1000, []
    for kind in ("rail", "weighted"):
        for knots in (50, 100, 200, 400):
            y = os.path.join(tmpdir, f"mpc_{knots}.yaml")
            open(y, "w").write(open(os.path.join(ROOT, "yaml", "iris_px4", "mpc", "mpc.yaml")).read().replace("knots: 40", f"knots: {knots}"))
            mpc = (mpcmod.RailMpc(xs, 20, y, create_solver=False) if kind == "rail"
                   else mpcmod.WeightedMpc(host.Trajectory(traj), 20, y, create_solver=False))
            mpc.updateProblem(t0_ms)
            T = mpc.knots - 1
            i0 = t0_ms // 20
            idx = np.minimum(i0 + np.arange(T + 1), len(xs) - 1)
            x0 = wl.noisy_x0(xs[i0], B, 9000)
            g = capi.BatchSolver(mpc, B)
            costs, pool = mpc.cost_tables()
            g.update_costs(0, costs, 0, pool)
            pr = capi.default_params(); pr.maxiter = mpc.iters; pr.convergence_init = 1e-3
            g.set_params(pr)
            xs_b = np.broadcast_to(xs[idx], (B, T + 1, xs.shape[1])).copy(); xs_b[:, 0] = x0
            us_b = np.broadcast_to(us[np.minimum(idx[:-1], len(us) - 1)], (B, T, us.shape[1])).copy()
            g.set_x0(x0); g.set_candidate(xs_b, us_b, False)
            for _ in range(3):
                g.reset(); g.solve()
            ms, its = 0.0, 0
            for _ in range(5):
                g.reset(); g.solve()
                ms += g.solve_stats()[0]; its += g.total_iterations()
            rows.append({"controller": kind, "knots": knots, "T": T, "iters_per_instance": its / 5 / B,
                         "device_ms_per_batched_mpc_step": ms / 5, "ocp_iterations_per_s": its / (ms * 1e-3),
                         "us_per_instance_step": 1e3 * ms / 5 / B})
            g.close()
    return {"workload": "iris_px4 rail / weighted MPC horizon sweep (yaml/iris_px4/mpc/mpc.yaml, knots overridden), B = 1024 warm-started instances",
            "batch": B, "unit": UNIT, "timing": "CUDA events around each batched solve, 3 warm-ups, mean of 5", "cases": rows}


def divergent_leg(args):
    """A batch whose OCPs need very different numbers of iterations (hexacopter370 hover from noisy initial states: a few
    iterations for some OCPs, 100+ crawling ones for others).  Every kernel masks finished OCPs, but the batch runs until its
    slowest OCP is done, and the sequential kernels (Riccati sweep, rollouts) cost their per-OCP latency however few OCPs
    are left: the leg reports the throughput, the straggler factor (batch-iterations x batch / OCP-iterations executed) and
    how the batch-iterations split by the share of OCPs still active."""
    capi = importlib.import_module("eagle-mpc_b200.capi")
    tr, fp, wl, seed0, dt = load_problem("hexacopter370_hover")
    B = 4096
    x0 = wl.noisy_x0(fp.x0, B, seed0)
    g = capi.BatchSolver(fp, B)
    g.set_x0(x0); g.set_candidate(None, None, False)
    g.enable_kernel_timing(False)
    for _ in range(2):
        g.reset(); g.solve()
    ms, its, launches = 0.0, 0, 0
    for _ in range(3):
        g.reset(); g.solve()
        ms += g.solve_stats()[0]; its += g.total_iterations(); launches += g.launch_stats()[0]
    it = g.iters() + 1
    batch_iters = (launches / 3 - 2) / 8
    # the same workload through empc_solve_stream: 4 x B jobs through the B slots, refilled on the device as OCPs finish
    jobs = 4 * B
    x0_jobs = wl.noisy_x0(fp.x0, jobs, seed0)
    g.solve_stream(x0_jobs[:2 * B], want_trajectories=False)   # warm-up
    out = g.solve_stream(x0_jobs, want_trajectories=True)
    ms_stream = g.solve_stats()[0]; its_stream = g.total_iterations()
    assert int((out["iters"][:B] + 1).sum()) == int(it.sum())   # the first B jobs are the plain batch's OCPs: same iteration counts
    # ... and the uniform yardstick: every OCP cut at the same small maxiter, so that (nearly) all of them are active in every batch-iteration
    pu = capi.default_params(); pu.maxiter = 6
    g.set_params(pu); g.set_x0(x0); g.set_candidate(None, None, False)
    g.reset(); g.solve()
    g.reset(); g.solve()
    ms_uni = g.solve_stats()[0]; its_uni = g.total_iterations()
    g.close()
    alive = [(it > k).mean() for k in range(int(it.max()))]   # share of OCPs still iterating at batch-iteration k
    return {"workload": "hexacopter370_hover, B = 4096, x0 = YAML state + 0.05*U(-1,1), seeds 1000+b", "T": fp.T, "batch": B,
            "value": its / (ms * 1e-3), "unit": UNIT, "device_ms_per_solve": ms / 3,
            "iterations_per_ocp": {"min": int(it.min()), "median": float(np.median(it)), "mean": float(it.mean()), "p95": float(np.percentile(it, 95)), "max": int(it.max())},
            "batch_iterations": batch_iters, "straggler_factor": batch_iters * B / (its / 3),
            "batch_iterations_by_active_share": {">50%": int(sum(a > 0.5 for a in alive)), "5-50%": int(sum(0.05 < a <= 0.5 for a in alive)),
                                                 "<5%": int(sum(a <= 0.05 for a in alive))},
            "ms_per_batch_iteration": ms / 3 / batch_iters,
            "stream": {"what": "empc_solve_stream: %d jobs through the %d slots, refilled on the device as OCPs finish" % (jobs, B),
                       "value": its_stream / (ms_stream * 1e-3), "unit": UNIT, "device_ms": ms_stream, "ocp_iterations": its_stream,
                       "speedup_over_plain_batch": (its_stream / ms_stream) / (its / ms)},
            "uniform_yardstick": {"what": "the same batch with every OCP cut at maxiter 6 per pass (all OCPs active in nearly every batch-iteration)",
                                  "value": its_uni / (ms_uni * 1e-3), "unit": UNIT},
            "per_iteration_cost_vs_uniform": {"plain_batch": (its_uni / ms_uni) / (its / ms), "stream": (its_uni / ms_uni) / (its_stream / ms_stream)}}


def run_reference(args):
    """CPU arm: the in-repo restatement (oracle/) on all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tr, fp, wl, seed0, dt = load_problem(args.workload)
    ob = oracle_binding()
    cores = os.cpu_count() or 1
    n = args.cpu_sample or 32 * cores   # ~5 s of CPU work per step on this workload (bounded sample of the 4096-OCP batch)
    n = min(n, 1024)
    x0 = wl.noisy_x0(fp.x0, n, seed0)
    for _ in range(max(args.warmup, 0)):
        ob.solve_batch(fp, x0[:cores], cores)
    tot_it, tot_s = 0, 0.0
    for _ in range(args.steps):
        sec, it, _c = ob.solve_batch(fp, x0, cores)
        tot_it += int(it.sum()); tot_s += sec
    v = tot_it / tot_s
    sample = f"{n} OCPs of {args.workload} (B=4096 workload, seeds {seed0}..{seed0 + n - 1}) per step, {cores} threads"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "T": fp.T, "dt_ms": dt, "batch_per_step": n,
                      "note": "CPU restatement of the reference algorithm (oracle/), not Crocoddyl itself"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    capi = importlib.import_module("eagle-mpc_b200.capi")
    tr, fp, wl, seed0, dt = load_problem(args.workload)
    B = args.batch
    nx, nu, ndx, T = fp.nx, fp.nu, fp.ndx, fp.T
    # weak scaling: every rank solves its own B OCPs (seeds continue across ranks); no collective on the solver path
    x0 = wl.noisy_x0(fp.x0, B, seed0, first=rank * B)
    solver = capi.BatchSolver(fp, B, device=local_rank)
    solver.set_x0(x0)
    solver.set_candidate(None, None, False)
    solver.enable_kernel_timing(False)  # timed steps use the pipelined multi-stream schedule

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host buffers for the e2e leg
    x0_pin = torch.from_numpy(x0.copy()).pin_memory()
    xs_pin = torch.empty((B, T + 1, nx), dtype=torch.float64).pin_memory()
    us_pin = torch.empty((B, T, nu), dtype=torch.float64).pin_memory()
    cost_pin = torch.empty((B,), dtype=torch.float64).pin_memory()
    iters_pin = torch.empty((B,), dtype=torch.int32).pin_memory()

    def step_resident():
        solver.reset()
        solver.solve()
        return solver.total_iterations()

    def step_e2e():
        solver.set_x0_ptr(x0_pin.data_ptr())
        solver.set_candidate(None, None, False)
        solver.solve()
        solver.get_into("xs", xs_pin.data_ptr())
        solver.get_into("us", us_pin.data_ptr())
        solver.get_into("cost", cost_pin.data_ptr())
        solver.get_into("iters", iters_pin.data_ptr())
        return solver.total_iterations()

    for _ in range(args.warmup):
        step_resident()
    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    iters = 0
    dev_ms = 0.0
    launches = 0
    ms_k = np.zeros(4)
    units_k = np.zeros(4, dtype=np.int64)
    for _ in range(args.steps):
        iters += step_resident()
        ms, _units = solver.solve_stats()
        dev_ms += ms
        n_l, _mk = solver.launch_stats()
        launches += n_l
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # one extra (untimed) step on the serial schedule with CUDA events around every launch: per-kernel device times
    # and the number of OCPs each kernel family processed, for the roofline of the dominant kernel
    solver.enable_kernel_timing(True)
    step_resident()
    _ms, units_k = solver.solve_stats()
    launches_serial, ms_k = solver.launch_stats()
    solver.enable_kernel_timing(False)
    # ---- e2e: host buffers through the C ABI ----
    step_e2e()
    barrier()
    t1 = time.perf_counter()
    iters_e2e = 0
    for _ in range(args.steps):
        iters_e2e += step_e2e()
    barrier()
    wall_e2e = time.perf_counter() - t1
    h2d = x0_pin.numel() * 8
    d2h = (xs_pin.numel() + us_pin.numel() + cost_pin.numel()) * 8 + iters_pin.numel() * 4

    # max over ranks of the time, sum over ranks of the work
    stats = torch.tensor([wall, wall_e2e, dev_ms], dtype=torch.float64, device="cuda")
    work = torch.tensor([iters, iters_e2e, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    wall, wall_e2e, dev_ms = stats.tolist()
    iters_all, iters_e2e_all, launches_all = work.tolist()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        # roofline of the dominant kernel: algorithmic bytes (SURVEY.md §8d split per kernel family) / its device time
        D = 2 * ndx * ndx + 2 * ndx * nu + nu * nu + ndx + nu
        bytes_node = {
            "calc_diff": 8 * ((nx + nu) + (D + nx + 1)),
            "backward": 8 * ((D + ndx) + (nu * ndx + nu + ndx + ndx * ndx)),
            "rollout": 8 * ((nx + nu + nu * ndx + nu + ndx) + (nx + nu + 1)),  # one trial (the sequential reference's usual case)
        }
        names = ["calc_diff", "backward", "rollout", "decide"]
        kernels = {"calc_diff": "node_calc_kernel+node_diff_kernel", "backward": "backward_kernel",
                   "rollout": "rollout_kernel", "decide": "decide_kernel"}
        dom = int(np.argmax(ms_k[:3]))
        n_launch = max(1, (launches_serial - 2) // 8)  # batch-iterations of the instrumented step (8 launches each)
        ach = bytes_node[names[dom]] * float(units_k[dom]) * T / (ms_k[dom] * 1e-3) / 1e9
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM traffic of the dominant kernel: ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch at this
        # workload (ncu cannot run inside a timed bench).  The capture is stamped with the hash of the CUDA sources it
        # was taken from; a capture of other sources is refused (null) rather than reported stale.
        traffic, traffic_note = None, "no ncu capture of the current kernels (profiles/r2_traffic.json)"
        try:
            tr_json = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if tr_json.get("source_hash") != source_hash():
                traffic_note = "profiles/r2_traffic.json was captured from other kernel sources (%s): refused" % tr_json.get("source_hash")
            elif tr_json.get("workload") == args.workload and tr_json.get("batch") == B:
                traffic = tr_json["kernels"].get(names[dom], {}).get("dram_bytes_per_launch")
                traffic_note = "ncu --set full capture of these sources (profiles/r2_traffic.json)"
        except Exception:
            pass
        # exact FLOP count of the Riccati sweep per node against the measured FP64 tensor rate
        flops_bw = 2 * (2 * ndx**3 + 2 * ndx**2 * nu + ndx * nu**2 + ndx**2 * nu) + nu**3 / 3 + 2 * nu**2 * (ndx + 1) + 2 * ndx**2
        fp64_peak = 37.05  # TFLOP/s, mma.sync m8n8k4.f64 measured on this pool (profiles/r1_baseline.md)
        fp64_ach = flops_bw * float(units_k[1]) * T / (ms_k[1] * 1e-3) / 1e12
        ceiling = peak * 1e9 / (T * sum(bytes_node.values()))  # HBM-bound OCP-iterations/s per GPU (1 rollout trial)
        roofline = {"bound": "hbm", "kernel": kernels[names[dom]], "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_kind": peak_kind,
                    "algorithmic_bytes_per_node": bytes_node[names[dom]],
                    "algorithmic_bytes_per_launch": bytes_node[names[dom]] * float(units_k[dom]) * T / n_launch,
                    "avg_launch_ms": float(ms_k[dom] / n_launch),
                    "share_of_step": {n: float(ms_k[i] / ms_k.sum()) for i, n in enumerate(names)},
                    "ms_by_kernel_per_step": {n: float(ms_k[i]) for i, n in enumerate(names)},
                    "hbm_frac_by_kernel": {n: float(bytes_node[n] * float(units_k[i]) * T / (ms_k[i] * 1e-3) / 1e9 / peak)
                                           for i, n in enumerate(names[:3])},
                    "backward_fp64": {"achieved_tflops": fp64_ach, "peak_tflops": fp64_peak, "frac": fp64_ach / fp64_peak,
                                      "flops_per_node": flops_bw, "peak_kind": "measured DMMA (scripts/microbench/fp64_peak.cu)"},
                    "step_hbm_ceiling_ocp_iter_per_s": ceiling, "step_frac_of_hbm_ceiling": (iters_all / world) / wall / ceiling,
                    "note": "per-kernel times: CUDA events on the solver's stream around every launch of one extra instrumented "
                            "step (same schedule as the timed steps)"}
        out = {"metric": METRIC, "value": iters_all / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": args.workload, "T": T, "dt_ms": dt, "batch_per_gpu": B, "maxiter": 100,
                          "nx": nx, "ndx": ndx, "nu": nu, "x0": f"YAML initial state + 0.05*U(-1,1), seeds {seed0}+b",
                          "l2": "inputs_larger_than_L2 (node tiles: %.1f GB per GPU)" % (B * (T + 1) * fp.tile * 8 / 1e9),
                          "iterations_per_step": iters_all / args.steps, "device_ms_per_step": dev_ms / args.steps},
               "clocks": clocks,
               "e2e": {"value": iters_e2e_all / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": 1e3 * wall_e2e / args.steps,
                       "outputs": "xs, us, cost, iters (K and k stay on the device: empc_get_K / empc_get_k on request)"},
               "gpu_launches": int(launches_all), "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            ob = oracle_binding()
            cores = os.cpu_count() or 1
            n = args.cpu_sample or min(1024, 64 * cores)  # 10-20 s of CPU work
            sec, it, _c = ob.solve_batch(fp, x0[:n], cores)
            out["cpu_baseline"] = {"value": float(it.sum() / sec), "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"first {n} OCPs of the same batch, {cores} host threads, {sec:.1f} s; CPU restatement (oracle/), not Crocoddyl itself"}
            # also the way the reference itself runs -- one OCP at a time on one thread
            n1 = max(8, min(48, n))
            sec1, it1, _c1 = ob.solve_batch(fp, x0[:n1], 1)
            out["cpu_baseline"]["single_thread"] = {"value": float(it1.sum() / sec1), "unit": UNIT, "cores": 1,
                                                    "sample": f"first {n1} OCPs, one at a time on one thread, {sec1:.1f} s"}
        if world == 1 and not args.no_mpc:
            solver.close()
            out["mpc_step_latency"] = mpc_latency(args.mpc_steps, not args.no_cpu_baseline)
        if world == 1 and not args.no_config5:
            solver.close()
            out["config5"] = config5_leg(args)
        if world == 1 and not args.no_divergent:
            solver.close()
            out["divergent_batch"] = divergent_leg(args)
    solver.close()
    if not args.no_config4:
        c4 = config4_leg(args, torch, dist, world, rank, local_rank, barrier)
        if rank == 0:
            out["config4"] = c4
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
