"""CPU-side tests: host-side mirror (YAML/URDF readers, factories, flattening, reference quirks), the C ABI library
(loads, exports every symbol include/empc_b200.h declares — no compute calls without a GPU), oracle goldens, an LQR
closed form for the Riccati recursion and the multi-rank sharding helpers over gloo."""
import ctypes as C
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle_binding as ob
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
host = importlib.import_module("eagle-mpc_b200.host")
capi = importlib.import_module("eagle-mpc_b200.capi")
abi = importlib.import_module("eagle-mpc_b200.abi")
wl = importlib.import_module("eagle-mpc_b200.workloads")
sharding = importlib.import_module("eagle-mpc_b200.sharding")


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "empc_b200.h")).read()
    names = sorted(set(re.findall(r"\b(empc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 30
    lib = capi.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    p = capi.default_params()
    assert p.maxiter == 100 and p.convergence_init == 1e-2 and p.th_acceptnegstep == 2 and p.th_stop_gaps == 1.0
    assert C.sizeof(abi.Cost) == 40 and C.sizeof(abi.SolverParams) == 24 + 21 * 8 + 8
    b = capi.box_params(abi.SOLVER_BOXDDP)   # crocoddyl::SolverBoxDDP: th_stop_ = 5e-5, BoxQP(nu, 100, 0.1, 1e-5, 0), upstream stop rule
    assert b.solver_type == abi.SOLVER_BOXDDP and b.th_stop == 5e-5 and b.boxqp_maxiter == 100 and b.boxqp_th_grad == 1e-5
    assert b.stop_criteria == abi.STOP_CRITERIA_QU_NORM and b.stop_test == abi.STOP_TEST_FEASIBLE and p.solver_type == abi.SOLVER_SBFDDP


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product path must fail, not fall back (run only where no GPU is visible)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is visible")
    except ImportError:
        pass
    h = synth.make_problem(seed=0, na=3, T=5)
    with pytest.raises(capi.EmpcError):
        capi.BatchSolver(h, 2)


def test_parser_yaml_key_layout_and_quirks():
    p = host.parse_yaml("hexacopter370_flying_arm_3/trajectories/displacement.yaml")
    assert p["robot/urdf"] == "hexacopter370_description/urdf/hexacopter370_flying_arm_3.urdf"
    assert p["robot/platform/n_rotors"] == "6" and p["robot/platform/cf"] == "4.138394792004922e-06"
    assert p["robot/platform/rotors"].startswith("[{translation:[0.1602147,0.0925,0.0]")
    assert p["stages/nav_wp1/transition"] == "true" and p["stages/wp_1/transition"] == "false"  # presence, not value
    assert p["stages/nav_wp1/costs/limits_state/l_bound"] == "[0,0,0,0,0,0,-1.5,-1.5,-1.5,0,0,0,0,0,0,-3,-3,-3]"
    assert p["stages/wp_4/costs/placement_gripper/link_name"] == "flying_arm_3__gripper"
    m = host.parse_yaml("hexacopter370_flying_arm_3/mpc/mpc.yaml")
    assert m["mpc_controller/knots"] == "30" and m["mpc_controller/carrot_state_limits_u_bound"].startswith("[0,0,0")
    # trailing comma inside a vector literal is tolerated (yaml/hexacopter370/trajectories/displacement.yaml)
    d = host.parse_yaml("hexacopter370/trajectories/displacement.yaml")
    assert d["problem_params/dt"] == "10"


def _write_yaml(text):
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    f.write(text)
    f.close()
    return f.name


BASE = """trajectory:
  robot:
    name: "hexacopter370"
    urdf: "hexacopter370_description/urdf/hexacopter370.urdf"
    follow: "hexacopter370/platform/hexacopter370.yaml"
  stages:
    - name: "a"
      duration: 100
      costs:
        - name: "reg"
          type: "CostModelState"
          weight: 1
          %s
"""


def test_factory_errors_and_quirks():
    # scientific notation inside a vector literal is rejected (src/utils/converter_utils.cpp:39-40)
    # ... for a state reference the factory swallows the exception and falls back to state.zero() (src/factory/cost.cpp:41-47)
    y = _write_yaml(BASE % "reference: [0, 0, 0, 0, 0, 0, 1, 1e-1, 0, 0, 0, 0, 0]")
    fp = host.Trajectory(y).createProblem(20, add_barrier=False)
    ref = np.ctypeslib.as_array(fp.desc.pool, shape=(fp.desc.n_pool,))[fp.desc.costs[0].ref_off:][:13]
    assert np.array_equal(ref, [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0])
    # ... but not for a frame position (:81), where it propagates
    y = _write_yaml((BASE % "reference: [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]").replace(
        'type: "CostModelState"', 'type: "CostModelFrameTranslation"\n          link_name: "hexacopter370__base_link"\n          position: [0, 0, 1e-1]'))
    with pytest.raises(capi.EmpcError, match="Invalid string representation of a Matrix"):
        host.Trajectory(y)
    # wrong reference dimension
    y = _write_yaml(BASE % "reference: [0, 0, 0]")
    with pytest.raises(capi.EmpcError, match="State reference vector"):
        host.Trajectory(y)
    # unknown link
    y = _write_yaml((BASE % "reference: [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]").replace(
        'type: "CostModelState"', 'type: "CostModelFrameTranslation"\n          link_name: "nope"\n          position: [0, 0, 1]'))
    with pytest.raises(capi.EmpcError, match="does no exists"):
        host.Trajectory(y)
    # "active" key present (whatever its value) => the cost is added inactive (src/stage.cpp:55-61)
    y = _write_yaml(BASE % "reference: [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]\n          active: 1")
    fp = host.Trajectory(y).createProblem(20, add_barrier=False)
    assert fp.desc.costs[0].active == 0
    # two consecutive zero-duration stages are rejected
    two = BASE % "reference: [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]"
    two = two.replace("duration: 100", "duration: 0") + two[two.index("    - name"):].replace('"a"', '"b"').replace("duration: 100", "duration: 0")
    with pytest.raises(capi.EmpcError, match="Two consecutives stages"):
        host.Trajectory(_write_yaml(two))
    with pytest.raises(capi.EmpcError, match="trajectory or an mpc_controller"):
        host.Trajectory(_write_yaml("foo:\n  bar: 1\n"))


@pytest.mark.parametrize("name,T,nsets", [("hexacopter370_hover", 100, 2), ("hexacopter370_passthrough", 245, 6),
                                          ("hexacopter370_flying_arm_3_displacement", 400, 8),
                                          ("hextilt_flying_arm_5_push_slide", 100, 1), ("iris_px4_hover", 250, 2)])
def test_knot_layout_and_flattening(name, T, nsets):
    yaml, dt, _ = wl.CONFIGS[name]
    tr = host.Trajectory(yaml)
    fp = tr.createProblem(dt)
    assert fp.T == T and fp.desc.n_costsets == nsets  # src/trajectory.cpp:117-127 knot rule, one model per stage
    nc = np.ctypeslib.as_array(fp.desc.node_costset, shape=(T + 1,))
    assert nc[T] == nsets - 1  # terminal model = last stage's model
    for s in range(nsets):
        names = fp.cost_names(s)
        assert names == sorted(names)  # crocoddyl iterates a std::map
        running = s in nc[:T]
        assert ("barrier" in names) == running  # SolverSbFDDP::barrierInit touches running models only
    # platform: tau_f third row is all ones for flat rotors, bounds follow the platform YAML and effort limits
    if "hextilt" not in name:
        assert np.allclose(tr.tau_f[2], 1.0) and np.allclose(tr.tau_f[:2], 0.0)
    assert np.all(tr.u_ub[:tr.n_rotors] > tr.u_lb[:tr.n_rotors])
    assert np.allclose(tr.u_lb[tr.n_rotors:], -tr.u_ub[tr.n_rotors:])


def test_urdf_reader_merges_fixed_links_and_frames():
    tr = host.Trajectory("hexacopter370_flying_arm_3/trajectories/displacement.yaml")
    assert (tr.nq, tr.nv, tr.nu) == (10, 9, 9)
    fp = tr.createProblem(20)
    r = fp.desc.robot
    assert r.n_joints == 4 and list(r.parent[:4]) == [-1, 0, 1, 2]
    # the gripper link is fixed to the last arm link: its mass is merged there and its frame hangs on joint 3
    assert abs(r.mass[3] - (0.08 + 0.03)) < 1e-12
    frames = {r.frame_joint[f] for f in range(r.n_frames)}
    assert frames == {0, 3}


def test_oracle_matches_golden_fixtures():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_named_problems.json")))
    for name in ("hexacopter370_flying_arm_3_displacement", "hextilt_flying_arm_5_push_slide", "hexacopter370_passthrough"):
        yaml, dt, _ = wl.CONFIGS[name]
        fp = host.Trajectory(yaml).createProblem(dt)
        for rec in gold[name]["ocps"][:2]:
            o = ob.Oracle(fp)
            o.set_x0(np.array(rec["x0"]))
            o.solve()
            assert int(o.get("iter")) == rec["iter"] and int(o.get("feasible")) == rec["feasible"]
            assert abs(o.get("cost") - rec["cost"]) <= 1e-9 * abs(rec["cost"])
            xs, us = o.get("xs"), o.get("us")
            assert np.allclose(xs[-1], rec["xs_T"], rtol=0, atol=1e-8)
            for t, u in rec["us_samples"].items():
                assert np.allclose(us[int(t)], u, rtol=0, atol=1e-8)


def test_riccati_matches_lqr_closed_form():
    """Backward pass of the oracle on its own linearisation == textbook discrete Riccati recursion (dense numpy)."""
    h = synth.make_problem(seed=4, na=3, T=8, all_costs=False)
    o = ob.Oracle(h)
    rng = np.random.default_rng(1)
    xs = np.stack([synth.random_state(rng, h, 0.2) for _ in range(h.T + 1)])
    us = rng.uniform(2, 8, size=(h.T, h.nu)); us[:, 6:] = rng.uniform(-1, 1, size=(h.T, 3))
    o.set_x0(xs[0]); o.set_candidate(xs, us, False)
    o.phase_calc_diff(0.1)
    assert o.phase_backward(1e-8, False) == 1
    tiles, fs = o.get("tiles"), o.get("fs")
    off = h.tile_offsets(); ndx, nu = h.ndx, h.nu
    blk = lambda t, n, r, c: tiles[t, off[n]:off[n] + r * c].reshape(r, c)
    Vxx = blk(h.T, "Lxx", ndx, ndx) + 1e-8 * np.eye(ndx)
    Vx = tiles[h.T, off["Lx"]:off["Lx"] + ndx] + Vxx @ fs[h.T]
    K, k = o.get("K"), o.get("k")
    for t in range(h.T - 1, -1, -1):
        Fx, Fu = blk(t, "Fx", ndx, ndx), blk(t, "Fu", ndx, nu)
        Qxx = blk(t, "Lxx", ndx, ndx) + Fx.T @ Vxx @ Fx
        Qxu = blk(t, "Lxu", ndx, nu) + Fx.T @ Vxx @ Fu
        Quu = blk(t, "Luu", nu, nu) + Fu.T @ Vxx @ Fu + 1e-8 * np.eye(nu)
        Qx = tiles[t, off["Lx"]:off["Lx"] + ndx] + Fx.T @ Vx
        Qu = tiles[t, off["Lu"]:off["Lu"] + nu] + Fu.T @ Vx
        Kt = np.linalg.solve(Quu, Qxu.T); kt = np.linalg.solve(Quu, Qu)
        assert np.allclose(K[t], Kt, rtol=1e-7, atol=1e-9 * np.abs(Kt).max())
        assert np.allclose(k[t], kt, rtol=1e-7, atol=1e-9 * max(1, np.abs(kt).max()))
        Vx = Qx - Kt.T @ Qu
        Vxx = Qxx - Qxu @ Kt
        Vxx = 0.5 * (Vxx + Vxx.T) + 1e-8 * np.eye(ndx)
        Vx = Vx + Vxx @ fs[t]
    assert np.allclose(o.get("Vx")[0], Vx, rtol=1e-6, atol=1e-8 * np.abs(Vx).max())


def test_shard_ranges_cover_the_batch():
    for total in (1, 7, 4096, 16384):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
sharding = importlib.import_module("eagle-mpc_b200.sharding")
wl = importlib.import_module("eagle-mpc_b200.workloads")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total, nx = 11, 19
x0 = wl.noisy_x0(np.eye(1, nx, 6)[0], total, 2024) if rank == 0 else None
mine = sharding.scatter_rows(x0, total, (nx,), dist)
b, e = sharding.shard_range(total, rank, world)
ref = wl.noisy_x0(np.eye(1, nx, 6)[0], e - b, 2024, first=b)   # each rank can also regenerate its slice from the seeds
assert mine.shape == (e - b, nx) and np.array_equal(mine.numpy(), ref), "scatter mismatch"
out = sharding.gather_rows(mine * 2.0, total, dist)          # stand-in for "solve": no collective in between
if rank == 0:
    assert np.array_equal(out.numpy(), 2.0 * x0), "gather mismatch"
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)                     # the bench's max-over-ranks timing reduction
assert t.item() == world
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_scatter_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29431", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2
