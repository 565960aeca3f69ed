"""bench.py's contract, the parts that run without a GPU: the reference arm (`--impl reference`, the reference's own CPU
solver on the same workload) prints ONE JSON line with the keys the driver reads, and the all-cores CPU port used for
`cpu_baseline_all_cores` agrees with the single-threaded reference. (Small mesh: the arm's code path, not its speed.)"""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


@pytest.mark.parametrize("solver", ["bicgstab", "cg"])
def test_reference_arm_prints_the_contract_line(solver):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "14", "--steps", "3",
                          "--warmup", "2", "--solver", solver], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "krylov_iterations_per_sec" and line["unit"] == "it/s"
    assert line["steps"] == 3 and line["warmup"] == 2 and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1000.0) < 1e-6 * 1000.0
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert line["config"]["solver"] == solver and line["config"]["cells"] == 6 * 14 ** 3
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("solver", ["bicgstab", "cg"])
def test_all_cores_port_agrees_with_the_reference(solver):
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_omp.so")):
        pytest.skip("oracle/liboracle_omp.so not built (no OpenMP-capable compiler)")
    bench = _bench_module()
    mesh, x_star = bench.build_problem(types.SimpleNamespace(cell="tet", n=16))
    got = bench.cpu_all_cores_rate(mesh, x_star, solver, budget_s=0.2)
    assert "unavailable" not in got, got
    assert got["kind"] == "port" and got["cores"] >= 1 and got["value"] > 0 and np.isfinite(got["value"])


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: the driver launches the arm under torchrun like the product arm; rank 0 alone runs and prints, the other
    ranks exit 0 without work."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--axis", "12", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2 and json.loads(lines[0])["impl"] == "reference"


def test_reference_arm_never_maps_the_cuda_library():
    """The process that times the reference's CPU solver prepares its inputs through oracle/libsb_meshprep.so (the
    product's host-only mesh sources, no device code) and maps nothing of the product's CUDA library."""
    code = (
        "import runpy, sys\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--axis', '12', '--steps', '2', '--warmup', '1']\n"
        "runpy.run_path('bench.py', run_name='__main__')\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('MAPS', 'libstormb200.so' in maps, 'libsb_meshprep.so' in maps, 'libref_solvers.so' in maps or 'liboracle.so' in maps)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAPS False True True" in out.stdout, out.stdout[-500:]


def test_committed_bench_line_has_the_contract_keys():
    """The round's bench line as measured on the B200 (profiles/r02_bench_n1.json): every key the measurement contract
    names, with the meaning it names (roofline per launch of the dominant kernel, CPU baseline beside it, end-to-end
    number with its copy volumes, clock record, parity of both solvers at full size)."""
    import json
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["metric"] == "krylov_iterations_per_sec" and line["unit"] == "it/s" and line["dtype"] == "f64"
    assert line["vs_baseline"] is None and line["higher_is_better"] is True and "workload" in line["config"]
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    # per variant: its own bytes (the <r~,v> epilogue reads a third vector), never more than the measured copy bandwidth
    n = line["config"]["cells"]
    uy, yy = r["variants"]
    assert uy["algorithmic_bytes_per_launch"] - yy["algorithmic_bytes_per_launch"] == 8 * n
    assert all(0.5 < v["frac"] < 1.0 and 0.0 < v["final_stage_ms"] < v["avg_launch_ms"] for v in r["variants"])
    assert 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1          # no wasted re-reads
    c = line["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] == 1 and c["unit"] == "it/s" and c["value"] > 0 and c["sample"]
    e = line["e2e"]
    assert e["unit"] == "it/s" and 0 < e["value"] < line["value"] and e["h2d_bytes_per_step"] == 16 * n / line["steps"]
    assert line["gpu_launches"] >= 8 * line["steps"]                             # 5 steps + 3 final stages per iteration
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    # parity at full size: BiCGStab leaves the 1e-10 bar early (reduction order), CG stays inside it
    assert line["parity"]["measured_path_coef_rows"]["first_iteration_over_1e-10"] >= 5
    assert line["parity_other_solver"]["measured_path_coef_rows"]["first_iteration_over_1e-10"] is None
