"""The measurement scripts at least parse their arguments (no GPU): a syntax or import error in a script would
otherwise only show on a GPU box, where a call costs minutes. gather_locality.py is host-only and runs for real."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = ["scale_ab.py", "solver_sweep.py", "apply_sweep.py", "config3_gmres.py", "config4_projection.py",
           "time_to_solution.py", "gather_locality.py", "playground_cahn_hilliard.py"]   # generic_probe.py and order_sweep.py take a bare number


@pytest.mark.parametrize("script", SCRIPTS)
def test_script_prints_its_usage(script):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), "--help"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "usage" in res.stdout.lower()


def test_gather_locality_counts_on_a_small_mesh():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gather_locality.py"), "--axis", "14", "--parts", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "metis rank 1" in res.stdout and "slab rank 0" in res.stdout and "sectors per 64-row slice" in res.stdout


def test_gpu_session_script_is_valid_bash():
    assert subprocess.run(["bash", "-n", os.path.join(ROOT, "scripts", "gpu_session_r02.sh")]).returncode == 0
