"""World-size > 1 tests. The ranks run tests/_dist_worker.py under torch.distributed.run:
on CPU with gloo (host-side logic of the N>1 path: partition broadcast, local meshes, halo maps, rank-ordered
reductions; runs in this container) and on >= 2 GPUs (the distributed apply / CG / BiCGStab through the C ABI,
bit-exact against the oracle in the ranks' concatenated order with the segmented reduction tree)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_dist_worker.py")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(world, mode, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert res.returncode == 0, f"worker failed\n--- stdout\n{res.stdout[-4000:]}\n--- stderr\n{res.stderr[-6000:]}"
    for r in range(world):
        assert f"rank {r}: OK" in res.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_host_side_of_the_distributed_path_gloo(world):
    launch(world, "cpu", timeout=600)


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
def test_distributed_solvers_bit_exact_two_ranks_sharing_one_gpu():
    """World size 2 on whatever is there -- on a one-GPU box both ranks map to device 0 (multigpu.local_device): their
    kernels are time-sliced, every peer wait costs a time slice instead of an NVLink round trip, but the protocol
    (CUDA-IPC peer mapping, halo pushes, sequence flags, mailbox all-reduce, folded reductions, the persistent kernel's
    grid barriers) is exactly the one that runs with a GPU per rank, and the results are bit-exact against the oracle."""
    env_backup = os.environ.get("SB_SPIN_TIMEOUT_S")
    os.environ["SB_SPIN_TIMEOUT_S"] = "60"
    try:
        launch(2, "p2p", timeout=900)
    finally:
        if env_backup is None:
            os.environ.pop("SB_SPIN_TIMEOUT_S", None)
        else:
            os.environ["SB_SPIN_TIMEOUT_S"] = env_backup


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_distributed_solvers_bit_exact_two_gpus(mode):
    if _gpu_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    launch(2, mode, timeout=900)


@pytest.mark.gpu
def test_distributed_solvers_bit_exact_all_gpus():
    n = _gpu_count()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    launch(min(n, 8), "p2p", timeout=900)
