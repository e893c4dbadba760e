"""Multi-GPU shard equivalence on real GPUs: G ranks under torchrun, instances sharded by
contiguous range; the NCCL-gathered trajectory, the trajectory the step kernels store straight into every rank's
symmetric-memory buffer (NVLS multimem.st or NVLink peer stores), and a single-GPU run over all instances must be
BITWISE equal; the all-reduced rollout cost vector must match the single-GPU one.  Skips with fewer than two GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("multicast", [True, False])
def test_sharded_gather_and_cost_allreduce_equal_single_gpu(built_lib, multicast):
    g = _gpu_count()
    if g < 2:
        pytest.skip(f"needs >= 2 GPUs, {g} visible")
    world = 8 if g >= 8 else (4 if g >= 4 else 2)
    env = dict(os.environ)
    env.pop("CDPR_NO_MULTICAST", None)
    if not multicast:
        env["CDPR_NO_MULTICAST"] = "1"
    port = 29500 + (os.getpid() % 400) + (0 if multicast else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[config 4]" in r.stdout and "bitwise equal to 1-GPU run: True" in r.stdout
    assert "[config 5]" in r.stdout
    assert r.stdout.count("== NCCL all-gather result: True") == world or "symmetric memory unavailable" in r.stdout


def test_cpp_host_drives_configs_4_and_5_without_python(built_lib):
    """The same two exchanges from a C++ host through the C ABI alone (cdpr_comm_*: one process, peer memory over NVLink, the
    library's own kernels): gathered trajectory bitwise equal to the 1-GPU run on every device, rank-ordered cost all-reduce
    identical on every device."""
    g = _gpu_count()
    if g < 2:
        pytest.skip(f"needs >= 2 GPUs, {g} visible")
    from cdpr_simulation_b200 import build as b
    b.build_host()
    r = subprocess.run([b.HOST_MULTI_BIN, str(8 if g >= 8 else (4 if g >= 4 else 2))], capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0 and "MULTI-GPU C++ HOST CHECK PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
