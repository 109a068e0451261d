"""GPU test of the N > 1 path: spawns one process per GPU with torchrun and checks parity of the
distributed solve (halo exchange, allreduce, sliced hierarchy + replicated tail) against the
oracle.  Needs at least two GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mp_gpu_check.py")] + args
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("kind,dims,rep,share", [
    ("lap7", ("12", "11", "6"), "40", "0"), ("lap7", ("16", "16", "10"), "262144", "0"),
    ("lap27", ("8", "8", "6"), "30", "0"), ("convdif", ("16", "8", "6"), "40", "0"),
    # replicated setup only (no work sharing between the ranks)
    ("lap7", ("12", "11", "6"), "40", "off"),
    # NCCL send/recv halo exchange instead of the peer-memory (CUDA IPC) path
    ("lap7", ("12", "11", "6"), "40", "nccl"),
    # sliced-ELL kernel on every level: off-diagonal block fused into the SpMV kernel (in-kernel
    # wait on the neighbours' flags), and the same layout with the separate correction kernel
    ("lap7", ("16", "16", "10"), "40", "sell-fused"), ("lap27", ("8", "8", "6"), "30", "sell-fused"),
    ("convdif", ("16", "8", "6"), "40", "sell-fused"), ("lap7", ("16", "16", "10"), "40", "sell-unfused"),
    ("lap7", ("12", "11", "6"), "40", "sell-nccl"),
    # uneven row slabs that cut through grid planes (asymmetric halo sizes), all three exchange paths
    ("lap7", ("14", "9", "7"), "40", "ragged"), ("lap27", ("8", "8", "6"), "30", "ragged-sell"),
    ("convdif", ("16", "8", "6"), "40", "ragged-nccl")])
def test_two_rank_solve_matches_oracle(gpu, kind, dims, rep, share):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = {"HDK_REPLICATE_ROWS": rep}
    if share == "off":
        env["HDK_SETUP_SHARE"] = "0"
    elif share == "nccl":
        env["HDK_HALO_IPC"] = "0"
    elif share.startswith("ragged"):
        env["MPCHECK_RAGGED"] = "1"
        env["HDK_SHARE_MIN_ROWS"] = "0"
        if share == "ragged-sell":
            env["HDK_SELL_MIN_ROWS"] = "0"
            env["HDK_SELL_MIN_ROWS_DIST"] = "0"
        if share == "ragged-nccl":
            env["HDK_HALO_IPC"] = "0"
    elif share.startswith("sell"):
        env["HDK_SELL_MIN_ROWS"] = "0"
        env["HDK_SELL_MIN_ROWS_DIST"] = "0"
        if share == "sell-unfused":
            env["HDK_FUSE_OFFD"] = "0"
        if share == "sell-nccl":
            env["HDK_HALO_IPC"] = "0"
    else:
        env["HDK_SHARE_MIN_ROWS"] = share  # share the interpolation / RAP rows even on these tiny levels
    r = _run(2, [kind, *dims], env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout
