"""One process per GPU (the torchrun model bench.py runs under): rank-mode
handles, NCCL id broadcast, CUDA-IPC peer connection, gathers on download."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import proj_entropy_b200 as g
from conftest import ROOT
from oracle import oracle_py as oracle

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_rank_mode_single_rank_matches_plain_handle():
    n = 3000
    m, x, y, vx, vy = g.plummer(n, seed=4)
    with g.GravityCuda(n, rank=0, nranks=1, device=0, kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx, vy)
        h.step(2, 3)
        a = h.download()
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx, vy)
        h.step(2, 3)
        b = h.download()
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
    with pytest.raises(g.GravityError, match="rank"):
        g.GravityCuda(n, rank=3, nranks=2, device=0)


def run_ranks(tmp_path, nproc, exchange, n, mode="parity"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "rank_worker.py"), str(tmp_path), str(exchange), str(n), mode]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    return [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(nproc)]


@pytest.mark.multigpu
@pytest.mark.parametrize("exchange", [0, 1])
def test_two_ranks_under_torchrun(gpu_count, exchange, tmp_path):
    """The mode bench.py scales in, checked against the ORACLE on rank 0: accelerations after two
    steps <= 1e-12, the state after five steps, the energy; every rank downloads identical bits;
    shard-only downloads; a re-upload without masses."""
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    n = 20000
    res = run_ranks(tmp_path, 2, exchange, n)
    assert all(r["world"] == 2 and r["same_after_reupload"] and r["shard_download_ok"] and r["step_host_ok"] for r in res)
    assert res[0]["accel_err"] <= 1e-12
    assert res[0]["pos_err"] <= 1e-12 and res[0]["vel_err"] <= 1e-11 and res[0]["energy_err"] <= 1e-12
    # every rank downloads the same full state
    for name in ("x", "vy"):
        a = np.load(tmp_path / ("%s_rank0.npy" % name))
        b = np.load(tmp_path / ("%s_rank1.npy" % name))
        assert np.array_equal(a, b)


@pytest.mark.multigpu
def test_a_peer_that_never_steps_times_out_and_kills_the_handle(gpu_count, tmp_path):
    """spin_timeout_ms: the waiting rank gets an error after the limit (not a hang), nothing stale
    is committed or published afterwards, and the handle refuses further work (fail-stop)."""
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    res = run_ranks(tmp_path, 2, 1, 20000, mode="timeout")
    r0 = res[0]
    assert r0["error"] and "timed out" in r0["error"], r0
    assert r0["seconds"] < 5.0
    assert r0["second_error"] and "dead" in r0["second_error"], r0
