"""The N > 1 host plumbing on CPU: torchrun, world_size 2, gloo backend."""
import json
import os
import socket
import subprocess
import sys

from conftest import ROOT
from proj_entropy_b200.dist import shard_bounds


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torchrun(args, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port())] + args
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)


def test_rank_plumbing_world_size_2(tmp_path):
    p = torchrun([os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)])
    assert p.returncode == 0, p.stderr[-2000:]
    outs = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(2)]
    for r, o in enumerate(outs):
        assert o["rank"] == r and o["world"] == 2
        assert o["id_ok"]                       # the 128-byte NCCL id reaches every rank intact
        assert o["max"] == 11.0 and o["sum"] == 3.0
    assert outs[0]["shard"] == [0, 512] and outs[1]["shard"] == [512, 1000]


def test_shard_geometry_covers_every_body_once():
    for n in (1, 7, 255, 256, 257, 1000, 65536, 1 << 20, (1 << 20) + 3):
        for p in (1, 2, 4, 8):
            cover = []
            for r in range(p):
                b, e = shard_bounds(n, p, r)
                assert 0 <= b <= e <= n
                assert b % 256 == 0 or b == n      # shards start on a j-tile boundary
                cover.append((b, e))
            assert cover[0][0] == 0 and max(e for _, e in cover) == n
            assert all(cover[k][1] == cover[k + 1][0] or cover[k + 1][0] == n for k in range(p - 1))
            assert sum(e - b for b, e in cover) == n


def test_reference_arm_under_torchrun_prints_one_line():
    p = torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                  "--bodies", "4096", "--cpu-rows", "512"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                      # rank 0 alone runs and prints
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "pairs/s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["e2e"]["h2d_bytes_per_step"] == 0
