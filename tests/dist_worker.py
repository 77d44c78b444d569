"""Run under torchrun (gloo, world_size 2) by tests/test_dist_gloo.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proj_entropy_b200.dist import Ranks, shard_bounds

r = Ranks("gloo")
payload = bytes(range(128)) if r.rank == 0 else None
got = r.broadcast_bytes(payload, 128)
out = {
    "rank": r.rank, "world": r.world,
    "id_ok": got == bytes(range(128)),
    "max": r.max(10.0 + r.rank), "sum": r.sum(1.0 + r.rank),
    "shard": shard_bounds(1000, r.world, r.rank),
}
r.barrier()
with open(os.path.join(sys.argv[1], "rank%d.json" % r.rank), "w") as f:
    json.dump(out, f)
r.close()
