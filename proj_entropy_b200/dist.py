"""Rank plumbing for one-process-per-GPU runs (torchrun): rendezvous, the
broadcast of the library's NCCL id, barriers and max/sum over ranks.

torch.distributed is plumbing only -- the data path (the per-step all-gather of
positions) runs on libgravity_cuda's own NCCL communicator.  The helpers work
on the "gloo" backend with CPU tensors too, which is how the CPU test-suite
covers them (tests/test_dist_gloo.py)."""
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(n, nranks, rank, tile=256):
    """[begin, end) of the i-bodies rank `rank` owns, exactly as libgravity_cuda
    shards them (set_geometry in csrc/gravity_cuda.cu): contiguous blocks of
    chunk = roundup(ceil(n / nranks), tile) bodies; trailing ranks may be short
    or empty."""
    per = (n + nranks - 1) // nranks
    chunk = max(tile, ((per + tile - 1) // tile) * tile)
    begin = min(n, rank * chunk)
    return begin, min(n, begin + chunk)


class Ranks:
    def __init__(self, backend=None):
        self.rank, self.world, self.local_rank = env_rank()
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            backend = backend or "nccl"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = "cuda"
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, v, op):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._reduce(v, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, v):
        return self._reduce(v, self.dist.ReduceOp.SUM if self.dist else None)

    def broadcast_bytes(self, payload, nbytes, src=0):
        """Rank `src` passes `payload` (bytes of length nbytes); everyone gets it."""
        if self.dist is None:
            return payload
        import torch
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        if self.rank == src:
            buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        self.dist.broadcast(buf, src)
        return bytes(buf.cpu().numpy().tobytes())

    def all_gather_bytes(self, payload):
        """Every rank passes `payload` (same length everywhere); returns the
        concatenation in rank order."""
        if self.dist is None:
            return bytes(payload)
        import torch
        mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(self.device)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine)
        return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None
