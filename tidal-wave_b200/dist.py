"""Multi-GPU plumbing: one process per GPU, launched by torch.distributed.run.

The path shards by independent units (image pairs): there is NO data-path collective.  The only
cross-rank traffic is (i) the barrier around the timed region, (ii) a MAX over ranks of the timed duration and
(iii) a SUM of the Report counters (/root/reference/src/message_queue.h:44-48) at the end.  Backend: gloo (CPU tensors).
"""
from __future__ import annotations

import os


class Dist:
    def __init__(self, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.pg = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # The path has no data-path collective (independent pairs), so the process group only carries a barrier and
            # three scalars: gloo on CPU tensors does that without touching the GPUs (and without NCCL's stdout banner,
            # which would break the one-JSON-line contract of bench.py).  backend="nccl" remains selectable.
            if backend is None:
                backend = "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
                self.device = "cuda"
            else:
                dist.init_process_group(backend)
            self.pg = dist

    def barrier(self):
        if self.pg is not None:
            self.pg.barrier()

    def _reduce(self, values, op):
        if self.pg is None:
            return list(values)
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device=self.device)
        self.pg.all_reduce(t, op=op)
        return [float(v) for v in t.tolist()]

    def reduce_max(self, value: float) -> float:
        if self.pg is None:
            return float(value)
        return self._reduce([value], self.pg.ReduceOp.MAX)[0]

    def reduce_sum(self, values):
        if self.pg is None:
            return list(values)
        return self._reduce(values, self.pg.ReduceOp.SUM)

    def close(self):
        if self.pg is not None:
            self.pg.barrier()
            self.pg.destroy_process_group()
            self.pg = None


def shard(n_units: int, rank: int, world: int) -> range:
    """Static contiguous partition of n independent units (pairs) over ranks; every unit belongs to exactly one rank."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def whole_job_throughput(units_per_rank: int, world: int, max_elapsed_s: float) -> float:
    """value = the units all ranks processed / the slowest rank's time (bench contract)."""
    return units_per_rank * world / max_elapsed_s
