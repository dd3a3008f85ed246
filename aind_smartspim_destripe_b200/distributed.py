"""One-process-per-GPU plumbing for Z-slab sharding.

Planes are independent (reference zarr_destriper.py:319-327), so the data path has NO
collective: ``torch.distributed`` is used only for the start/stop barrier and for reducing
timings / plane counts (NCCL on the GPU box, gloo in CPU tests).
"""

from __future__ import annotations

import os
from typing import Tuple


def env_rank() -> Tuple[int, int, int]:
    return (
        int(os.environ.get("RANK", "0")),
        int(os.environ.get("WORLD_SIZE", "1")),
        int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))),
    )


def init(backend: str = None) -> Tuple[int, int, int]:
    """Initialise the default process group when WORLD_SIZE > 1; returns (rank, world, local_rank)."""
    rank, world, local = env_rank()
    if world > 1:
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _device():
    import torch
    import torch.distributed as dist

    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def _reduce(value: float, op_name: str) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(value: float) -> float:
    return _reduce(value, "MAX")


def sum_over_ranks(value: float) -> float:
    return _reduce(value, "SUM")


def shutdown():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
