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
            kwargs = {}
            if backend == "nccl":
                torch.cuda.set_device(local)
                kwargs["device_id"] = torch.device(f"cuda:{local}")  # no rank -> GPU guessing at the first barrier
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def gpu_numa_node(device: int):
    """NUMA node of a GPU from sysfs (None when unknown)."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fp:
            node = int(fp.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device: int):
    """Pin this process to the CPUs of the GPU's NUMA node so that the pinned staging buffers it
    allocates afterwards are node-local (the reference leaves placement to the OS; with one
    process per GPU, cross-socket pinned memory halves the PCIe throughput).  Returns the node or
    None if nothing was changed."""
    node = gpu_numa_node(device)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fp:
            spec = fp.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


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
