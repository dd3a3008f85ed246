"""CPU restatement of the reference chunk worker and of its multiprocessing scheduler.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``execute_worker``: ``/root/reference/code/aind_smartspim_destripe/zarr_destriper.py:253-336``.
  The two helpers it takes from the un-vendored ``aind-large-scale-prediction==1.0.0``
  (``recover_global_position``, ``unpad_global_coords``; pins in
  ``environment/Dockerfile:15``) are restated for the reference's only configuration,
  overlap ``(0, 0, 0)`` (``zarr_destriper.py:1018-1022``).
* ``run_planes_multiprocess``: the scheduling *shape* of ``producer``/``consumer``
  (``zarr_destriper.py:797-906``): N OS processes, each filtering whole chunks one plane
  at a time.  Used only as the timed CPU baseline.
"""

from __future__ import annotations

import multiprocessing as mp
import os
from typing import Sequence, Tuple

import numpy as np

from . import plane_filter as fl


def recover_global_position(super_chunk_slice, internal_slices):
    """Global ZYX slices = super-chunk offset + internal offset (zero-overlap case)."""
    pos = []
    for sc, inner in zip(super_chunk_slice, internal_slices):
        start = (sc.start or 0) + (inner.start or 0)
        stop = (sc.start or 0) + inner.stop
        pos.append(slice(start, stop))
    pos = tuple(pos)
    return pos, tuple(p.start for p in pos), tuple(p.stop for p in pos)


def unpad_global_coords(global_coord_pos, block_shape, overlap_prediction_chunksize, dataset_shape):
    """With overlap (0,0,0): the block's own global slices and the full local block."""
    assert all(int(o) == 0 for o in overlap_prediction_chunksize), "only zero overlap is restated"
    local = tuple(slice(0, int(n)) for n in block_shape[-3:])
    return tuple(global_coord_pos[-3:]), local


def pad_array_n_d(arr: np.ndarray, dim: int = 5) -> np.ndarray:
    """zarr_destriper.py:157-179"""
    if dim > 5:
        raise ValueError("Padding more than 5 dimensions is not supported.")
    while arr.ndim < dim:
        arr = arr[np.newaxis, ...]
    return arr


def execute_worker(
    data,
    batch_super_chunk,
    batch_internal_slice,
    cells_config,
    no_cells_config,
    overlap_prediction_chunksize,
    output_destriped_zarr,
    shadow_correction,
    dataset_name,
    logger=None,
):
    """zarr_destriper.py:253-336"""
    data = np.squeeze(data, axis=0)
    global_coord_pos, _, _ = recover_global_position(batch_super_chunk, batch_internal_slice)
    unpadded_global_slice, unpadded_local_slice = unpad_global_coords(
        global_coord_pos, data.shape, overlap_prediction_chunksize, output_destriped_zarr.shape
    )
    unpadded_local_slice = list((slice(0, 1), slice(0, 1)) + unpadded_local_slice)
    output_slices = list((slice(0, 1), slice(0, 1)) + unpadded_global_slice)
    for idx in range(output_destriped_zarr.ndim):
        if output_slices[idx].stop > output_destriped_zarr.shape[idx]:
            rest = output_slices[idx].stop - output_destriped_zarr.shape[idx]
            unpadded_local_slice[idx] = slice(
                unpadded_local_slice[idx].start, unpadded_local_slice[idx].stop - rest
            )
            output_slices[idx] = slice(output_slices[idx].start, output_destriped_zarr.shape[idx])
    output_slices = tuple(output_slices)
    unpadded_local_slice = tuple(unpadded_local_slice)

    filtered_data = np.zeros_like(data)
    input_tile_path = dataset_name.replace(".zarr", "")
    for plane_idx in range(data.shape[-3]):
        filtered_data[plane_idx, ...] = fl.filter_stripes(
            image=data[plane_idx, ...],
            input_tile_path=input_tile_path,
            no_cells_config=no_cells_config,
            cells_config=cells_config,
            shadow_correction=shadow_correction,
            microscope_high_int=2500,
        )
    filtered_data = pad_array_n_d(
        arr=filtered_data[unpadded_local_slice[2:]], dim=output_destriped_zarr.ndim
    )
    output_destriped_zarr[output_slices] = filtered_data


# --------------------------------------------------------------------------- CPU baseline
def get_cpu_limit() -> int:
    """``utils/utils.py:197-227`` semantics: CO_CPUS env -> cgroup quota -> core count."""
    co = os.environ.get("CO_CPUS")
    if co:
        return int(co)
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as fp:
            quota = int(fp.read())
        if quota > 0:
            return max(1, quota // 100000)
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as fp:
            q, p = fp.read().split()
        if q != "max":
            return max(1, int(q) // int(p))
    except (OSError, ValueError):
        pass
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_G = {}


def _init_worker(cfg):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    _G.update(cfg)


def _filter_block(block: np.ndarray) -> np.ndarray:
    out = np.empty(block.shape, dtype=np.uint16 if _G["shadow"] is not None else np.float32)
    for z in range(block.shape[0]):
        out[z] = fl.filter_stripes(
            image=block[z].astype(np.float32),
            input_tile_path="0_0",
            no_cells_config=_G["no_cells"],
            cells_config=_G["cells"],
            shadow_correction=_G["shadow"],
            microscope_high_int=2500,
        )
    return out


def run_planes_multiprocess(stack, no_cells_config, cells_config, shadow_correction, n_workers):
    """Filter a (Z,H,W) stack with ``n_workers`` processes, one contiguous Z-block each."""
    n_workers = max(1, min(int(n_workers), stack.shape[0]))
    blocks = np.array_split(stack, n_workers, axis=0)
    cfg = dict(no_cells=no_cells_config, cells=cells_config, shadow=shadow_correction)
    if n_workers == 1:
        _init_worker(cfg)
        return _filter_block(blocks[0])
    ctx = mp.get_context("fork")
    with ctx.Pool(n_workers, initializer=_init_worker, initargs=(cfg,)) as pool:
        outs = pool.map(_filter_block, blocks)
    return np.concatenate(outs, axis=0)
