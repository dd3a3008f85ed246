"""Chunk worker and plane scheduler of the reference, re-designed for one GPU per process.

* ``execute_worker`` keeps the signature of
  ``/root/reference/code/aind_smartspim_destripe/zarr_destriper.py:253-264`` but filters the
  whole ``(Z, H, W)`` chunk in ONE engine call instead of a serial plane loop (:319-327).
* ``destripe_volume`` replaces ``producer`` / ``consumer`` (:797-906): instead of pickling
  819 MB blocks through a ``multiprocessing.Queue`` to ``co_cpus`` processes, a reader thread
  fills pinned host buffers, the engine streams them through the GPU (H2D / compute / D2H on
  three CUDA streams inside ``dstr_filter_chunk``) and a writer thread stores the result.
  Decode (read), device and write times are reported separately.
* ``z_slab`` partitions a tile into contiguous Z-slabs aligned to the output zarr's Z-chunk
  (64, :1069) for the one-process-per-GPU launch; planes are independent, so there is no
  collective.
"""

from __future__ import annotations

import queue
import threading
import time
from typing import Optional, Sequence, Tuple

import numpy as np

from . import engine as _eng
from . import filtering as fl


# ----------------------------------------------------------------------------------------------
# zero-overlap restatement of the two aind-large-scale-prediction helpers execute_worker uses
# (SURVEY.md Appendix A.7); the reference always runs with overlap (0, 0, 0)
# (zarr_destriper.py:1018-1022).
def recover_global_position(super_chunk_slice: Sequence[slice], internal_slices: Sequence[slice]):
    pos = []
    for sc, inner in zip(super_chunk_slice, internal_slices):
        base = sc.start or 0
        pos.append(slice(base + (inner.start or 0), base + inner.stop))
    pos = tuple(pos)
    return pos, tuple(p.start for p in pos), tuple(p.stop for p in pos)


def unpad_global_coords(global_coord_pos, block_shape, overlap_prediction_chunksize, dataset_shape):
    if any(int(o) != 0 for o in overlap_prediction_chunksize):
        raise NotImplementedError("only the reference's zero overlap is supported")
    local = tuple(slice(0, int(n)) for n in block_shape[-3:])
    return tuple(global_coord_pos[-3:]), local


def pad_array_n_d(arr: np.ndarray, dim: int = 5) -> np.ndarray:
    """Left-pad singleton axes up to ``dim`` dimensions (reference zarr_destriper.py:157-179)."""
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
    engine: Optional["_eng.DestripeEngine"] = None,
):
    """Destripe one ``(1, Z, H, W)`` chunk and write it into ``output_destriped_zarr``.

    Mirrors reference ``execute_worker`` (zarr_destriper.py:253-336): same slice arithmetic,
    ``microscope_high_int=2500`` (:326), result written with ``__setitem__`` into the 5-D
    (or N-D) sink.  ``data`` may be float32 (reference DataLoader contract, :1049) or uint16.
    """
    data = np.squeeze(np.asarray(data), axis=0)
    global_coord_pos, _, _ = recover_global_position(batch_super_chunk, batch_internal_slice)
    unpadded_global_slice, unpadded_local_slice = unpad_global_coords(
        global_coord_pos=global_coord_pos,
        block_shape=data.shape,
        overlap_prediction_chunksize=overlap_prediction_chunksize,
        dataset_shape=output_destriped_zarr.shape,
    )
    unpadded_local_slice = list((slice(0, 1), slice(0, 1)) + tuple(unpadded_local_slice))
    output_slices = list((slice(0, 1), slice(0, 1)) + tuple(unpadded_global_slice))
    for idx in range(output_destriped_zarr.ndim):
        if output_slices[idx].stop > output_destriped_zarr.shape[idx]:
            rest = output_slices[idx].stop - output_destriped_zarr.shape[idx]
            unpadded_local_slice[idx] = slice(
                unpadded_local_slice[idx].start, unpadded_local_slice[idx].stop - rest
            )
            output_slices[idx] = slice(output_slices[idx].start, output_destriped_zarr.shape[idx])
    output_slices = tuple(output_slices)
    unpadded_local_slice = tuple(unpadded_local_slice)

    input_tile_path = dataset_name.replace(".zarr", "")
    filtered = fl.filter_planes(
        data,
        input_tile_path=input_tile_path,
        no_cells_config=no_cells_config,
        cells_config=cells_config,
        shadow_correction=shadow_correction,
        microscope_high_int=2500,
        engine=engine,
    )
    if filtered.dtype != np.uint16:
        # no shadow correction: the reference's implicit float -> uint16 store truncates;
        # saturate instead of wrapping
        filtered = np.clip(filtered, 0, 65535)
    filtered = pad_array_n_d(arr=filtered[unpadded_local_slice[2:]], dim=output_destriped_zarr.ndim)
    output_destriped_zarr[output_slices] = filtered


# ----------------------------------------------------------------------------------------------
def compute_pyramid(data, n_lvls: int, scale_axis, chunks="auto", engine: Optional["_eng.DestripeEngine"] = None):
    """Multiscale levels ``[level 0, ..., level n_lvls-1]`` of a uint16 volume on the GPU.

    Mirrors reference ``compute_pyramid`` (zarr_destriper.py:365-407): windowed mean with
    ``preserve_dtype`` (truncation), every level built from the previous one; only the
    reference's scale ``2`` on the last three axes (leading axes 1) is implemented.
    """
    data = np.asarray(data)
    scale_axis = tuple(int(s) for s in scale_axis)
    if data.dtype != np.uint16 or data.ndim < 3:
        raise NotImplementedError("B200 engine: compute_pyramid expects a uint16 array with >= 3 dimensions")
    if scale_axis[-3:] != (2, 2, 2) or any(s != 1 for s in scale_axis[:-3]) or len(scale_axis) != data.ndim:
        raise NotImplementedError("B200 engine: only scale (…, 1, 2, 2, 2) is implemented")
    lead = data.shape[:-3]
    if int(np.prod(lead)) != 1:
        raise NotImplementedError("B200 engine: leading (T, C) axes must be singleton")
    levels = [data]
    vol = data.reshape(data.shape[-3:])
    eng = engine
    for _ in range(1, n_lvls):
        if eng is None:
            eng = _eng.get_engine(vol.shape[1], vol.shape[2])
        vol = eng.downscale2x(vol)
        levels.append(vol.reshape(lead + vol.shape))
    return levels


def z_slab(n_planes: int, rank: int, world_size: int, align: int = 64) -> Tuple[int, int]:
    """Contiguous Z range ``[z0, z1)`` of ``rank``; boundaries are multiples of ``align``."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    n_blocks = (n_planes + align - 1) // align
    base, extra = divmod(n_blocks, world_size)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return min(b0 * align, n_planes), min(b1 * align, n_planes)


def destripe_volume(
    volume,
    output,
    no_cells_config: dict,
    cells_config: dict,
    shadow_correction: Optional[dict] = None,
    dataset_name: str = "0_0.zarr",
    chunk_planes: int = 64,
    z_range: Optional[Tuple[int, int]] = None,
    device: Optional[int] = None,
    microscope_high_int: int = 2500,
    queue_depth: int = 2,
    pyramid_outputs: Optional[Sequence] = None,
    io_threads: int = 4,
):
    """Stream ``volume[z0:z1]`` (array-like ``(Z, H, W)``, uint16 or float32) through the GPU.

    ``output`` is any ``(Z, H, W)`` sink with ``__setitem__`` (numpy array, zarr array).
    ``pyramid_outputs`` (optional): up to two sinks ``(Z//2, H//2, W//2)`` and ``(Z//4, H//4, W//4)``
    that receive multiscale levels 1 and 2, computed on the device while each destriped chunk is
    still resident (requires ``shadow_correction``, i.e. uint16 output, and ``chunk_planes`` and
    the slab start to be multiples of 4).

    ``io_threads``: the reader and the writer each split a chunk into that many Z-ranges and move
    them concurrently (array slicing / Zarr decode releases the GIL), which is what the reference
    spreads over ``co_cpus`` processes.

    Returns a timing dict: ``read_s`` (decode / host I/O), ``device_s`` (pinned H2D + kernels +
    D2H inside the engine), ``write_s`` and ``wall_s``.
    """
    Z, H, W = volume.shape[-3:]
    z0, z1 = (0, Z) if z_range is None else z_range
    eng = _eng.DestripeEngine(H, W, max_planes=min(chunk_planes, 16), device=_eng.default_device() if device is None else device)
    in_dtype = np.uint16 if np.dtype(volume.dtype) == np.uint16 else np.float32
    out_dtype = np.uint16 if shadow_correction is not None else np.float32
    n_pyr = 0 if pyramid_outputs is None else len(pyramid_outputs)
    if n_pyr:
        if out_dtype != np.uint16 or chunk_planes % 4 or z0 % 4 or n_pyr > 2:
            raise ValueError("pyramid_outputs need uint16 output, chunk_planes % 4 == 0 and an aligned slab")
    n_buf = queue_depth + 1
    pyr_bufs = [[_eng.PinnedBuffer((max(chunk_planes >> (k + 1), 1), H >> (k + 1), W >> (k + 1)), np.uint16)
                 for k in range(n_pyr)] for _ in range(n_buf)]
    in_bufs = [_eng.PinnedBuffer((chunk_planes, H, W), in_dtype) for _ in range(n_buf)]
    out_bufs = [_eng.PinnedBuffer((chunk_planes, H, W), out_dtype) for _ in range(n_buf)]
    free_in: "queue.Queue[int]" = queue.Queue()
    free_out: "queue.Queue[int]" = queue.Queue()
    for i in range(n_buf):
        free_in.put(i)
        free_out.put(i)
    ready: "queue.Queue" = queue.Queue()
    done: "queue.Queue" = queue.Queue()
    times = dict(read_s=0.0, device_s=0.0, write_s=0.0)
    errors = []

    from concurrent.futures import ThreadPoolExecutor

    io_threads = max(1, int(io_threads))
    rpool = ThreadPoolExecutor(io_threads)
    wpool = ThreadPoolExecutor(io_threads)

    def _split(n):
        step = max(1, (n + io_threads - 1) // io_threads)
        return [(s0, min(s0 + step, n)) for s0 in range(0, n, step)]

    def _read_part(i, a, s0, s1):
        in_bufs[i].array[s0:s1] = volume[a + s0 : a + s1]

    def _write_part(j, a, s0, s1):
        res = out_bufs[j].array[s0:s1]
        output[a + s0 : a + s1] = res if out_dtype == np.uint16 else np.clip(res, 0, 65535)

    def reader():
        try:
            for a in range(z0, z1, chunk_planes):
                b = min(a + chunk_planes, z1)
                i = free_in.get()
                t = time.perf_counter()
                list(rpool.map(lambda r: _read_part(i, a, r[0], r[1]), _split(b - a)))
                times["read_s"] += time.perf_counter() - t
                ready.put((i, a, b))
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
        finally:
            ready.put(None)

    def writer():
        try:
            while True:
                item = done.get()
                if item is None:
                    return
                j, a, b = item
                t = time.perf_counter()
                list(wpool.map(lambda r: _write_part(j, a, r[0], r[1]), _split(b - a)))
                for k in range(n_pyr):
                    sh = k + 1
                    n_k = (b - a) >> sh
                    pyramid_outputs[k][(a >> sh) : (a >> sh) + n_k] = pyr_bufs[j][k].array[:n_k]
                times["write_s"] += time.perf_counter() - t
                free_out.put(j)
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    t_wall = time.perf_counter()
    rt = threading.Thread(target=reader, daemon=True)
    wt = threading.Thread(target=writer, daemon=True)
    rt.start()
    wt.start()
    tile = dataset_name.replace(".zarr", "")
    try:
        while True:
            item = ready.get()
            if item is None:
                break
            i, a, b = item
            j = free_out.get()
            t = time.perf_counter()
            if n_pyr:
                eng.set_pyramid_outputs(pyr_bufs[j][0].array, pyr_bufs[j][1].array if n_pyr > 1 else None)
            fl.filter_planes(
                in_bufs[i].array[: b - a],
                tile,
                no_cells_config,
                cells_config,
                shadow_correction,
                microscope_high_int,
                out=out_bufs[j].array[: b - a],
                engine=eng,
            )
            times["device_s"] += time.perf_counter() - t
            free_in.put(i)
            done.put((j, a, b))
    finally:
        done.put(None)
        rt.join()
        wt.join()
        times["wall_s"] = time.perf_counter() - t_wall
        rpool.shutdown(wait=True)
        wpool.shutdown(wait=True)
        for pb in in_bufs + out_bufs + [p for ps in pyr_bufs for p in ps]:
            pb.free()
        eng.close()
    if errors:
        raise errors[0]
    times["planes"] = z1 - z0
    return times
