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


# Engine contexts + pinned staging buffers of destripe_volume, kept between calls: page-locking a few GB and creating
# the contexts costs 2-3 s, several times the streaming time of a whole tile, and the tiles of a channel (or the
# chunks of a benchmark) all have the same geometry.  At most `_POOL_MAX` idle sets are kept; `release_volume_resources`
# (also registered with atexit) frees them.
_pool_lock = threading.Lock()
_pool: "dict[tuple, list]" = {}
_POOL_MAX = 2


def _pool_take(key):
    with _pool_lock:
        sets = _pool.get(key)
        if sets:
            return sets.pop()
    return None


def _free_resource_set(res):
    for pb in res["in"] + res["out"] + [p for ps in res["pyr"] for p in ps]:
        pb.free()
    for e in res["engines"]:
        e.close()


def _pool_give(key, res):
    evicted = []
    with _pool_lock:
        _pool.setdefault(key, []).append(res)
        idle = [(k, r) for k, v in _pool.items() for r in v]
        while len(idle) > _POOL_MAX:
            k, r = idle.pop(0)  # oldest first (dict and list order)
            _pool[k].remove(r)
            evicted.append(r)
    for r in evicted:
        _free_resource_set(r)


def release_volume_resources():
    """Free the idle engine contexts / pinned buffers kept by ``destripe_volume(..., reuse_resources=True)``."""
    with _pool_lock:
        sets = [r for v in _pool.values() for r in v]
        _pool.clear()
    for r in sets:
        _free_resource_set(r)


import atexit as _atexit

_atexit.register(release_volume_resources)


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
    device_workers: int = 2,
    reuse_resources: bool = True,
):
    """Stream ``volume[z0:z1]`` (array-like ``(Z, H, W)``, uint16 or float32) through the GPU.

    ``output`` is any ``(Z, H, W)`` sink with ``__setitem__`` (numpy array, zarr array).
    ``pyramid_outputs`` (optional): up to two sinks ``(Z//2, H//2, W//2)`` and ``(Z//4, H//4, W//4)``
    that receive multiscale levels 1 and 2, computed on the device while each destriped chunk is
    still resident (requires ``shadow_correction``, i.e. uint16 output, and ``chunk_planes`` and
    the slab start to be multiples of 4).

    ``io_threads``: the reader and the writer each split a chunk into that many Z-ranges and move
    them concurrently (array slicing / Zarr decode releases the GIL), which is what the reference
    spreads over ``co_cpus`` processes.  Ranges written to a chunked sink (``output.chunks``) are cut
    on its Z-chunk boundaries, so no two threads ever touch the same stored chunk.
    ``device_workers``: engine contexts fed from the same queue (each call is synchronous in its own
    host thread), so the upload / kernels / download of consecutive chunks overlap.

    ``reuse_resources``: keep the engine contexts and pinned buffers for the next call with the same geometry
    (``release_volume_resources()`` frees them); ``setup_s`` / ``teardown_s`` of the returned dict show the cost.

    A failure in the reader, a device worker or the writer stops all of them and is re-raised here
    (the reference's consumers would hang on ``join``, zarr_destriper.py:1171).

    Returns a timing dict: ``read_s`` (decode / host I/O), ``device_s`` (pinned H2D + kernels +
    D2H inside the engine, summed over the workers), ``write_s`` and ``wall_s``.
    """
    from concurrent.futures import ThreadPoolExecutor

    Z, H, W = volume.shape[-3:]
    z0, z1 = (0, Z) if z_range is None else z_range
    in_dtype = np.uint16 if np.dtype(volume.dtype) == np.uint16 else np.float32
    out_dtype = np.uint16 if shadow_correction is not None else np.float32
    n_pyr = 0 if pyramid_outputs is None else len(pyramid_outputs)
    if n_pyr:
        if out_dtype != np.uint16 or chunk_planes % 4 or z0 % 4 or n_pyr > 2:
            raise ValueError("pyramid_outputs need uint16 output, chunk_planes % 4 == 0 and an aligned slab")
    device_workers = max(1, int(device_workers))
    dev = _eng.default_device() if device is None else device
    n_buf = max(2, queue_depth) + device_workers - 1
    engines, in_bufs, out_bufs, pyr_bufs = [], [], [], []
    rpool = wpool = None
    stop = threading.Event()
    errors = []
    errors_lock = threading.Lock()

    def _fail(exc):
        with errors_lock:
            errors.append(exc)
        stop.set()

    def _get(q):
        """Blocking get that gives up when another stage failed."""
        while True:
            try:
                return q.get(timeout=0.2)
            except queue.Empty:
                if stop.is_set():
                    raise _Stopped()

    io_threads = max(1, int(io_threads))
    sink_chunks = getattr(output, "chunks", None)
    z_chunk = int(sink_chunks[-3]) if sink_chunks is not None and len(sink_chunks) >= 3 else 1

    def _split(a, n, align=1):
        """Z-ranges [s0, s1) (relative to ``a``) cut at multiples of ``align`` in absolute Z."""
        step = max(1, (n + io_threads - 1) // io_threads)
        if align > 1:
            step = ((step + align - 1) // align) * align
            first = min(n, (align - a % align) % align)  # leading partial stored chunk: one writer
            cuts = ([0] if first else []) + list(range(first, n, step))
        else:
            cuts = list(range(0, n, step))
        return [(c, min(nx, n)) for c, nx in zip(cuts, cuts[1:] + [n]) if c < n]

    # optional zero-copy protocols: a source with ``read_into(dst, z0, z1)`` fills the pinned buffer itself (a
    # decoder writing straight into it), a sink with ``write_from(src, z0, z1)`` consumes the pinned result
    read_into = getattr(volume, "read_into", None)
    write_from = getattr(output, "write_from", None)

    def _read_part(i, a, s0, s1):
        if read_into is not None:
            read_into(in_bufs[i].array[s0:s1], a + s0, a + s1)
        else:
            in_bufs[i].array[s0:s1] = volume[a + s0 : a + s1]

    def _write_part(j, a, s0, s1):
        res = out_bufs[j].array[s0:s1]
        if out_dtype != np.uint16:
            res = np.clip(res, 0, 65535)
        if write_from is not None:
            write_from(res, a + s0, a + s1)
        else:
            output[a + s0 : a + s1] = res

    free_in: "queue.Queue[int]" = queue.Queue()
    free_out: "queue.Queue[int]" = queue.Queue()
    ready: "queue.Queue" = queue.Queue()
    done: "queue.Queue" = queue.Queue()
    times = dict(read_s=0.0, device_s=0.0, write_s=0.0)
    times_lock = threading.Lock()

    def reader():
        try:
            for a in range(z0, z1, chunk_planes):
                b = min(a + chunk_planes, z1)
                i = _get(free_in)
                t = time.perf_counter()
                list(rpool.map(lambda r: _read_part(i, a, r[0], r[1]), _split(a, b - a)))
                times["read_s"] += time.perf_counter() - t
                ready.put((i, a, b))
        except _Stopped:
            pass
        except BaseException as exc:
            _fail(exc)
        finally:
            for _ in range(device_workers):
                ready.put(None)

    tile = dataset_name.replace(".zarr", "")

    def device_worker(eng):
        try:
            while True:
                item = _get(ready)
                if item is None:
                    return
                i, a, b = item
                j = _get(free_out)
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
                with times_lock:
                    times["device_s"] += time.perf_counter() - t
                free_in.put(i)
                done.put((j, a, b))
        except _Stopped:
            pass
        except BaseException as exc:
            _fail(exc)
        finally:
            done.put(None)

    def writer():
        try:
            live = device_workers
            while live:
                item = _get(done)
                if item is None:
                    live -= 1
                    continue
                j, a, b = item
                t = time.perf_counter()
                list(wpool.map(lambda r: _write_part(j, a, r[0], r[1]), _split(a, b - a, z_chunk)))
                for k in range(n_pyr):
                    sh = k + 1
                    n_k = (b - a) >> sh
                    pyramid_outputs[k][(a >> sh) : (a >> sh) + n_k] = pyr_bufs[j][k].array[:n_k]
                times["write_s"] += time.perf_counter() - t
                free_out.put(j)
        except _Stopped:
            pass
        except BaseException as exc:
            _fail(exc)

    t_wall = time.perf_counter()
    threads = []
    t_setup_end = t_stream_end = None
    try:
        # engine contexts and pinned buffers are created concurrently: page-locking GBs of host memory and the
        # first CUDA context / module load each take seconds when done one after the other
        pool_key = (int(dev), int(H), int(W), int(chunk_planes), np.dtype(in_dtype).str, np.dtype(out_dtype).str,
                    int(n_buf), int(device_workers), int(n_pyr))
        pooled = _pool_take(pool_key) if reuse_resources else None
        if pooled is not None:
            engines, in_bufs, out_bufs, pyr_bufs = pooled["engines"], pooled["in"], pooled["out"], pooled["pyr"]
        with ThreadPoolExecutor(max(4, device_workers)) as setup:
            if pooled is not None:
                setup = None
            f_eng = [] if setup is None else [setup.submit(_eng.DestripeEngine, H, W, min(chunk_planes, 16), dev) for _ in range(device_workers)]
            if setup is not None:
                f_in = [setup.submit(_eng.PinnedBuffer, (chunk_planes, H, W), in_dtype) for _ in range(n_buf)]
                f_out = [setup.submit(_eng.PinnedBuffer, (chunk_planes, H, W), out_dtype) for _ in range(n_buf)]
                f_pyr = [[setup.submit(_eng.PinnedBuffer, (max(chunk_planes >> (k + 1), 1), H >> (k + 1), W >> (k + 1)), np.uint16)
                          for k in range(n_pyr)] for _ in range(n_buf)]
                for f in f_eng:
                    engines.append(f.result())
                in_bufs = [f.result() for f in f_in]
                out_bufs = [f.result() for f in f_out]
                pyr_bufs = [[f.result() for f in fs] for fs in f_pyr]
        for i in range(n_buf):
            free_in.put(i)
            free_out.put(i)
        rpool = ThreadPoolExecutor(io_threads)
        wpool = ThreadPoolExecutor(io_threads)
        t_setup_end = time.perf_counter()
        threads = [threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)]
        threads += [threading.Thread(target=device_worker, args=(e,), daemon=True) for e in engines]
        for t in threads:
            t.start()
        # the writer ends when every device worker has ended; an error anywhere sets `stop`, and every
        # blocking get polls it, so no stage can wait forever on a stage that died
        for t in threads:
            while t.is_alive():
                t.join(timeout=0.05)
        t_stream_end = time.perf_counter()
    except BaseException as exc:  # KeyboardInterrupt included
        _fail(exc)
        for t in threads:
            t.join(timeout=5.0)
    finally:
        for pool in (rpool, wpool):
            if pool is not None:
                pool.shutdown(wait=True)
        res = dict(engines=engines, **{"in": in_bufs, "out": out_bufs, "pyr": pyr_bufs})
        complete = (len(engines) == device_workers and len(in_bufs) == n_buf and len(out_bufs) == n_buf)
        if reuse_resources and not errors and complete and t_stream_end is not None:
            for e in engines:
                if n_pyr:
                    e.set_pyramid_outputs(None, None)
            _pool_give(pool_key, res)
        else:
            _free_resource_set(res)
        t_end = time.perf_counter()
        times["wall_s"] = t_end - t_wall
        if t_setup_end is not None and t_stream_end is not None:
            # engine contexts + pinned buffers / the streaming pipeline itself / freeing them again
            times["setup_s"] = t_setup_end - t_wall
            times["stream_s"] = t_stream_end - t_setup_end
            times["teardown_s"] = t_end - t_stream_end
    if errors:
        raise errors[0]
    times["planes"] = z1 - z0
    return times


class _Stopped(Exception):
    """Raised inside a pipeline stage when another stage failed."""


# ----------------------------------------------------------------------------------------------
# Tile driver (SURVEY.md §8 "next" row f1): destripe_zarr / destripe_channel of the reference
# (zarr_destriper.py:909-1267) on destripe_volume.  Zarr I/O goes through ``zarr_store`` (the
# ``zarr`` package when importable); TIFF flats / darks through ``destriper.imread``.
# ----------------------------------------------------------------------------------------------
def get_microscope_flats(channel_name: str, derivatives_folder):
    """Per-hemisphere microscope flats + the X/Y-folder -> side map from ``metadata.json``
    (reference zarr_destriper.py:70-153).  Returns ``(flatfields | None, tile_config | None)``."""
    import json
    import re
    from pathlib import Path

    from .destriper import imread

    derivatives_folder = Path(derivatives_folder)
    waves = [p for p in str(channel_name).split("_") if p.isdigit()]
    metadata_path = derivatives_folder / "metadata.json"
    if not (metadata_path.exists() and waves):
        return None, None
    with open(metadata_path) as fp:
        tile_config = json.load(fp).get("tile_config")
    if tile_config is None:
        raise ValueError("Please, verify metadata.json")
    wave = int(waves[0])
    sides: dict = {}
    for entry in tile_config.values():
        if int(entry.get("Laser")) != wave:
            continue
        x_folder, y_folder, side = entry.get("X"), entry.get("Y"), entry.get("Side")
        if x_folder is None or y_folder is None or side is None:
            raise KeyError("Please, check the data in metadata.json")
        sides.setdefault(x_folder, {})[y_folder] = int(side)

    def natural(p):
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", p.name)]

    flats = [np.asarray(imread(p)) for p in sorted(derivatives_folder.glob(f"FlatReal{wave}_*.tif"), key=natural)]
    if len(flats) != 2:
        raise ValueError(f"Error while reading the microscope flatfields: found {len(flats)}, expected 2")
    return flats, sides


def _compute_scales(scale_num_levels, scale_factor, pixelsizes, chunks, data_shape):
    """coordinateTransformations + chunk shapes per level (reference :410-500)."""
    scale = [1.0, 1.0, float(pixelsizes[0]), float(pixelsizes[1]), float(pixelsizes[2])]
    z, y, x = data_shape[2:]
    transforms, chunk_sizes = [], []
    for lvl in range(scale_num_levels):
        if lvl:
            scale = scale[:2] + [scale[2 + i] * scale_factor[i] for i in range(3)]
            z, y, x = (int(np.ceil(v / f)) for v, f in zip((z, y, x), scale_factor))
        transforms.append([{"type": "scale", "scale": list(scale)}])
        chunk_sizes.append((1, 1, min(z, chunks[2]), min(y, chunks[3]), min(x, chunks[4])))
    return transforms, chunk_sizes


def _build_ome(data_shape, image_name, channel_colors, channel_minmax, channel_startend):
    """``omero`` block (reference :531-597)."""
    channels = []
    for i in range(data_shape[1]):
        channels.append({
            "active": True, "coefficient": 1, "color": f"{channel_colors[i]:06x}", "family": "linear",
            "inverted": False, "label": image_name,
            "window": {"end": float(channel_startend[i][1]), "max": float(channel_minmax[i][1]),
                       "min": float(channel_minmax[i][0]), "start": float(channel_startend[i][0])},
        })
    return {"id": 1, "name": image_name, "version": "0.4", "channels": channels,
            "rdefs": {"defaultT": 0, "defaultZ": data_shape[2] // 2, "model": "color"}}


def ome_ngff_metadata(data_shape, chunks, image_name, n_lvls, scale_factors, voxel_size) -> dict:
    """The ``.zattrs`` the reference writes next to the levels (OME-NGFF 0.4: ``multiscales`` with
    one scale transform per level, ``omero`` display block; reference :600-674 and :703-727)."""
    transforms, _ = _compute_scales(n_lvls, scale_factors, voxel_size, chunks, data_shape)
    axes = [
        {"name": "t", "type": "time", "unit": "millisecond"},
        {"name": "c", "type": "channel"},
        {"name": "z", "type": "space", "unit": "micrometer"},
        {"name": "y", "type": "space", "unit": "micrometer"},
        {"name": "x", "type": "space", "unit": "micrometer"},
    ]
    datasets = [{"path": str(i), "coordinateTransformations": transforms[i]} for i in range(n_lvls)]
    n_c = data_shape[1]
    info = np.iinfo(np.uint16)
    return {
        "omero": _build_ome(data_shape, image_name, [0x690AFE] * n_c, [(info.min, info.max)] * n_c,
                            [(0.0, 350.0)] * n_c),
        "multiscales": [{"version": "0.4", "axes": axes, "datasets": datasets}],
    }


class _PlanesView:
    """(Z, H, W) view of one (t, c) of a 5-D TCZYX array; Z-range indexing only."""

    def __init__(self, arr, t: int = 0, c: int = 0):
        self.arr, self.t, self.c = arr, t, c
        self.shape = tuple(arr.shape[2:])
        self.dtype = np.dtype(arr.dtype)
        ch = getattr(arr, "chunks", None)
        if ch is not None:
            self.chunks = tuple(ch[2:])  # lets destripe_volume cut concurrent writes on stored-chunk boundaries

    def __getitem__(self, key):
        return np.asarray(self.arr[self.t, self.c, key])

    def __setitem__(self, key, value):
        self.arr[self.t, self.c, key] = value


def destripe_zarr(
    dataset_path,
    multiscale: str,
    output_destriped_zarr,
    prediction_chunksize: Tuple[int, ...],
    target_size_mb: int,
    n_workers: int,
    batch_size: int,
    super_chunksize: Optional[Tuple[int, ...]],
    results_folder,
    derivatives_path,
    xyz_resolution,
    parameters: dict,
    flatfield=None,
    lazy_callback_fn=None,
    n_levels: int = 3,
    compressor="default",
    rank: Optional[int] = None,
    world_size: Optional[int] = None,
):
    """Destripe one tile ``<dataset_path>/<multiscale>`` (TCZYX) into ``<output_destriped_zarr>/0``
    and write the 2x multiscale levels + OME-NGFF metadata (reference zarr_destriper.py:909-1211).

    Same arguments as the reference.  ``target_size_mb``, ``batch_size`` and ``super_chunksize``
    described its DataLoader and are accepted and ignored: chunks of ``prediction_chunksize[0]``
    planes stream through pinned buffers instead.  ``n_workers`` is the number of chunk
    decode / encode threads (0 = 8).  With ``world_size`` > 1 (default: ``WORLD_SIZE`` / ``RANK`` of
    a torchrun launch) every rank processes its own Z-slab; there is no collective.
    """
    from pathlib import Path

    from . import zarr_store as zs
    from .destriper import imread
    from .distributed import env_rank

    if lazy_callback_fn is not None:
        raise NotImplementedError("lazy_callback_fn is not supported by the GPU tile driver")
    no_cells_config, cells_config = parameters["no_cells_config"], parameters["cells_config"]
    from_env = rank is None or world_size is None  # explicit rank / world_size: the caller runs rank 0 first
    if from_env:
        rank, world_size, _ = env_rank()
    threads = int(n_workers) if n_workers and n_workers > 0 else 8
    output_destriped_zarr = Path(output_destriped_zarr)
    dataset_name = output_destriped_zarr.name

    src = zs.open_array(Path(dataset_path) / str(multiscale), "r", threads)
    if len(src.shape) != 5:
        raise ValueError(f"expected a TCZYX array, got shape {src.shape}")
    T, C, Z, H, W = src.shape

    # dark / flat exactly as the reference resolves them (:1106-1138)
    darkfield, tile_config = None, None
    retrospective = flatfield is not None
    derivatives_path = Path(derivatives_path)
    if derivatives_path.exists():
        dark_path = derivatives_path / "DarkMaster_cropped.tif"
        if not dark_path.exists():
            raise FileNotFoundError(f"Please, provide the current dark from the microscope! Provided path: {dark_path}")
        darkfield = np.asarray(imread(dark_path))
        if flatfield is None:
            flatfield, tile_config = get_microscope_flats(output_destriped_zarr.parent.name, derivatives_path)
            if flatfield is not None:
                flatfield = fl.normalize_image(flatfield)
    shadow = None
    if flatfield is not None:
        if darkfield is None:
            raise FileNotFoundError(f"a flat field needs the microscope dark: {derivatives_path} does not exist")
        shadow = {"retrospective": retrospective, "flatfield": flatfield, "darkfield": darkfield,
                  "tile_config": tile_config}

    chunk_planes = int(prediction_chunksize[-3])
    if len(prediction_chunksize) >= 3 and (int(prediction_chunksize[-2]) < H or int(prediction_chunksize[-1]) < W):
        import warnings

        warnings.warn(
            f"prediction_chunksize {tuple(prediction_chunksize)} is smaller than the plane ({H}, {W}): the reference "
            "would filter every YX block on its own (zarr_destriper.py:253-336); this driver always filters whole "
            "planes, so the result differs from the reference's per-block output for this chunking",
            stacklevel=2,
        )
    # stream whole source chunks when that stays small: a 128-deep source chunk read 64 planes at a
    # time would be decoded twice
    src_cz = int(getattr(src, "chunks", (1, 1, chunk_planes))[2])
    if src_cz > chunk_planes and int(np.lcm(src_cz, chunk_planes)) <= 256:
        chunk_planes = int(np.lcm(src_cz, chunk_planes))
    out_chunks = (1, 1, 64, 128, 128)
    fused = shadow is not None and n_levels in (2, 3) and chunk_planes % 4 == 0
    shapes = [(T, C, Z >> k, H >> k, W >> k) for k in range(n_levels)]
    # slab boundaries: no two ranks inside one stored chunk of any level (a chunk of level k spans
    # min(64, Z >> k) << k planes of level 0), and whole streaming chunks
    align = chunk_planes
    for k in range(n_levels):
        align = int(np.lcm(align, max(1, min(out_chunks[2], shapes[k][2])) << k))
    if rank == 0:
        zs.create_group(output_destriped_zarr, overwrite=True)
        for k in range(n_levels):
            zs.ZarrArray.create(output_destriped_zarr / str(k), shapes[k], out_chunks, np.uint16, compressor, "/")
        voxel = [xyz_resolution[-1], xyz_resolution[-2], xyz_resolution[-3]]
        zs.write_attrs(output_destriped_zarr,
                       ome_ngff_metadata(shapes[0], out_chunks, dataset_name, n_levels, [2, 2, 2], voxel))
    if world_size > 1 and from_env:
        from . import distributed as dist

        dist.init()
        dist.barrier()  # rank 0 has created the arrays
    levels = [zs.ZarrArray.open(output_destriped_zarr / str(k), "w", threads) for k in range(n_levels)]
    z0, z1 = z_slab(Z, rank, world_size, align)
    totals = dict(read_s=0.0, device_s=0.0, write_s=0.0, wall_s=0.0, setup_s=0.0, stream_s=0.0, teardown_s=0.0, planes=0)
    for t in range(T):
        for c in range(C):
            pyr = [_PlanesView(levels[k], t, c) for k in range(1, n_levels)] if fused else None
            if z1 > z0:
                tm = destripe_volume(_PlanesView(src, t, c), _PlanesView(levels[0], t, c), no_cells_config, cells_config,
                                     shadow, dataset_name=dataset_name, chunk_planes=chunk_planes, z_range=(z0, z1),
                                     pyramid_outputs=pyr, io_threads=1)
                for k in totals:
                    totals[k] += tm.get(k, 0.0)
            if not fused and n_levels > 1 and z1 > z0:
                # float output (no shadow correction) or unusual chunking: levels from the written data
                # (streamed in pieces of whole coarsest-level windows, so host memory stays at one piece)
                win = 1 << (n_levels - 1)
                a0 = z0 - z0 % win
                piece = int(np.lcm(max(chunk_planes, 64), win))
                for p0 in range(a0, z1, piece):
                    prev = np.asarray(levels[0][t, c, p0 : min(z1, p0 + piece)])
                    for k in range(1, n_levels):
                        if prev.shape[0] < 2:
                            break  # a trailing odd plane has no window at this level (floor semantics)
                        prev = compute_pyramid(prev, 2, [2, 2, 2])[-1]
                        room = max(0, shapes[k][2] - (p0 >> k))
                        if prev.shape[0] and room:
                            levels[k][t, c, (p0 >> k) : (p0 >> k) + min(room, prev.shape[0])] = prev[:room]
    for lv in levels:
        lv.close()
    if hasattr(src, "close"):
        src.close()
    return totals


def tiles_of_rank(tiles: Sequence, rank: int, world_size: int, tile_parallel: Optional[bool] = None):
    """Tiles a rank processes.  ``tile_parallel`` (default: as soon as there are at least as many tiles as
    ranks): whole tiles are dealt round-robin to the ranks (BASELINE config 5: 8 tiles on 8 GPUs, SURVEY.md
    section 8e); otherwise every rank takes part in every tile with its own Z-slab.  Returns (tiles, flag)."""
    tiles = list(tiles)
    if tile_parallel is None:
        tile_parallel = world_size > 1 and len(tiles) >= world_size
    if not tile_parallel or world_size <= 1:
        return tiles, False
    return [t for i, t in enumerate(tiles) if i % world_size == rank], True


def destripe_channel(zarr_dataset_path, derivatives_path, channel_name, results_folder, xyz_resolution,
                     estimated_channel_flats, laser_tiles, parameters, tile_parallel: Optional[bool] = None):
    """Every ``*.zarr`` tile of a channel with the flat field of its laser side
    (reference zarr_destriper.py:1214-1267, which loops the tiles sequentially :1231).

    Under a multi-rank launch (``WORLD_SIZE`` / ``RANK``) the tiles are independent units of work: with
    ``tile_parallel`` each rank destripes whole tiles on its own GPU (no collective, no shared chunk);
    otherwise all ranks work on one tile at a time as Z-slabs."""
    from pathlib import Path

    from .destriper import imread
    from .distributed import env_rank

    rank, world_size, _ = env_rank()
    channel_dataset = Path(zarr_dataset_path) / channel_name
    destriped_data_folder = Path(results_folder) / "destriped_data"
    destriped_data_folder.mkdir(parents=True, exist_ok=True)
    timings = {}
    mine, per_tile = tiles_of_rank(sorted(channel_dataset.glob("*.zarr")), rank, world_size, tile_parallel)
    for tile_path in mine:
        output_folder = destriped_data_folder / channel_name / tile_path.name
        flatfield_path = None
        for side, tiles in laser_tiles.items():
            if tile_path.stem.rsplit(".", 1)[0] in tiles:
                flatfield_path = estimated_channel_flats[int(side)]
                break
        if flatfield_path is None:
            raise ValueError(f"Tile {tile_path} not found in {laser_tiles}")
        flatfield = np.asarray(imread(str(flatfield_path)))
        timings[tile_path.name] = destripe_zarr(
            dataset_path=tile_path, multiscale="0", output_destriped_zarr=output_folder,
            prediction_chunksize=(64, 1600, 2000), target_size_mb=3072, n_workers=0, batch_size=1,
            super_chunksize=(384, 1600, 2000), results_folder=results_folder, derivatives_path=derivatives_path,
            xyz_resolution=xyz_resolution, parameters=parameters, flatfield=flatfield, lazy_callback_fn=None,
            # a rank that owns the whole tile runs it like a single-process job (no slab split, no barrier)
            rank=0 if per_tile else None, world_size=1 if per_tile else None)
    return timings
