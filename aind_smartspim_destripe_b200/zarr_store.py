"""Minimal Zarr v2 directory store: what the tile driver needs when the ``zarr`` package is absent.

The reference opens tiles and creates its outputs with ``zarr`` / ``ome_zarr`` / ``numcodecs``
(``zarr_destriper.py:1027-1074``: uint16 arrays, chunks ``(1, 1, 64, 128, 128)``, blosc-zstd-3 with
byte shuffle, ``dimension_separator="/"``).  None of those packages is installed here, so this
module reads and writes the same on-disk format (``.zarray`` / ``.zgroup`` / ``.zattrs`` JSON, one
file per chunk, C order, full-size edge chunks padded with ``fill_value``) with the codecs the
standard library has (``null``, ``zlib``, ``bz2``, ``lzma``) plus the reference's own ``blosc``
(zstd / lz4 with byte shuffle: ``blosc1.py`` over the system ``libzstd`` / ``liblz4``); anything else
is delegated to ``numcodecs`` when it can be imported and otherwise rejected by name.  Chunk
decode / encode runs on a thread pool (the codecs release the GIL), which replaces the reference's
``co_cpus`` worker processes on the I/O side.
"""

from __future__ import annotations

import bz2
import itertools
import json
import lzma
import os
import threading
import shutil
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np


class _Codec:
    def __init__(self, config: Optional[dict], itemsize: int = 1):
        self.config = config
        cid = None if config is None else config.get("id")
        self.id = cid
        if cid is None:
            self.encode, self.decode = (lambda b: b), (lambda b: b)
        elif cid == "zlib":
            level = int(config.get("level", 1))
            self.encode, self.decode = (lambda b: zlib.compress(b, level)), zlib.decompress
        elif cid == "bz2":
            level = int(config.get("level", 1))
            self.encode, self.decode = (lambda b: bz2.compress(b, level)), bz2.decompress
        elif cid == "lzma":
            self.encode, self.decode = lzma.compress, lzma.decompress
        elif cid == "blosc" and _blosc_builtin_ok(config):
            # the reference's codec (zarr_destriper.py:1066-1074): Blosc1 frames over the system libzstd / liblz4
            from . import blosc1

            cname = config.get("cname", "lz4")
            clevel = int(config.get("clevel", 5))
            shuffle = int(config.get("shuffle", 1))
            if shuffle == -1:  # numcodecs AUTOSHUFFLE: byte shuffle unless the items are single bytes
                shuffle = 1 if itemsize > 1 else 0
            blocksize = int(config.get("blocksize", 0))
            self.encode = lambda b: blosc1.compress(b, itemsize, clevel, shuffle, cname, blocksize)
            self.decode = blosc1.decompress
        else:
            try:
                import numcodecs  # noqa: WPS433
            except ImportError as exc:
                raise NotImplementedError(
                    f"zarr compressor {config!r} needs numcodecs, which is not installed; supported without it: "
                    "null, zlib, bz2, lzma, blosc (zstd / lz4, no or byte shuffle)"
                ) from exc
            codec = numcodecs.get_codec(config)
            self.encode, self.decode = (lambda b: bytes(codec.encode(b))), (lambda b: bytes(codec.decode(b)))


def _blosc_builtin_ok(config: dict) -> bool:
    from . import blosc1

    return config.get("cname", "lz4") in blosc1.COMPRESSOR_CODE and int(config.get("shuffle", 1)) in (0, 1, -1) \
        and blosc1.available(config.get("cname", "lz4"))


def default_compressor() -> Optional[dict]:
    """The reference's output codec: blosc-zstd-3 with byte shuffle (zarr_destriper.py:1066-1074)."""
    return {"id": "blosc", "cname": "zstd", "clevel": 3, "shuffle": 1, "blocksize": 0}


def _norm_key(key, shape) -> Tuple[Tuple[slice, ...], Tuple[int, ...]]:
    """Basic indexing only: ints and step-1 slices; returns full slices and the axes to squeeze."""
    if not isinstance(key, tuple):
        key = (key,)
    if any(k is Ellipsis for k in key):
        i = key.index(Ellipsis)
        key = key[:i] + (slice(None),) * (len(shape) - len(key) + 1) + key[i + 1 :]
    key = key + (slice(None),) * (len(shape) - len(key))
    if len(key) != len(shape):
        raise IndexError("too many indices")
    out, squeeze = [], []
    for ax, (k, n) in enumerate(zip(key, shape)):
        if isinstance(k, (int, np.integer)):
            k = int(k) + (n if k < 0 else 0)
            if not 0 <= k < n:
                raise IndexError("index out of range")
            out.append(slice(k, k + 1))
            squeeze.append(ax)
        elif isinstance(k, slice):
            start, stop, step = k.indices(n)
            if step != 1:
                raise NotImplementedError("only step-1 slices")
            out.append(slice(start, max(start, stop)))
        else:
            raise NotImplementedError("only basic indexing")
    return tuple(out), tuple(squeeze)


class ZarrArray:
    """One Zarr v2 array in a directory (``<path>/.zarray`` + chunk files)."""

    def __init__(self, path, meta: dict, mode: str = "r", threads: int = 8):
        self.path = Path(path)
        self.meta = meta
        self.mode = mode
        if meta.get("zarr_format") != 2:
            raise NotImplementedError("only zarr_format 2")
        if meta.get("order", "C") != "C":
            raise NotImplementedError("only C-order arrays")
        if meta.get("filters"):
            raise NotImplementedError("zarr filters are not supported")
        self.shape = tuple(int(v) for v in meta["shape"])
        self.chunks = tuple(int(v) for v in meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill_value = meta.get("fill_value", 0) or 0
        self.sep = meta.get("dimension_separator", ".")
        self.codec = _Codec(meta.get("compressor"), self.dtype.itemsize)
        self.threads = max(1, int(threads))
        self._pool: Optional[ThreadPoolExecutor] = None
        self._made_dirs = set()

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def open(cls, path, mode: str = "r", threads: int = 8) -> "ZarrArray":
        with open(Path(path) / ".zarray") as fp:
            return cls(path, json.load(fp), mode, threads)

    @classmethod
    def create(cls, path, shape: Sequence[int], chunks: Sequence[int], dtype, compressor="default",
               dimension_separator: str = "/", fill_value=0, overwrite: bool = True, threads: int = 8) -> "ZarrArray":
        path = Path(path)
        if path.exists():
            if not overwrite:
                raise FileExistsError(str(path))
            shutil.rmtree(path)
        path.mkdir(parents=True)
        if compressor == "default":
            compressor = default_compressor()
        meta = {
            "zarr_format": 2,
            "shape": [int(v) for v in shape],
            "chunks": [int(min(c, s)) if s > 0 else int(c) for c, s in zip(chunks, shape)],
            "dtype": np.dtype(dtype).str,
            "compressor": compressor,
            "fill_value": fill_value,
            "order": "C",
            "filters": None,
            "dimension_separator": dimension_separator,
        }
        _Codec(compressor, np.dtype(dtype).itemsize)  # fail before anything is written if the codec is unavailable
        with open(path / ".zarray", "w") as fp:
            json.dump(meta, fp, indent=4)
        return cls(path, meta, "w", threads)

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def __repr__(self):
        return f"<ZarrArray {self.path} {self.shape} {self.dtype} chunks={self.chunks} codec={self.codec.id}>"

    # ---- chunk level ----------------------------------------------------------------------
    def _chunk_path(self, idx: Tuple[int, ...]) -> Path:
        return self.path / self.sep.join(str(i) for i in idx) if self.sep == "/" else self.path / ".".join(map(str, idx))

    def read_chunk(self, idx: Tuple[int, ...]) -> np.ndarray:
        p = self._chunk_path(idx)
        try:
            raw = p.read_bytes()
        except FileNotFoundError:
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        buf = self.codec.decode(raw)
        return np.frombuffer(buf, dtype=self.dtype).reshape(self.chunks)

    def write_chunk(self, idx: Tuple[int, ...], block: np.ndarray):
        if block.shape != self.chunks:
            raise ValueError("write_chunk needs a full chunk")
        p = self._chunk_path(idx)
        if p.parent not in self._made_dirs:
            p.parent.mkdir(parents=True, exist_ok=True)
            self._made_dirs.add(p.parent)
        block = np.ascontiguousarray(block, dtype=self.dtype)
        data = self.codec.encode(block.reshape(-1).view(np.uint8).data)  # buffer protocol: no intermediate bytes copy
        tmp = p.with_name(p.name + f".{os.getpid()}.{threading.get_ident()}.tmp")  # unique per writer thread
        with open(tmp, "wb") as fp:
            fp.write(data)
        os.replace(tmp, p)

    def _pool_map(self, fn, items):
        items = list(items)
        if self.threads == 1 or len(items) <= 1:
            for it in items:
                fn(it)
            return
        if self._pool is None:
            self._pool = ThreadPoolExecutor(self.threads)
        list(self._pool.map(fn, items))

    def _chunk_ranges(self, sel: Tuple[slice, ...]):
        per_axis = []
        for s, c in zip(sel, self.chunks):
            if s.stop <= s.start:
                return []
            per_axis.append(range(s.start // c, (s.stop - 1) // c + 1))
        return itertools.product(*per_axis)

    # ---- region access --------------------------------------------------------------------
    def __getitem__(self, key) -> np.ndarray:
        sel, squeeze = _norm_key(key, self.shape)
        out = np.empty(tuple(s.stop - s.start for s in sel), dtype=self.dtype)

        def load(idx):
            block = self.read_chunk(idx)
            src, dst = [], []
            for i, s, c in zip(idx, sel, self.chunks):
                lo, hi = max(s.start, i * c), min(s.stop, (i + 1) * c)
                src.append(slice(lo - i * c, hi - i * c))
                dst.append(slice(lo - s.start, hi - s.start))
            out[tuple(dst)] = block[tuple(src)]

        self._pool_map(load, self._chunk_ranges(sel))
        return out.squeeze(axis=squeeze) if squeeze else out

    def __setitem__(self, key, value):
        if self.mode == "r":
            raise PermissionError("array opened read-only")
        sel, squeeze = _norm_key(key, self.shape)
        region = tuple(s.stop - s.start for s in sel)
        value = np.asarray(value)
        if squeeze:
            value = np.expand_dims(value, squeeze) if value.ndim == len(region) - len(squeeze) else value
        value = np.broadcast_to(value, region) if value.shape != region else value
        if value.dtype != self.dtype:
            value = value.astype(self.dtype)  # like zarr: implicit cast on assignment

        def store(idx):
            src, dst, full = [], [], True
            for i, s, c, n in zip(idx, sel, self.chunks, self.shape):
                lo, hi = max(s.start, i * c), min(s.stop, (i + 1) * c)
                src.append(slice(lo - s.start, hi - s.start))
                dst.append(slice(lo - i * c, hi - i * c))
                full &= lo == i * c and hi == min((i + 1) * c, n)
            part = value[tuple(src)]
            if full and part.shape == self.chunks:  # a whole interior chunk: encode the slice itself
                self.write_chunk(idx, part)
                return
            if full:  # covers every stored element of an edge chunk: pad, no read-modify-write
                block = np.full(self.chunks, self.fill_value, dtype=self.dtype)
            else:
                block = self.read_chunk(idx).copy()
            block[tuple(dst)] = part
            self.write_chunk(idx, block)

        self._pool_map(store, self._chunk_ranges(sel))

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None


# ---- groups ------------------------------------------------------------------------------------
def create_group(path, attrs: Optional[dict] = None, overwrite: bool = False) -> Path:
    path = Path(path)
    if overwrite and path.exists():
        shutil.rmtree(path)
    path.mkdir(parents=True, exist_ok=True)
    with open(path / ".zgroup", "w") as fp:
        json.dump({"zarr_format": 2}, fp)
    if attrs is not None:
        write_attrs(path, attrs)
    return path


def write_attrs(path, attrs: dict):
    with open(Path(path) / ".zattrs", "w") as fp:
        json.dump(attrs, fp, indent=4)


def read_attrs(path) -> dict:
    p = Path(path) / ".zattrs"
    if not p.exists():
        return {}
    with open(p) as fp:
        return json.load(fp)


def open_array(path, mode: str = "r", threads: int = 8):
    """The ``zarr`` package's array when it is importable, else :class:`ZarrArray`."""
    try:
        import zarr  # noqa: WPS433

        return zarr.open(str(path), mode="r" if mode == "r" else "r+")
    except ImportError:
        return ZarrArray.open(path, mode, threads)
