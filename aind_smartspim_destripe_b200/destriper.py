"""TIFF / RAW directory front-end on the B200 engine (SURVEY.md §8 "next" row f4).

Mirrors ``/root/reference/code/aind_smartspim_destripe/destriper.py`` (``imsave`` :49-110,
``read_filter_save`` :113-227, ``_find_all_images`` :230-264, ``batch_filter`` :267-378) and
``readers.py`` (``raw_imread`` :34-61, ``imread`` :64-89): walk a directory tree, read every
supported image, destripe it with ``filter_stripes`` semantics, write it under the same relative
path.  Instead of ``multiprocessing.Pool.imap`` over single images, ``workers`` I/O threads feed
same-shape images in batches of ``chunks`` planes to one GPU engine call.

File formats: ``.raw`` (8-byte width/height header + uint16, either byte order) is built in;
TIFF uses ``tifffile`` when it is importable and otherwise a minimal baseline codec
(uncompressed, single-sample strips) that covers the SmartSPIM acquisition files; PNG goes through
``imageio`` like the reference when it is importable, else through a built-in codec (8 / 16-bit,
non-interlaced; inflate in Python, row filters in the native library).
"""

from __future__ import annotations

import logging
import os
import shutil
import struct
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List, Optional, Union

import numpy as np

from . import filtering as fl

PathLike = Union[Path, str]
SUPPORTED_READING_EXTENSIONS = [".tif", ".tiff", ".raw", ".png"]
SUPPORTED_OUTPUT_EXTENSIONS = [".tif", ".tiff", ".png"]
logger = logging.getLogger(__name__)

try:  # pragma: no cover - not installed in the build container
    import tifffile as _tifffile
except ImportError:
    _tifffile = None
try:  # pragma: no cover
    import imageio as _iio
except ImportError:
    _iio = None


def _get_extension(path) -> str:
    return Path(path).suffix


# ---------------------------------------------------------------------------- minimal TIFF codec
_TIFF_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 16: "Q"}
_SAMPLE_DTYPES = {(1, 8): "u1", (1, 16): "u2", (1, 32): "u4", (2, 8): "i1", (2, 16): "i2", (2, 32): "i4",
                  (3, 32): "f4", (3, 64): "f8"}


def _tiff_read(path: str) -> np.ndarray:
    """Baseline TIFF: first IFD, uncompressed, one sample per pixel, strips."""
    with open(path, "rb") as fp:
        data = fp.read()
    bo = {b"II": "<", b"MM": ">"}.get(data[:2])
    if bo is None or struct.unpack(bo + "H", data[2:4])[0] != 42:
        raise ValueError(f"{path}: not a classic TIFF file")
    (ifd,) = struct.unpack(bo + "I", data[4:8])
    (n,) = struct.unpack(bo + "H", data[ifd : ifd + 2])
    tags = {}
    for k in range(n):
        off = ifd + 2 + 12 * k
        tag, typ, count = struct.unpack(bo + "HHI", data[off : off + 8])
        size = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 16: 8}.get(typ, 0) * count
        if size == 0:
            continue
        pos = off + 8 if size <= 4 else struct.unpack(bo + "I", data[off + 8 : off + 12])[0]
        fmt = {1: "B", 3: "H", 4: "I", 16: "Q"}.get(typ)
        if fmt is None:
            continue
        tags[tag] = struct.unpack(bo + fmt * count, data[pos : pos + size])
    width, height = tags[256][0], tags[257][0]
    bits = tags.get(258, (1,))[0]
    if tags.get(259, (1,))[0] != 1 or tags.get(277, (1,))[0] != 1:
        raise NotImplementedError(f"{path}: only uncompressed single-sample TIFFs are supported without tifffile")
    fmt = tags.get(339, (1,))[0]
    dt = np.dtype(bo + _SAMPLE_DTYPES[(fmt, bits)])
    offsets, counts = tags[273], tags.get(279)
    if counts is None:
        counts = (width * height * dt.itemsize,)
    buf = b"".join(data[o : o + c] for o, c in zip(offsets, counts))
    return np.frombuffer(buf, dtype=dt, count=width * height).reshape(height, width)


def _tiff_write(path: str, img: np.ndarray):
    """Little-endian uncompressed single-strip baseline TIFF."""
    img = np.ascontiguousarray(img)
    if img.ndim != 2:
        raise NotImplementedError("minimal TIFF writer handles 2-D images")
    kind = {"u": 1, "i": 2, "f": 3}[img.dtype.kind]
    raw = img.astype(img.dtype.newbyteorder("<"), copy=False).tobytes()
    h, w = img.shape
    entries = [
        (256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, img.dtype.itemsize * 8), (259, 3, 1, 1), (262, 3, 1, 1),
        (273, 4, 1, 8), (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(raw)), (339, 3, 1, kind),
    ]
    ifd_off = 8 + len(raw) + (len(raw) & 1)
    out = bytearray(b"II" + struct.pack("<HI", 42, ifd_off) + raw + (b"\0" if len(raw) & 1 else b""))
    out += struct.pack("<H", len(entries))
    for tag, typ, count, value in entries:
        out += struct.pack("<HHI", tag, typ, count) + (struct.pack("<HH", value, 0) if typ == 3 else struct.pack("<I", value))
    out += struct.pack("<I", 0)
    with open(path, "wb") as fp:
        fp.write(out)


# ---------------------------------------------------------------------------- readers.py mirror
_PNG_SIG = b"\x89PNG\r\n\x1a\n"


def _png_read(path) -> np.ndarray:
    """Built-in PNG reader for the acquisition formats (8 / 16-bit greyscale, gray+alpha, RGB(A); non-interlaced):
    chunk parsing and inflate here, row reconstruction in the native library (``dstr_png_unfilter``)."""
    import ctypes as C
    import zlib

    from . import engine as _eng

    with open(path, "rb") as fp:
        buf = fp.read()
    if buf[:8] != _PNG_SIG:
        raise OSError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos + 8 <= len(buf):
        n, kind = struct.unpack(">I4s", buf[pos : pos + 8])
        body = buf[pos + 8 : pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    if hdr is None:
        raise OSError(f"{path}: missing IHDR")
    w, h, depth, ctype, _, _, interlace = hdr
    channels = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if channels is None or depth not in (8, 16) or interlace:
        raise NotImplementedError(f"{path}: PNG colour type {ctype} / depth {depth} / interlace {interlace} is not supported without imageio")
    bpp = channels * depth // 8
    stride = w * bpp
    scan = zlib.decompress(b"".join(idat))
    if len(scan) != h * (stride + 1):
        raise OSError(f"{path}: truncated PNG data")
    out = np.empty(h * stride, dtype=np.uint8)
    rc = _eng.load_library().dstr_png_unfilter(scan, h, stride, bpp, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OSError(f"{path}: bad PNG filter byte")
    img = out.view(">u2").astype(np.uint16) if depth == 16 else out
    return img.reshape(h, w) if channels == 1 else img.reshape(h, w, channels)


def _png_write(path, img, compression=1):
    """Greyscale 8 / 16-bit PNG (filter type 0, zlib level ``compression``)."""
    import zlib

    img = np.asarray(img)
    if img.ndim != 2 or img.dtype not in (np.uint8, np.uint16):
        raise NotImplementedError("the built-in PNG writer stores 2-D uint8 / uint16 images")
    h, w = img.shape
    depth = 8 * img.dtype.itemsize
    rows = np.empty((h, 1 + w * img.dtype.itemsize), dtype=np.uint8)
    rows[:, 0] = 0
    rows[:, 1:] = img.astype(img.dtype.newbyteorder(">"), copy=False).view(np.uint8).reshape(h, -1)

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)

    data = _PNG_SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, 0, 0, 0, 0))
    data += chunk(b"IDAT", zlib.compress(rows.tobytes(), int(compression))) + chunk(b"IEND", b"")
    with open(path, "wb") as fp:
        fp.write(data)


def raw_imread(path):
    """``.raw``: two uint32 (width, height) then uint16 pixels; endianness by the smaller width
    (reference readers.py:34-61; like there the array is shaped ``(width, height)``)."""
    be = np.memmap(path, dtype=">u4", mode="r", shape=(2,))
    width_be, height_be = int(be[0]), int(be[1])
    del be
    le = np.memmap(path, dtype="<u4", mode="r", shape=(2,))
    width_le, height_le = int(le[0]), int(le[1])
    del le
    if width_le < width_be:
        width, height, dtype = width_le, height_le, "<u2"
    else:
        width, height, dtype = width_be, height_be, ">u2"
    return np.memmap(path, dtype=dtype, mode="r", offset=8, shape=(width, height))


def imread(path: PathLike) -> np.ndarray:
    """Load a tiff, raw or png image (reference readers.py:64-89)."""
    path = str(path)
    ext = _get_extension(path)
    if ext == ".raw":
        return raw_imread(path)
    if ext in (".tif", ".tiff"):
        return _tifffile.imread(path) if _tifffile is not None else _tiff_read(path)
    if ext == ".png":
        return _iio.imread(path) if _iio is not None else _png_read(path)
    return None


def imsave(path, img, compression=1, output_format: Optional[str] = None):
    """Save as TIFF (default for every input format) or PNG (reference destriper.py:49-110)."""
    ext = _get_extension(path)
    if output_format is None:
        if ext not in (".raw", ".png", ".tif", ".tiff"):
            raise NotImplementedError(f"We can't save in {ext} format, available: {SUPPORTED_OUTPUT_EXTENSIONS}")
        target, kind = os.path.splitext(path)[0] + ".tiff", ".tiff"
    else:
        if output_format not in SUPPORTED_OUTPUT_EXTENSIONS:
            raise ValueError(
                f"Output format {output_format} is not valid! Supported extensions are: {SUPPORTED_OUTPUT_EXTENSIONS}"
            )
        target, kind = os.path.splitext(path)[0] + output_format, output_format
    if kind in (".tif", ".tiff"):
        if _tifffile is not None:  # pragma: no cover
            _tifffile.imwrite(target, img, compressionargs={"level": compression})
        else:
            _tiff_write(target, img)
    else:
        if _iio is not None:  # pragma: no cover
            _iio.v3.imwrite(target, img, compress_level=compression)
        else:
            _png_write(target, img, compression)


def _find_all_images(search_path: PathLike, input_path: PathLike, output_path: PathLike) -> List[Path]:
    """All supported images below ``search_path``; mirrors the directory tree under
    ``output_path`` (reference destriper.py:230-264)."""
    input_path, output_path, search_path = Path(input_path), Path(output_path), Path(search_path)
    assert search_path.is_dir()
    found = []
    for p in sorted(search_path.iterdir()):
        if p.is_file():
            if p.suffix in SUPPORTED_READING_EXTENSIONS:
                found.append(p)
        elif p.is_dir():
            o = output_path.joinpath(p.relative_to(input_path))
            if not o.exists():
                o.mkdir(parents=True)
            found.extend(_find_all_images(p, input_path, output_path))
    return found


def _read_with_retries(output_dir, input_path, retries=3):
    """Three read attempts, then the path is appended to ``destripe_log.txt`` (reference :166-192)."""
    for i in range(retries):
        try:
            img = imread(input_path)
            if img is None:
                raise OSError(f"unsupported image {input_path}")
            return img
        except Exception:
            if i == retries - 1:
                log = os.path.join(output_dir, "destripe_log.txt")
                if not os.path.exists(log):
                    with open(log, "w") as fp:
                        fp.write("Error reading the following images.  We will interpolate their content.")
                with open(log, "a+") as fp:
                    fp.write("\n{}".format(str(input_path)))
                return None
            time.sleep(0.05)


def _cast_output(filtered: np.ndarray, dtype) -> np.ndarray:
    dtype = np.dtype(dtype)
    if dtype.kind in "ui":  # the reference's astype wraps out-of-range values; saturate instead
        info = np.iinfo(dtype)
        filtered = np.clip(filtered, info.min, info.max)
    return filtered.astype(dtype)


def _save_with_retries(output_path, img, compression, output_format, retries=10):
    for _ in range(retries):  # reference :202-215 (OSError on NAS)
        try:
            imsave(output_path, img, compression=compression, output_format=output_format)
            return
        except OSError:
            logger.error(f"Retrying writing image in {output_path}...")
            time.sleep(0.05)


def read_filter_save(
    output_dir: PathLike,
    input_path: PathLike,
    output_path: PathLike,
    high_int_filter_params: dict,
    low_int_filter_params: dict,
    shadow_correction: dict,
    compression: Optional[int] = 1,
    output_format: Optional[str] = None,
    output_dtype: Optional[type] = None,
):
    """Read one image, destripe it, save it (reference destriper.py:113-227)."""
    raw_image = _read_with_retries(output_dir, input_path)
    if raw_image is None:
        return
    dtype = raw_image.dtype
    if output_dtype is not None and isinstance(output_dtype, type):
        dtype = output_dtype
    filtered = fl.filter_stripes(
        image=np.asarray(raw_image),
        input_tile_path=input_path,
        no_cells_config=low_int_filter_params,
        cells_config=high_int_filter_params,
        shadow_correction=shadow_correction,
    )
    _save_with_retries(output_path, _cast_output(filtered, dtype), compression, output_format)


def batch_filter(
    input_path: PathLike,
    output_path: PathLike,
    workers: int,
    chunks: int,
    high_int_filt_params: dict,
    low_int_filt_params: dict,
    shadow_correction: dict,
    compression: Optional[int] = 1,
    output_format: Optional[str] = None,
    output_dtype: Optional[type] = None,
):
    """Destripe every image below ``input_path`` into the same tree under ``output_path``
    (reference destriper.py:267-378).  ``workers`` = I/O threads, ``chunks`` = planes per GPU call."""
    input_path, output_path = Path(input_path), Path(output_path)
    output_path.mkdir(parents=True, exist_ok=True)
    error_path = os.path.join(output_path, "destripe_log.txt")
    if os.path.exists(error_path):
        os.remove(error_path)
    img_paths = _find_all_images(input_path, input_path, output_path)
    logger.info(f"Found {len(img_paths)} compatible images")
    for file in input_path.iterdir():
        if Path(file).suffix in [".txt", ".ini"]:
            shutil.copyfile(file, os.path.join(output_path, os.path.split(file)[1]))
    chunks = max(1, int(chunks))
    per_tile_flat = isinstance(shadow_correction, dict) and not shadow_correction.get("retrospective")

    def side_key(p):  # planes sharing one engine call must resolve to the same flat field
        return tuple(str(p).split("_")[:2]) if per_tile_flat else None

    with ThreadPoolExecutor(max(1, int(workers))) as pool:
        for start in range(0, len(img_paths), chunks):
            batch = img_paths[start : start + chunks]
            images = list(pool.map(lambda p: _read_with_retries(output_path, p), batch))
            groups = {}
            for p, img in zip(batch, images):
                if img is not None:
                    groups.setdefault((img.shape, img.dtype.str, side_key(p)), []).append((p, img))
            writes = []
            for items in groups.values():
                stack = np.stack([np.asarray(img) for _, img in items])
                tile = items[0][0]
                filtered = fl.filter_planes(
                    stack, tile, low_int_filt_params, high_int_filt_params, shadow_correction, 2700
                )
                for (p, img), res in zip(items, filtered):
                    dtype = output_dtype if (output_dtype is not None and isinstance(output_dtype, type)) else img.dtype
                    o = output_path.joinpath(p.relative_to(input_path))
                    if not o.parent.exists():
                        o.parent.mkdir(parents=True)
                    writes.append((o, _cast_output(res, dtype)))
            list(pool.map(lambda w: _save_with_retries(w[0], w[1], compression, output_format), writes))
    if os.path.exists(error_path):
        logger.error("An error happened, see destripe log for more details")
