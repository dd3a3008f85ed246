"""Blosc1 frame codec (zstd / lz4, byte shuffle) over the system ``libzstd.so.1`` / ``liblz4.so.1``.

The reference writes its output chunks with ``Blosc(cname="zstd", clevel=3, shuffle=Blosc.SHUFFLE,
blocksize=0)`` (``/root/reference/code/aind_smartspim_destripe/zarr_destriper.py:1066-1074``) and its
input tiles carry the same codec; ``numcodecs`` / ``python-blosc`` are not installable here, so the
container format is restated from the published c-blosc 1.x format (``README_HEADER.rst`` /
``blosc.c``) and the entropy coders come from the shared libraries through ``ctypes`` (the calls
release the GIL, so chunks are coded concurrently on a thread pool).

Frame layout (little endian):

    byte 0  version (2)      byte 1  versionlz (1)     byte 2  flags        byte 3  typesize
    4..7    nbytes           8..11   blocksize         12..15  cbytes (whole frame)
    flags: 0x01 byte shuffle, 0x02 memcpyed (raw copy follows the header), 0x04 bit shuffle,
           0x10 blocks are NOT split into per-byte streams, bits 5..7 compressor (1 lz4, 4 zstd)
    then   int32 bstarts[nblocks]  (offsets of the blocks from the start of the frame)
    block: for each stream (1 if not split, else typesize): int32 csize, then csize bytes; a
           stream whose csize equals its raw length is stored verbatim

The shuffle is applied per block: byte j of element i goes to position j * n_elements + i; trailing
bytes that do not fill an element are copied unchanged.  zstd level = 2 * clevel - 1 (clevel < 9),
as in c-blosc's ``zstd_wrap_compress``.
"""

from __future__ import annotations

import ctypes as C
import ctypes.util
import struct
import threading
from typing import Optional

import numpy as np

BLOSC_VERSION_FORMAT = 2
BLOSC_MIN_HEADER = 16
BLOSC_MIN_BUFFERSIZE = 128  # smaller buffers are stored raw
FLAG_SHUFFLE, FLAG_MEMCPYED, FLAG_BITSHUFFLE, FLAG_DONT_SPLIT = 0x01, 0x02, 0x04, 0x10
COMPRESSOR_CODE = {"lz4": 1, "zstd": 4}
COMPRESSOR_NAME = {v: k for k, v in COMPRESSOR_CODE.items()}


class BloscError(RuntimeError):
    pass


_libs = {}
_lib_lock = threading.Lock()


def _load(name: str):
    with _lib_lock:
        if name in _libs:
            return _libs[name]
        path = ctypes.util.find_library(name) or {"zstd": "libzstd.so.1", "lz4": "liblz4.so.1"}[name]
        try:
            lib = C.CDLL(path)
        except OSError as exc:  # pragma: no cover - both libraries are part of the image
            raise BloscError(f"blosc codec needs lib{name} ({path}): {exc}") from exc
        if name == "zstd":
            lib.ZSTD_compressBound.restype = C.c_size_t
            lib.ZSTD_compressBound.argtypes = [C.c_size_t]
            lib.ZSTD_compress.restype = C.c_size_t
            lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
            lib.ZSTD_decompress.restype = C.c_size_t
            lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            lib.ZSTD_isError.restype = C.c_uint
            lib.ZSTD_isError.argtypes = [C.c_size_t]
            lib.ZSTD_maxCLevel.restype = C.c_int
        else:
            lib.LZ4_compressBound.restype = C.c_int
            lib.LZ4_compressBound.argtypes = [C.c_int]
            lib.LZ4_compress_default.restype = C.c_int
            lib.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            lib.LZ4_decompress_safe.restype = C.c_int
            lib.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _libs[name] = lib
        return lib


def available(cname: str = "zstd") -> bool:
    try:
        _load(cname)
        return True
    except (BloscError, KeyError):
        return False


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


_WIDE = {2: np.dtype("<u2"), 4: np.dtype("<u4"), 8: np.dtype("<u8")}


def _shuffle(block: np.ndarray, typesize: int) -> np.ndarray:
    """Byte shuffle of one block (uint8 array): byte j of element i -> position j * n + i."""
    n = block.size // typesize
    if typesize <= 1 or n == 0:
        return block
    out = np.empty_like(block)
    body = block[: n * typesize]
    wide = _WIDE.get(typesize)
    if wide is not None and body.ctypes.data % typesize == 0:
        # whole elements as integers: one vectorised shift-and-narrow pass per byte plane (a strided
        # uint8 transpose is several times slower in numpy)
        v = body.view(wide)
        for j in range(typesize):
            out[j * n : (j + 1) * n] = (v >> np.array(8 * j, dtype=wide)).astype(np.uint8)
    else:
        out[: n * typesize] = body.reshape(n, typesize).T.reshape(-1)
    out[n * typesize :] = block[n * typesize :]
    return out


def _unshuffle(block: np.ndarray, typesize: int) -> np.ndarray:
    n = block.size // typesize
    if typesize <= 1 or n == 0:
        return block
    out = np.empty_like(block)
    wide = _WIDE.get(typesize)
    if wide is not None and out.ctypes.data % typesize == 0:
        v = out[: n * typesize].view(wide)
        v[:] = block[:n]
        for j in range(1, typesize):
            v |= block[j * n : (j + 1) * n].astype(wide) << np.array(8 * j, dtype=wide)
    else:
        out[: n * typesize] = block[: n * typesize].reshape(typesize, n).T.reshape(-1)
    out[n * typesize :] = block[n * typesize :]
    return out


def _auto_blocksize(nbytes: int, typesize: int, clevel: int) -> int:
    """A block size in the spirit of c-blosc's compute_blocksize for the high-ratio codecs (any
    value is valid: it is recorded in the header).  256 KB at the reference's clevel 3."""
    bs = {0: 32, 1: 64, 2: 128, 3: 256, 4: 256, 5: 512, 6: 512, 7: 1024, 8: 1024, 9: 2048}[max(0, min(9, clevel))] * 1024
    bs = min(bs, nbytes)
    if bs > typesize:
        bs -= bs % typesize
    return max(bs, 1)


def _zstd_level(clevel: int, lib) -> int:
    if clevel >= 9:
        return int(lib.ZSTD_maxCLevel())
    if clevel == 8:
        return int(lib.ZSTD_maxCLevel()) - 2
    return max(1, 2 * clevel - 1)


def _native():
    """The C++ implementation of the same format inside libdstr_b200.so (csrc/dstr_blosc.cpp), if built."""
    try:
        from . import engine

        lib = engine.load_library()
        return lib if hasattr(lib, "dstr_blosc_compress") else None
    except Exception:  # library not built: the pure-Python path below is complete
        return None


def compress(data, typesize: int = 2, clevel: int = 3, shuffle: int = 1, cname: str = "zstd", blocksize: int = 0,
             native: bool = True):
    """Blosc1 frame of ``data`` (bytes-like).  ``shuffle``: 0 none, 1 byte shuffle (bit shuffle is not written).
    ``native``: use the C++ implementation in the engine library when it is available (the whole call then runs
    without the GIL); the pure-Python path writes the same format."""
    lib = _native() if native else None
    if lib is not None and cname in COMPRESSOR_CODE and shuffle in (0, 1) and lib.dstr_blosc_available(COMPRESSOR_CODE[cname]):
        src = np.frombuffer(data, dtype=np.uint8)
        dst = np.empty(src.size + BLOSC_MIN_HEADER, dtype=np.uint8)
        r = lib.dstr_blosc_compress(src.ctypes.data if src.size else None, src.size, int(typesize), int(clevel), int(shuffle),
                                    COMPRESSOR_CODE[cname], int(blocksize), dst.ctypes.data, dst.size)
        if r < 0:
            raise BloscError(f"dstr_blosc_compress failed ({r})")
        return dst[:r].data
    return _compress_py(data, typesize, clevel, shuffle, cname, blocksize)


def _compress_py(data, typesize: int = 2, clevel: int = 3, shuffle: int = 1, cname: str = "zstd", blocksize: int = 0) -> bytes:
    if cname not in COMPRESSOR_CODE:
        raise NotImplementedError(f"blosc compressor {cname!r} (supported: zstd, lz4)")
    if shuffle not in (0, 1):
        raise NotImplementedError("blosc bit shuffle is not implemented for writing")
    src = np.frombuffer(data, dtype=np.uint8)
    nbytes = int(src.size)
    if nbytes > 0x7FFFFFFF - BLOSC_MIN_HEADER:
        raise BloscError("buffer too large for a Blosc1 frame")
    typesize = int(typesize) if 1 <= int(typesize) <= 255 else 1
    flags = (FLAG_SHUFFLE if (shuffle and typesize > 1) else 0) | FLAG_DONT_SPLIT | (COMPRESSOR_CODE[cname] << 5)
    bs = int(blocksize) if blocksize else _auto_blocksize(nbytes, typesize, clevel)
    bs = max(1, min(bs, max(nbytes, 1)))

    def memcpyed() -> bytes:
        hdr = struct.pack("<BBBBIII", BLOSC_VERSION_FORMAT, 1, (flags | FLAG_MEMCPYED), typesize, nbytes, bs,
                          nbytes + BLOSC_MIN_HEADER)
        return hdr + src.tobytes()

    if clevel == 0 or nbytes < BLOSC_MIN_BUFFERSIZE:
        return memcpyed()
    lib = _load(cname)
    nblocks = (nbytes + bs - 1) // bs
    parts = []
    bstarts = []
    pos = BLOSC_MIN_HEADER + 4 * nblocks
    if cname == "zstd":
        level = _zstd_level(clevel, lib)
        bound = int(lib.ZSTD_compressBound(bs))
    else:
        bound = int(lib.LZ4_compressBound(bs))
    dst = np.empty(bound, dtype=np.uint8)
    for b in range(nblocks):
        blk = src[b * bs : min((b + 1) * bs, nbytes)]
        if flags & FLAG_SHUFFLE:
            blk = _shuffle(blk, typesize)
        blk = np.ascontiguousarray(blk)
        if cname == "zstd":
            c = int(lib.ZSTD_compress(_ptr(dst), bound, _ptr(blk), blk.size, level))
            if lib.ZSTD_isError(c):
                raise BloscError("ZSTD_compress failed")
        else:
            c = int(lib.LZ4_compress_default(_ptr(blk), _ptr(dst), blk.size, bound))
            if c <= 0:
                raise BloscError("LZ4_compress_default failed")
        if c >= blk.size:  # incompressible stream: stored verbatim, csize == raw size
            payload = blk.tobytes()
        else:
            payload = dst[:c].tobytes()
        bstarts.append(pos)
        parts.append(struct.pack("<i", len(payload)))
        parts.append(payload)
        pos += 4 + len(payload)
    if pos >= nbytes + BLOSC_MIN_HEADER:  # no gain: c-blosc falls back to a plain copy
        return memcpyed()
    hdr = struct.pack("<BBBBIII", BLOSC_VERSION_FORMAT, 1, flags, typesize, nbytes, bs, pos)
    return b"".join([hdr, struct.pack(f"<{nblocks}i", *bstarts)] + parts)


def frame_info(frame) -> dict:
    buf = memoryview(frame)
    if len(buf) < BLOSC_MIN_HEADER:
        raise BloscError("truncated Blosc frame")
    version, versionlz, flags, typesize, nbytes, blocksize, cbytes = struct.unpack_from("<BBBBIII", buf, 0)
    return dict(version=version, versionlz=versionlz, flags=flags, typesize=typesize, nbytes=nbytes,
                blocksize=blocksize, cbytes=cbytes, shuffle=bool(flags & FLAG_SHUFFLE),
                bitshuffle=bool(flags & FLAG_BITSHUFFLE), memcpyed=bool(flags & FLAG_MEMCPYED),
                split=not (flags & FLAG_DONT_SPLIT), cname=COMPRESSOR_NAME.get(flags >> 5))


def decompress(frame, out: Optional[np.ndarray] = None, native: bool = True):
    """Decode a Blosc1 frame (zstd or lz4 streams, split or unsplit blocks, byte shuffle, memcpyed)."""
    lib = _native() if native else None
    if lib is not None:
        info = frame_info(frame)
        if not info["bitshuffle"] and (info["memcpyed"] or info["cname"] in COMPRESSOR_CODE):
            src = np.frombuffer(frame, dtype=np.uint8)
            res = np.empty(info["nbytes"], dtype=np.uint8) if out is None else out.reshape(-1).view(np.uint8)
            if res.size != info["nbytes"]:
                raise BloscError("output buffer size does not match the frame")
            r = lib.dstr_blosc_decompress(src.ctypes.data, src.size, res.ctypes.data if res.size else None, res.size)
            if info["nbytes"] and r != info["nbytes"]:
                raise BloscError(f"corrupt Blosc frame (dstr_blosc_decompress returned {r})")
            return res.data if out is None else out
    return _decompress_py(frame, out)


def _decompress_py(frame, out: Optional[np.ndarray] = None) -> bytes:
    info = frame_info(frame)
    src = np.frombuffer(frame, dtype=np.uint8)
    nbytes, bs, typesize, flags = info["nbytes"], info["blocksize"], info["typesize"], info["flags"]
    if info["version"] != BLOSC_VERSION_FORMAT:
        raise BloscError(f"unsupported Blosc format version {info['version']}")
    if info["cbytes"] > src.size:
        raise BloscError("truncated Blosc frame")
    if info["bitshuffle"]:
        raise NotImplementedError("blosc bit shuffle is not implemented")
    res = np.empty(nbytes, dtype=np.uint8) if out is None else out.reshape(-1).view(np.uint8)
    if res.size != nbytes:
        raise BloscError("output buffer size does not match the frame")
    if info["memcpyed"]:
        res[:] = src[BLOSC_MIN_HEADER : BLOSC_MIN_HEADER + nbytes]
        return res.tobytes() if out is None else out
    if nbytes == 0:
        return b"" if out is None else out
    cname = info["cname"]
    if cname not in COMPRESSOR_CODE:
        raise NotImplementedError(f"blosc compressor code {flags >> 5} (supported: zstd, lz4)")
    lib = _load(cname)
    nblocks = (nbytes + bs - 1) // bs
    bstarts = struct.unpack_from(f"<{nblocks}i", src, BLOSC_MIN_HEADER)
    tmp = np.empty(bs, dtype=np.uint8)
    for b in range(nblocks):
        blen = min(bs, nbytes - b * bs)
        # c-blosc splits a block into `typesize` streams only when the block is a whole multiple of the typesize
        # (the leftover block of a buffer is never split)
        split = info["split"] and typesize > 1 and (blen % typesize == 0) and not (b == nblocks - 1 and blen != bs)
        nstreams = typesize if split else 1
        slen = blen // nstreams
        pos = bstarts[b]
        target = tmp[:blen] if (flags & FLAG_SHUFFLE and typesize > 1) else res[b * bs : b * bs + blen]
        for s in range(nstreams):
            (csize,) = struct.unpack_from("<i", src, pos)
            pos += 4
            dst = target[s * slen : (s + 1) * slen]
            if csize == slen:
                dst[:] = src[pos : pos + csize]
            else:
                if cname == "zstd":
                    r = int(lib.ZSTD_decompress(_ptr(dst), slen, _ptr(src) + pos, csize))
                    bad = bool(lib.ZSTD_isError(r)) or r != slen
                else:
                    r = int(lib.LZ4_decompress_safe(_ptr(src) + pos, _ptr(dst), csize, slen))
                    bad = r != slen
                if bad:
                    raise BloscError(f"corrupt Blosc stream (block {b}, stream {s})")
            pos += csize
        if flags & FLAG_SHUFFLE and typesize > 1:
            res[b * bs : b * bs + blen] = _unshuffle(tmp[:blen], typesize)
    return res.tobytes() if out is None else out
