"""Blosc1 frame codec (aind_smartspim_destripe_b200/blosc1.py): the reference's chunk compressor
(zarr_destriper.py:1066-1074) restated from the published c-blosc 1.x container format.  No other
blosc implementation is installed here, so the frames are checked against the format description:
round trips, header fields, and hand-built frames (memcpyed, split streams, verbatim streams) that
the decoder must accept."""
import struct

import numpy as np
import pytest

from aind_smartspim_destripe_b200 import blosc1
from aind_smartspim_destripe_b200 import zarr_store as zs


def _smooth_u16(n, seed=0):
    rng = np.random.default_rng(seed)
    return (2000 + 300 * np.sin(np.arange(n) / 37.0) + rng.poisson(20, n)).astype(np.uint16)


@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("cname", ["zstd", "lz4"])
@pytest.mark.parametrize("n,typesize", [(0, 2), (50, 2), (64 * 128 * 128, 2), (1000003, 2), (262144 + 6, 4), (777, 1), (4099, 3)])
def test_round_trip_and_header(cname, n, typesize, native):
    """Two implementations of the format (pure Python here, C++ in libdstr_b200.so) write identical
    frames and decode each other's."""
    data = _smooth_u16(n, seed=n).tobytes()[: n * 2]
    frame = bytes(blosc1.compress(data, typesize=typesize, clevel=3, shuffle=1, cname=cname, native=native))
    info = blosc1.frame_info(frame)
    assert info["version"] == 2 and info["typesize"] == typesize and info["nbytes"] == len(data)
    assert info["cbytes"] == len(frame)
    assert bytes(blosc1.decompress(frame, native=native)) == data
    assert bytes(blosc1.decompress(frame, native=not native)) == data
    other = bytes(blosc1.compress(data, typesize=typesize, clevel=3, shuffle=1, cname=cname, native=not native))
    assert other == frame  # same container, same streams
    if len(data) >= 1 << 16:
        assert not info["memcpyed"] and info["cname"] == cname and info["shuffle"] == (typesize > 1)
        assert len(frame) < 0.8 * len(data)  # smooth 16-bit data compresses once the bytes are shuffled


def test_incompressible_data_falls_back_to_a_copy():
    data = np.random.default_rng(0).integers(0, 256, 300000, dtype=np.uint8).tobytes()
    for native in (False, True):
        frame = bytes(blosc1.compress(data, typesize=1, clevel=3, shuffle=0, native=native))
        info = blosc1.frame_info(frame)
        assert info["memcpyed"] and len(frame) == len(data) + 16
        assert bytes(blosc1.decompress(frame, native=native)) == data


def test_decoder_accepts_split_streams_and_verbatim_streams():
    """A frame as c-blosc writes it for the split codecs: every full block = `typesize` streams (here: the
    low-byte stream zstd-compressed, the high-byte stream stored verbatim), the leftover block unsplit."""
    lib = blosc1._load("zstd")
    typesize, bs = 2, 4096
    vals = _smooth_u16(bs // 2 * 2 + 100, seed=3)
    raw = vals.tobytes()
    nbytes = len(raw)
    nblocks = (nbytes + bs - 1) // bs
    body, bstarts = [], []
    pos = 16 + 4 * nblocks
    for b in range(nblocks):
        blk = np.frombuffer(raw[b * bs : (b + 1) * bs], dtype=np.uint8)
        sh = blosc1._shuffle(blk, typesize)
        leftover = b == nblocks - 1 and blk.size != bs
        streams = [sh] if leftover else [sh[: blk.size // 2], sh[blk.size // 2 :]]
        bstarts.append(pos)
        for k, st in enumerate(streams):
            st = np.ascontiguousarray(st)
            if k == 1:  # verbatim: csize == raw length
                payload = st.tobytes()
            else:
                dst = np.empty(int(lib.ZSTD_compressBound(st.size)), dtype=np.uint8)
                c = int(lib.ZSTD_compress(dst.ctypes.data, dst.size, st.ctypes.data, st.size, 3))
                payload = dst[:c].tobytes()
            body.append(struct.pack("<i", len(payload)) + payload)
            pos += 4 + len(payload)
    flags = blosc1.FLAG_SHUFFLE | (blosc1.COMPRESSOR_CODE["zstd"] << 5)  # split allowed (no DONT_SPLIT bit)
    frame = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, bs, pos) + struct.pack(f"<{nblocks}i", *bstarts) + b"".join(body)
    assert blosc1.frame_info(frame)["split"]
    assert bytes(blosc1.decompress(frame, native=False)) == raw
    assert bytes(blosc1.decompress(frame, native=True)) == raw


def test_corrupt_frames_raise():
    data = _smooth_u16(100000).tobytes()
    frame = bytearray(blosc1.compress(data))
    for native in (False, True):
        with pytest.raises(blosc1.BloscError):
            blosc1.decompress(bytes(frame[: len(frame) // 2]), native=native)
    frame[40] ^= 0xFF
    frame[41] ^= 0xFF
    for native in (False, True):
        with pytest.raises(blosc1.BloscError):
            blosc1.decompress(bytes(frame), native=native)


def test_zarr_array_with_the_reference_compressor(tmp_path):
    """Chunks (1, 1, 64, 128, 128) uint16 with blosc-zstd-3-shuffle, like zarr_destriper.py:1066-1074."""
    assert zs.default_compressor() == {"id": "blosc", "cname": "zstd", "clevel": 3, "shuffle": 1, "blocksize": 0}
    vol = _smooth_u16(70 * 150 * 140, seed=9).reshape(70, 150, 140)
    arr = zs.ZarrArray.create(tmp_path / "0", (1, 1) + vol.shape, (1, 1, 64, 128, 128), np.uint16, "default", "/")
    arr[0, 0] = vol
    arr.close()
    back = zs.ZarrArray.open(tmp_path / "0")
    assert back.meta["compressor"]["id"] == "blosc" and back.meta["compressor"]["cname"] == "zstd"
    np.testing.assert_array_equal(back[0, 0], vol)
    raw = (tmp_path / "0" / "0" / "0" / "0" / "0" / "0").read_bytes()
    info = blosc1.frame_info(raw)
    assert info["nbytes"] == 64 * 128 * 128 * 2 and info["typesize"] == 2 and info["shuffle"] and info["cname"] == "zstd"
