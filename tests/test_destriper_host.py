"""CPU tests of the TIFF/RAW directory front-end's host logic:
file codecs, directory walk, retry/log behaviour.  No GPU calls."""
import os
import struct

import numpy as np
import pytest

from aind_smartspim_destripe_b200 import destriper as D


def _write_raw(path, img, big_endian):
    bo = ">" if big_endian else "<"
    with open(path, "wb") as fp:
        fp.write(struct.pack(bo + "II", img.shape[0], img.shape[1]))
        fp.write(img.astype(bo + "u2").tobytes())


@pytest.mark.parametrize("big_endian", [False, True])
def test_raw_imread_both_byte_orders(tmp_path, big_endian):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 60000, (24, 40), dtype=np.uint16)
    p = tmp_path / "plane.raw"
    _write_raw(p, img, big_endian)
    got = D.imread(p)
    assert got.shape == (24, 40) and got.dtype.itemsize == 2
    np.testing.assert_array_equal(np.asarray(got), img)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_tiff_round_trip(tmp_path, dtype):
    rng = np.random.default_rng(4)
    img = (rng.random((33, 47)) * 200).astype(dtype)
    D.imsave(str(tmp_path / "a.tif"), img)
    assert (tmp_path / "a.tiff").exists()  # reference imsave always writes .tiff by default
    got = D.imread(tmp_path / "a.tiff")
    assert got.dtype == np.dtype(dtype)
    np.testing.assert_array_equal(got, img)


def test_tiff_reader_big_endian_multi_strip(tmp_path):
    img = np.arange(6 * 5, dtype=np.uint16).reshape(6, 5) * 1000
    rows_per_strip, strips = 4, [img[:4], img[4:]]
    body = b"".join(s.astype(">u2").tobytes() for s in strips)
    off0, off1 = 8, 8 + strips[0].size * 2
    ifd_off = 8 + len(body)
    arr_off = ifd_off + 2 + 12 * 9 + 4
    entries = [
        (256, 3, 1, struct.pack(">HH", 5, 0)), (257, 3, 1, struct.pack(">HH", 6, 0)),
        (258, 3, 1, struct.pack(">HH", 16, 0)), (259, 3, 1, struct.pack(">HH", 1, 0)),
        (262, 3, 1, struct.pack(">HH", 1, 0)), (273, 4, 2, struct.pack(">I", arr_off)),
        (277, 3, 1, struct.pack(">HH", 1, 0)), (278, 3, 1, struct.pack(">HH", rows_per_strip, 0)),
        (279, 4, 2, struct.pack(">I", arr_off + 8)),
    ]
    blob = b"MM" + struct.pack(">HI", 42, ifd_off) + body + struct.pack(">H", len(entries))
    for tag, typ, cnt, val in entries:
        blob += struct.pack(">HHI", tag, typ, cnt) + val
    blob += struct.pack(">I", 0) + struct.pack(">II", off0, off1) + struct.pack(">II", strips[0].size * 2, strips[1].size * 2)
    p = tmp_path / "be.tif"
    p.write_bytes(blob)
    np.testing.assert_array_equal(D.imread(p), img)


def test_imsave_rejects_unknown_formats(tmp_path):
    img = np.zeros((4, 4), np.uint16)
    with pytest.raises(ValueError):
        D.imsave(str(tmp_path / "a.tif"), img, output_format=".jpg")
    with pytest.raises(NotImplementedError):
        D.imsave(str(tmp_path / "a.bmp"), img)
    assert D.imread(tmp_path / "a.bmp") is None


def test_find_all_images_mirrors_tree(tmp_path):
    src, dst = tmp_path / "in", tmp_path / "out"
    (src / "Ex_488" / "1000" / "1000_2000").mkdir(parents=True)
    (src / "empty").mkdir()
    dst.mkdir()
    for rel in ["Ex_488/1000/1000_2000/000010.tif", "Ex_488/1000/1000_2000/000020.raw", "top.png",
                "notes.txt", "Ex_488/skip.json"]:
        (src / rel).write_bytes(b"x")
    found = D._find_all_images(src, src, dst)
    assert sorted(str(p.relative_to(src)) for p in found) == [
        "Ex_488/1000/1000_2000/000010.tif", "Ex_488/1000/1000_2000/000020.raw", "top.png"]
    assert (dst / "Ex_488" / "1000" / "1000_2000").is_dir() and (dst / "empty").is_dir()


def test_unreadable_image_goes_to_log(tmp_path):
    bad = tmp_path / "broken.tif"
    bad.write_bytes(b"not a tiff")
    assert D._read_with_retries(str(tmp_path), bad) is None
    log = (tmp_path / "destripe_log.txt").read_text()
    assert str(bad) in log and log.startswith("Error reading")
    bad2 = tmp_path / "broken2.tif"
    bad2.write_bytes(b"")
    D._read_with_retries(str(tmp_path), bad2)
    assert (tmp_path / "destripe_log.txt").read_text().count("\n") == 2


def test_cast_output_saturates():
    out = D._cast_output(np.array([-5.0, 10.7, 70000.0]), np.uint16)
    assert out.dtype == np.uint16 and list(out) == [0, 10, 65535]
    assert D._cast_output(np.array([1.5]), np.float32).dtype == np.float32


def _png_filtered(img, filters):
    """Encode a greyscale image with the given per-row PNG filter types (the spec's forward filters)."""
    import zlib

    h, w = img.shape
    bpp = img.dtype.itemsize
    raw = img.astype(img.dtype.newbyteorder(">")).view(np.uint8).reshape(h, -1).astype(np.int32)
    rows = bytearray()
    for y in range(h):
        ft = filters[y % len(filters)]
        cur = raw[y]
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        up = raw[y - 1] if y else np.zeros_like(cur)
        ul = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - up
        elif ft == 3:
            f = cur - ((left + up) >> 1)
        else:
            p = left + up - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
            f = cur - pred
        rows += bytes([ft]) + (f & 0xFF).astype(np.uint8).tobytes()

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)

    half = len(rows) // 2  # two IDAT chunks
    z = zlib.compress(bytes(rows), 6)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8 * bpp, 0, 0, 0, 0))
            + chunk(b"IDAT", z[: len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2 :]) + chunk(b"IEND", b""))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_builtin_png_codec_all_row_filters(tmp_path, dtype):
    """readers.py:64-89 reads PNG through imageio; without it the built-in codec must decode every filter type."""
    rng = np.random.default_rng(2)
    img = rng.integers(0, np.iinfo(dtype).max, (37, 29)).astype(dtype)
    img[5:20] = (np.arange(29) * 7).astype(dtype)[None, :]
    p = tmp_path / "f.png"
    p.write_bytes(_png_filtered(img, [0, 1, 2, 3, 4]))
    out = D.imread(p)
    assert out.dtype == dtype
    np.testing.assert_array_equal(out, img)
    q = tmp_path / "w.png"
    D.imsave(str(q), img, compression=3, output_format=".png")
    np.testing.assert_array_equal(D.imread(q), img)
    (tmp_path / "bad.png").write_bytes(b"not a png")
    with pytest.raises(OSError):
        D.imread(tmp_path / "bad.png")
