"""CPU tests of the minimal Zarr v2 store and the OME-NGFF metadata of the tile driver
(reference zarr_destriper.py:1027-1074, :410-674)."""
import json

import numpy as np
import pytest

from aind_smartspim_destripe_b200 import zarr_destriper as zd
from aind_smartspim_destripe_b200 import zarr_store as zs


@pytest.mark.parametrize("compressor", [None, {"id": "zlib", "level": 1}, {"id": "bz2", "level": 1}])
@pytest.mark.parametrize("sep", ["/", "."])
def test_round_trip_ragged_chunks(tmp_path, compressor, sep):
    rng = np.random.default_rng(0)
    data = rng.integers(0, 60000, (1, 1, 70, 150, 130), dtype=np.uint16)
    arr = zs.ZarrArray.create(tmp_path / "a", data.shape, (1, 1, 64, 128, 128), np.uint16, compressor, sep, threads=4)
    arr[...] = data
    meta = json.loads((tmp_path / "a" / ".zarray").read_text())
    assert meta["chunks"] == [1, 1, 64, 128, 128] and meta["dtype"] == "<u2" and meta["zarr_format"] == 2
    assert meta["dimension_separator"] == sep and meta["order"] == "C"
    first = tmp_path / "a" / ("0/0/0/0/0" if sep == "/" else "0.0.0.0.0")
    assert first.is_file()
    if compressor is None:
        assert first.stat().st_size == 64 * 128 * 128 * 2  # edge chunks are stored full size too
    back = zs.ZarrArray.open(tmp_path / "a")
    np.testing.assert_array_equal(back[...], data)
    np.testing.assert_array_equal(back[0, 0, 60:68, 120:140, 5], data[0, 0, 60:68, 120:140, 5])
    np.testing.assert_array_equal(back[0, 0, -1], data[0, 0, -1])
    with pytest.raises(PermissionError):
        back[0, 0, 0] = 1


def test_partial_writes_merge_and_missing_chunks_read_fill(tmp_path):
    arr = zs.ZarrArray.create(tmp_path / "b", (10, 20), (4, 8), np.float32, None, "/", fill_value=0)
    assert not list((tmp_path / "b").glob("[0-9]*"))
    np.testing.assert_array_equal(arr[...], np.zeros((10, 20), np.float32))
    arr[1:3, 2:5] = 7  # read-modify-write of one chunk, implicit cast
    arr[3:9, 6:18] = np.arange(6 * 12, dtype=np.int64).reshape(6, 12)
    ref = np.zeros((10, 20), np.float32)
    ref[1:3, 2:5] = 7
    ref[3:9, 6:18] = np.arange(72).reshape(6, 12)
    np.testing.assert_array_equal(zs.ZarrArray.open(tmp_path / "b")[...], ref)


def test_unknown_codec_is_rejected_by_name(tmp_path):
    # blosc with zstd / lz4 and byte shuffle is built in (blosc1.py); other variants need numcodecs
    with pytest.raises(NotImplementedError, match="blosclz"):
        zs.ZarrArray.create(tmp_path / "c", (4,), (4,), np.uint16, {"id": "blosc", "cname": "blosclz", "clevel": 3})
    with pytest.raises(NotImplementedError, match="numcodecs"):
        zs.ZarrArray.create(tmp_path / "c2", (4,), (4,), np.uint16, {"id": "blosc", "cname": "zstd", "shuffle": 2})
    assert zs.default_compressor()["id"] == "blosc"
    with pytest.raises(NotImplementedError):
        zs._norm_key((slice(0, 4, 2),), (4,))


def test_ome_ngff_metadata_matches_reference_layout():
    md = zd.ome_ngff_metadata((1, 1, 2000, 1600, 2000), (1, 1, 64, 128, 128), "tile_x_0000_y_0000.zarr", 3,
                              [2, 2, 2], [2.0, 1.8, 1.8])
    ms = md["multiscales"][0]
    assert [a["name"] for a in ms["axes"]] == ["t", "c", "z", "y", "x"] and ms["version"] == "0.4"
    assert [d["path"] for d in ms["datasets"]] == ["0", "1", "2"]
    assert ms["datasets"][0]["coordinateTransformations"] == [{"type": "scale", "scale": [1.0, 1.0, 2.0, 1.8, 1.8]}]
    assert ms["datasets"][2]["coordinateTransformations"][0]["scale"] == [1.0, 1.0, 8.0, 7.2, 7.2]
    om = md["omero"]
    assert om["channels"][0]["color"] == "690afe" and om["channels"][0]["window"] == {
        "end": 350.0, "max": 65535.0, "min": 0.0, "start": 0.0}
    assert om["rdefs"]["defaultZ"] == 1000
    _, chunk_sizes = zd._compute_scales(3, [2, 2, 2], [2.0, 1.8, 1.8], (1, 1, 64, 128, 128), (1, 1, 100, 200, 300))
    assert chunk_sizes == [(1, 1, 64, 128, 128), (1, 1, 50, 100, 128), (1, 1, 25, 50, 75)]


def test_get_microscope_flats_reads_sides(tmp_path):
    from aind_smartspim_destripe_b200 import destriper as D

    (tmp_path / "metadata.json").write_text(json.dumps({"tile_config": {
        "t0": {"Laser": "488", "X": "1000", "Y": "2000", "Side": "0"},
        "t1": {"Laser": "488", "X": "1000", "Y": "3000", "Side": "1"},
        "t2": {"Laser": "561", "X": "1000", "Y": "2000", "Side": "1"}}}))
    for side in (0, 1):
        D._tiff_write(str(tmp_path / f"FlatReal488_{side}.tif"), np.full((4, 5), 100 + side, np.uint16))
    flats, sides = zd.get_microscope_flats("Ex_488_Em_525", tmp_path)
    assert sides == {"1000": {"2000": 0, "3000": 1}}
    assert len(flats) == 2 and flats[0][0, 0] == 100 and flats[1][0, 0] == 101
    assert zd.get_microscope_flats("Ex_none", tmp_path) == (None, None)
    with pytest.raises(ValueError):
        zd.get_microscope_flats("Ex_561_Em_600", tmp_path)  # no flats for that laser
