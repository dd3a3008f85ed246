"""GPU tests of the reference-facing API: filter_stripes dispatch + epilogue, the chunk worker,
the streamed volume pipeline and the helper entry points, against the oracle."""
import numpy as np
import pytest

from _parity import U16_FRACTION, u16_agreement
from aind_smartspim_destripe_b200 import engine as E
from aind_smartspim_destripe_b200 import filtering as fl
from aind_smartspim_destripe_b200 import synthetic as S
from aind_smartspim_destripe_b200 import zarr_destriper as zd
from oracle import plane_filter as OF
from oracle import worker as OW

pytestmark = pytest.mark.gpu


def _shadow(H, W, retrospective=True):
    flat, dark = S.synthetic_flat_dark(H + 8, W + 8)
    if retrospective:
        return dict(retrospective=True, flatfield=flat[:H, :W].copy(), darkfield=dark, tile_config=None)
    flat2 = (flat[:H, :W] * 1.1).astype(np.float32)
    return dict(retrospective=False, flatfield=[flat[:H, :W].copy(), flat2], darkfield=dark,
                tile_config={"10": {"20": 1, "30": 0}})


def test_get_foreground_background_mean_known_answers():
    # /root/reference/code/tests/test_filtering.py:41-114
    fg, bg, mask = fl.get_foreground_background_mean(np.array([10, 20, 400, 500, 600]), 0.3)
    assert fg == pytest.approx(500.0) and bg == pytest.approx(15.0)
    np.testing.assert_array_equal(mask, [0, 0, 1, 1, 1])
    fg, bg, mask = fl.get_foreground_background_mean(np.array([]), 0.3)
    assert fg == 0.0 and bg == 0.0 and mask.size == 0
    img = np.array([10, 20, 30, 40, 50])
    fg, bg, mask = fl.get_foreground_background_mean(img, 1.0)
    assert fg == 0.0 and bg == pytest.approx(img.mean())
    np.testing.assert_array_equal(mask, np.zeros_like(img))
    img = np.array([400, 420, 430, 440, 460])
    fg, bg, mask = fl.get_foreground_background_mean(img, 0.0)
    assert fg == pytest.approx(img.mean()) and bg == 0.0
    np.testing.assert_array_equal(mask, np.ones_like(img))


def test_plane_stats_match_oracle_on_planes():
    st = S.synthetic_stack(4, 200, 240, cells_every=2)
    eng = E.DestripeEngine(200, 240, max_planes=4)
    fg, bg, uc = eng.plane_stats(st, high_int=2500)
    for z in range(4):
        rfg, rbg, _ = OF.get_foreground_background_mean(st[z].astype(np.float32))
        assert fg[z] == pytest.approx(float(rfg), rel=1e-5) and bg[z] == pytest.approx(float(rbg), rel=1e-5)
        assert uc[z] == int(rfg > rbg and rfg > 2500)
    assert list(uc) == [0, 1, 0, 1]
    eng.close()


def test_flatfield_correction_known_answer():
    # /root/reference/code/tests/test_filtering.py:226-240 (truncation, not rounding)
    image_tiles = np.array([[[10, 20], [30, 40]]])
    out = fl.flatfield_correction(image_tiles, np.array([[[2, 2], [2, 2]]]), np.array([[[1, 1], [1, 1]]]))
    assert out.dtype == np.uint16
    np.testing.assert_array_equal(out, np.array([[[4, 9], [14, 19]]], dtype=np.uint16))
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 70000, (3, 40, 50))
    flat = rng.uniform(1, 2, (3, 40, 50)).astype(np.float32)
    dark = rng.integers(90, 110, (3, 40, 50)).astype(np.uint16)
    ref = OF.flatfield_correction(img.astype(np.float32), flat, dark)
    out = fl.flatfield_correction(img.astype(np.float32), flat, dark)
    frac, mx, _ = u16_agreement(out, ref)
    assert frac == 1.0 and mx <= 1


@pytest.mark.parametrize("retrospective", [True, False])
def test_filter_stripes_dispatch_and_shadow(retrospective, production_configs):
    no_cells, cells = production_configs
    H, W = 320, 400
    st = S.synthetic_stack(2, H, W, base_seed=11, cells_every=2)
    shadow = _shadow(H, W, retrospective)
    tile = "10_20"
    for z in range(2):
        img = st[z].astype(np.float32)
        ref = OF.filter_stripes(img, tile, no_cells, cells, shadow, 2500)
        out = fl.filter_stripes(img, tile, no_cells, cells, shadow, 2500)
        assert out.dtype == np.uint16 and out.shape == (H, W)
        frac, mx, exact = u16_agreement(out, ref)
        print(f"filter_stripes plane {z} retrospective={retrospective}: within+-1 {frac:.6f} exact {exact:.4f} max {mx}")
        assert frac >= U16_FRACTION
        ref_f = OF.filter_stripes(img, tile, no_cells, cells, None, 2500)
        out_f = fl.filter_stripes(img, tile, no_cells, cells, None, 2500)
        assert out_f.dtype == np.float64
        assert np.abs(out_f - ref_f).max() / ref_f.max() < 1e-4
    with pytest.raises(KeyError):
        fl.filter_stripes(st[0], "99_20", no_cells, cells, _shadow(H, W, False), 2500)
    bad = dict(shadow)
    bad["darkfield"] = np.zeros((H - 1, W), np.uint16)
    with pytest.raises(ValueError):
        fl.filter_stripes(st[0], tile, no_cells, cells, bad, 2500)


def test_filter_planes_equals_per_plane_calls_and_mixed_levels(production_configs):
    no_cells, cells = production_configs
    H, W = 256, 320
    st = S.synthetic_stack(5, H, W, base_seed=21, cells_every=2)
    shadow = _shadow(H, W)
    batch = fl.filter_planes(st, "0_0", no_cells, cells, shadow, 2500)
    for z in range(5):
        one = fl.filter_stripes(st[z], "0_0", no_cells, cells, shadow, 2500)
        np.testing.assert_array_equal(batch[z], one)
    # configs with different decomposition depth are grouped by the dispatch decision
    cells3 = dict(cells, level=3)
    mixed = fl.filter_planes(st, "0_0", no_cells, cells3, shadow, 2500)
    for z in range(5):
        ref = OF.filter_stripes(st[z].astype(np.float32), "0_0", no_cells, cells3, shadow, 2500)
        frac, mx, _ = u16_agreement(mixed[z], ref)
        assert frac >= U16_FRACTION


def test_execute_worker_matches_oracle_worker(production_configs):
    no_cells, cells = production_configs
    Z, H, W = 6, 200, 240
    st = S.synthetic_stack(Z, H, W, base_seed=31, cells_every=3).astype(np.float32)
    shadow = _shadow(H, W)
    sink = np.zeros((1, 1, 10, H, W), dtype=np.uint16)  # dataset shorter than the chunk end -> clamp
    ref_sink = np.zeros_like(sink)
    args = dict(
        batch_super_chunk=(slice(0, 12), slice(0, H), slice(0, W)),
        batch_internal_slice=(slice(6, 12), slice(0, H), slice(0, W)),
        cells_config=cells, no_cells_config=no_cells, overlap_prediction_chunksize=(0, 0, 0),
        shadow_correction=shadow, dataset_name="0_0.zarr", logger=None,
    )
    zd.execute_worker(data=st[None], output_destriped_zarr=sink, **args)
    OW.execute_worker(data=st[None], output_destriped_zarr=ref_sink, **args)
    assert np.all(sink[0, 0, :6] == 0)
    frac, mx, _ = u16_agreement(sink[0, 0, 6:10], ref_sink[0, 0, 6:10])
    print(f"execute_worker: within+-1 {frac:.6f} max {mx}")
    assert frac >= U16_FRACTION and sink[0, 0, 6:10].max() > 0


def test_destripe_volume_streams_and_matches_chunk_call(production_configs):
    no_cells, cells = production_configs
    Z, H, W = 40, 160, 192
    vol = S.synthetic_stack(Z, H, W, base_seed=41, cells_every=4, n_unique=8)
    shadow = _shadow(H, W)
    out = np.zeros((Z, H, W), np.uint16)
    t = zd.destripe_volume(vol, out, no_cells, cells, shadow, chunk_planes=16)
    assert t["planes"] == Z and t["wall_s"] > 0
    ref = fl.filter_planes(vol, "0_0", no_cells, cells, shadow, 2500)
    np.testing.assert_array_equal(out, ref)
    # Z-slab of "rank 1 of 2" only touches its own planes
    out2 = np.zeros((Z, H, W), np.uint16)
    z0, z1 = zd.z_slab(Z, 1, 2, align=16)
    zd.destripe_volume(vol, out2, no_cells, cells, shadow, chunk_planes=16, z_range=(z0, z1))
    np.testing.assert_array_equal(out2[z0:z1], ref[z0:z1])
    assert not out2[:z0].any()


def test_error_paths():
    eng = E.DestripeEngine(64, 80, max_planes=2)
    p = E.make_params(dict(level=None, sigma=8, max_threshold=3))
    with pytest.raises(ValueError):
        eng.filter_chunk(np.zeros((1, 64, 81), np.uint16), p)
    with pytest.raises(ValueError):  # shadow requested without flat/dark
        eng.filter_chunk(np.zeros((1, 64, 80), np.uint16), p, flags=E.FLAG_SHADOW)
    with pytest.raises(ValueError):
        eng.set_flat_dark(np.ones((64, 81), np.float32), np.ones((64, 80), np.float32))
    eng.close()


def test_downscale_and_fused_pyramid_match_oracle(production_configs):
    # next-row f3: xarray_multiscale windowed_mean + preserve_dtype (zarr_destriper.py:365-407)
    from oracle import pyramid as OP

    no_cells, cells = production_configs
    rng = np.random.default_rng(5)
    vol = rng.integers(0, 65536, (9, 37, 50)).astype(np.uint16)  # odd sizes are cropped
    eng = E.DestripeEngine(37, 50, max_planes=4)
    np.testing.assert_array_equal(eng.downscale2x(vol), OP.windowed_mean(vol, (2, 2, 2)))
    eng.close()
    lv = zd.compute_pyramid(vol[None, None], 3, (1, 1, 2, 2, 2))
    ref = OP.compute_pyramid(vol[None, None], 3, (1, 1, 2, 2, 2))
    assert len(lv) == 3
    for a, b in zip(lv, ref):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(NotImplementedError):
        zd.compute_pyramid(vol[None, None], 2, (1, 1, 2, 2, 4))

    # fused: pyramid levels emitted while the destriped chunk is resident == pyramid of the output
    Z, H, W = 24, 160, 192
    st = S.synthetic_stack(Z, H, W, base_seed=51, cells_every=3, n_unique=6)
    shadow = _shadow(H, W)
    out = np.zeros((Z, H, W), np.uint16)
    p1 = np.zeros((Z // 2, H // 2, W // 2), np.uint16)
    p2 = np.zeros((Z // 4, H // 4, W // 4), np.uint16)
    zd.destripe_volume(st, out, no_cells, cells, shadow, chunk_planes=8, pyramid_outputs=(p1, p2))
    ref = OP.compute_pyramid(out, 3, (2, 2, 2))
    np.testing.assert_array_equal(p1, ref[1])
    np.testing.assert_array_equal(p2, ref[2])
    np.testing.assert_array_equal(out, fl.filter_planes(st, "0_0", no_cells, cells, shadow, 2500))


@pytest.mark.parametrize("with_shadow", [False, True])
def test_batch_filter_directory_matches_oracle(tmp_path, production_configs, with_shadow):
    # /root/reference/code/aind_smartspim_destripe/destriper.py:267-378 (TIFF / RAW directory path)
    import struct

    from aind_smartspim_destripe_b200 import destriper as D

    no_cells, cells = production_configs
    H, W = 160, 192
    src, dst = tmp_path / "in", tmp_path / "out"
    tile = src / "Ex_488_Em_525" / "471320" / "471320_304840"
    tile.mkdir(parents=True)
    (src / "metadata.txt").write_text("acquisition")
    st = S.synthetic_stack(5, H, W, cells_every=2)
    small = S.synthetic_stack(1, 96, 112)[0]
    names = []
    for z in range(4):
        D._tiff_write(str(tile / f"{z:06d}.tiff"), st[z])
        names.append(tile / f"{z:06d}.tiff")
    with open(tile / "000004.raw", "wb") as fp:  # big-endian raw, like the acquisition software
        fp.write(struct.pack(">II", H, W))
        fp.write(st[4].astype(">u2").tobytes())
    names.append(tile / "000004.raw")
    D._tiff_write(str(tile / "000005.tiff"), small)  # different shape inside one batch
    names.append(tile / "000005.tiff")
    (tile / "000006.tiff").write_bytes(b"garbage")
    shadow = _shadow(H, W) if with_shadow else None
    if with_shadow:
        (tile / "000005.tiff").unlink()  # the flat field only fits the (H, W) planes
        names.pop()
    D.batch_filter(src, dst, workers=3, chunks=4, high_int_filt_params=cells, low_int_filt_params=no_cells,
                   shadow_correction=shadow)
    assert (dst / "metadata.txt").read_text() == "acquisition"
    assert str(tile / "000006.tiff") in (dst / "destripe_log.txt").read_text()
    for p in names:
        img = np.asarray(D.imread(p))
        out = D.imread((dst / p.relative_to(src)).with_suffix(".tiff"))
        # the reference feeds the file's uint16 planes (destriper.py:194): float64 flow (filtering.py:175)
        ref = OF.filter_stripes(img, str(p), no_cells, cells, shadow, 2700)
        ref = np.clip(ref, 0, 65535).astype(np.uint16)
        assert img.dtype.kind == "u" and img.dtype.itemsize == 2 and out.dtype == np.uint16 and out.shape == img.shape
        frac, worst, _ = u16_agreement(out, ref)
        assert frac >= U16_FRACTION, (p.name, frac, worst)


def _make_tile(path, vol):
    from aind_smartspim_destripe_b200 import zarr_store as zs

    zs.create_group(path)
    arr = zs.ZarrArray.create(path / "0", (1, 1) + vol.shape, (1, 1, 32, 64, 64), np.uint16, {"id": "zlib", "level": 1})
    arr[0, 0] = vol
    arr.close()


def test_destripe_zarr_tile_driver_matches_oracle(tmp_path, production_configs):
    # /root/reference/code/aind_smartspim_destripe/zarr_destriper.py:909-1211 (tile -> destriped zarr + pyramid)
    import json

    from aind_smartspim_destripe_b200 import destriper as D
    from aind_smartspim_destripe_b200 import zarr_store as zs
    from oracle import pyramid as OP

    no_cells, cells = production_configs
    Z, H, W = 72, 160, 192
    vol = S.synthetic_stack(Z, H, W, base_seed=77, cells_every=3, n_unique=6)
    tile = tmp_path / "SPIM.ome.zarr" / "Ex_488_Em_525" / "471320_304840.zarr"
    tile.parent.mkdir(parents=True)
    _make_tile(tile, vol)
    shadow = _shadow(H, W)
    deriv = tmp_path / "derivatives"
    deriv.mkdir()
    D._tiff_write(str(deriv / "DarkMaster_cropped.tif"), shadow["darkfield"])
    out = tmp_path / "results" / "Ex_488_Em_525" / tile.name
    params = {"no_cells_config": no_cells, "cells_config": cells}
    t = zd.destripe_zarr(tile, "0", out, (16, H, W), 3072, 0, 1, None, tmp_path, deriv, [1.8, 1.8, 2.0], params,
                         flatfield=shadow["flatfield"], compressor={"id": "zlib", "level": 1})
    assert t["planes"] == Z
    lv = [zs.ZarrArray.open(out / str(k)) for k in range(3)]
    assert lv[0].shape == (1, 1, Z, H, W) and lv[0].chunks == (1, 1, 64, 128, 128) and lv[0].dtype == np.uint16
    assert lv[1].shape == (1, 1, Z // 2, H // 2, W // 2) and lv[2].shape == (1, 1, Z // 4, H // 4, W // 4)
    got = lv[0][0, 0]
    for z in range(0, Z, 7):
        ref = OF.filter_stripes(vol[z].astype(np.float32), tile.name, no_cells, cells, shadow, 2500)
        frac, worst, _ = u16_agreement(got[z], ref)
        assert frac >= U16_FRACTION, (z, frac, worst)
    # the multiscale levels are the truncating 2x2x2 means of the written level 0 (bit-exact)
    pyr = OP.compute_pyramid(got, 3, [2, 2, 2])
    np.testing.assert_array_equal(lv[1][0, 0], pyr[1])
    np.testing.assert_array_equal(lv[2][0, 0], pyr[2])
    attrs = json.loads((out / ".zattrs").read_text())
    assert attrs["multiscales"][0]["datasets"][1]["coordinateTransformations"][0]["scale"] == [1.0, 1.0, 4.0, 3.6, 3.6]
    assert json.loads((out / ".zgroup").read_text()) == {"zarr_format": 2}

    # without a flat field there is no fused pyramid: the levels are built from the written level 0 in streamed
    # pieces (64 + 8 planes here), and a YX chunking smaller than the plane is announced as a deviation
    out_nf = tmp_path / "results_noflat" / "Ex_488_Em_525" / tile.name
    with pytest.warns(UserWarning, match="smaller than the plane"):
        zd.destripe_zarr(tile, "0", out_nf, (16, H // 2, W), 3072, 0, 1, None, tmp_path, tmp_path / "no_derivatives",
                         [1.8, 1.8, 2.0], params, flatfield=None, compressor=None)
    lv_nf = [zs.ZarrArray.open(out_nf / str(k)) for k in range(3)]
    pyr_nf = OP.compute_pyramid(lv_nf[0][0, 0], 3, [2, 2, 2])
    np.testing.assert_array_equal(lv_nf[1][0, 0], pyr_nf[1])
    np.testing.assert_array_equal(lv_nf[2][0, 0], pyr_nf[2])
    assert lv_nf[0][0, 0].any()

    # two "ranks" (run one after the other) write the same tile as one rank: slabs [0, 256) and [256, 320)
    Z2, H2, W2 = 320, 96, 112
    vol2 = S.synthetic_stack(Z2, H2, W2, base_seed=78, cells_every=5, n_unique=8)
    tile2 = tile.parent / "471320_330760.zarr"
    _make_tile(tile2, vol2)
    sh2 = _shadow(H2, W2)
    D._tiff_write(str(deriv / "DarkMaster_cropped.tif"), sh2["darkfield"])
    outs = [tmp_path / f"results_w{w}" / "Ex_488_Em_525" / tile2.name for w in (1, 2)]
    zd.destripe_zarr(tile2, "0", outs[0], (64, H2, W2), 3072, 2, 1, None, tmp_path, deriv, [1.8, 1.8, 2.0], params,
                     flatfield=sh2["flatfield"], compressor=None, rank=0, world_size=1)
    planes = [zd.destripe_zarr(tile2, "0", outs[1], (64, H2, W2), 3072, 2, 1, None, tmp_path, deriv, [1.8, 1.8, 2.0],
                               params, flatfield=sh2["flatfield"], compressor=None, rank=r, world_size=2)["planes"]
              for r in (0, 1)]
    assert planes == [256, 64]
    for k in range(3):
        np.testing.assert_array_equal(zs.ZarrArray.open(outs[1] / str(k))[...], zs.ZarrArray.open(outs[0] / str(k))[...])


def test_destripe_channel_picks_flat_by_laser_side(tmp_path, production_configs):
    # /root/reference/code/aind_smartspim_destripe/zarr_destriper.py:1214-1267
    from aind_smartspim_destripe_b200 import destriper as D
    from aind_smartspim_destripe_b200 import zarr_store as zs

    no_cells, cells = production_configs
    Z, H, W = 1600 // 100, 96, 112  # destripe_channel fixes prediction_chunksize to (64, 1600, 2000): Z <= 64 here
    vols = {name: S.synthetic_stack(Z, H, W, base_seed=s, n_unique=4) for name, s in [("100_200", 5), ("100_300", 9)]}
    data = tmp_path / "SPIM.ome.zarr"
    for name, vol in vols.items():
        (data / "Ex_488_Em_525").mkdir(parents=True, exist_ok=True)
        _make_tile(data / "Ex_488_Em_525" / f"{name}.zarr", vol)
    sh = _shadow(H, W)
    flats = [sh["flatfield"], (sh["flatfield"] * 1.25).astype(np.float32)]
    deriv = tmp_path / "derivatives"
    deriv.mkdir()
    D._tiff_write(str(deriv / "DarkMaster_cropped.tif"), sh["darkfield"])
    flat_paths = []
    for i, f in enumerate(flats):
        D._tiff_write(str(tmp_path / f"flat_{i}.tif"), f)
        flat_paths.append(tmp_path / f"flat_{i}.tif")
    params = {"no_cells_config": no_cells, "cells_config": cells}
    zd.destripe_channel(data, deriv, "Ex_488_Em_525", tmp_path / "results", [1.8, 1.8, 2.0], flat_paths,
                        {"0": ["100_200"], "1": ["100_300"]}, params)
    for side, name in enumerate(["100_200", "100_300"]):
        got = zs.ZarrArray.open(tmp_path / "results" / "destriped_data" / "Ex_488_Em_525" / f"{name}.zarr" / "0")[0, 0]
        shadow = dict(retrospective=True, flatfield=flats[side], darkfield=sh["darkfield"], tile_config=None)
        ref = fl.filter_planes(vols[name], name, no_cells, cells, shadow, 2500)
        np.testing.assert_array_equal(got, ref)
    with pytest.raises(ValueError):
        zd.destripe_channel(data, deriv, "Ex_488_Em_525", tmp_path / "r2", [1.8, 1.8, 2.0], flat_paths,
                            {"0": ["100_200"]}, params)


def test_asynchronous_host_chunks_match_synchronous_calls(production_configs):
    # DSTR_FLAG_NO_SYNC with pinned host buffers: chunks are queued back to back, results are valid after synchronize()
    no_cells, cells = production_configs
    H, W = 160, 192
    eng = E.DestripeEngine(H, W, max_planes=8)
    f, d = S.synthetic_flat_dark(H, W)
    eng.set_flat_dark(f, d)
    chunks = [S.synthetic_stack(10, H, W, base_seed=300 + 20 * k, cells_every=3) for k in range(3)]
    pn, pc = E.make_params(no_cells), E.make_params(cells)
    ref = [eng.filter_chunk(c, pn, cells=pc, high_int=2500, mode=E.MODE_DISPATCH, flags=E.FLAG_SHADOW) for c in chunks]
    pins = [E.PinnedBuffer(c.shape, np.uint16) for c in chunks]
    outs = [E.PinnedBuffer(c.shape, np.uint16) for c in chunks]
    for p, c in zip(pins, chunks):
        p.array[...] = c
    for p, o in zip(pins, outs):
        eng.filter_chunk(p.array, pn, cells=pc, out=o.array, high_int=2500, mode=E.MODE_DISPATCH,
                         flags=E.FLAG_SHADOW | E.FLAG_NO_SYNC)
    # another entry point drains the queue first instead of racing with it
    fg, bg, uc = eng.plane_stats(chunks[0], high_int=2500)
    eng.synchronize()
    for o, r in zip(outs, ref):
        np.testing.assert_array_equal(o.array, r)
    assert len(fg) == 10
    for b in pins + outs:
        b.free()
    eng.close()


class _FailingSink:
    """(Z, H, W) sink whose second write fails (disk full, codec error, ...)."""

    def __init__(self, shape):
        self.shape, self.calls = shape, 0

    def __setitem__(self, key, value):
        self.calls += 1
        if self.calls >= 2:
            raise OSError("sink failed")


class _FailingSource:
    def __init__(self, vol, fail_at):
        self.vol, self.shape, self.dtype, self.fail_at = vol, vol.shape, vol.dtype, fail_at

    def __getitem__(self, key):
        if key.start >= self.fail_at:
            raise OSError("source failed")
        return self.vol[key]


@pytest.mark.timeout(120)
def test_destripe_volume_surfaces_stage_failures_instead_of_hanging(production_configs):
    """A failing writer / reader / device stage stops the pipeline and re-raises (the reference's consumers
    would hang on join, zarr_destriper.py:1171).  More chunks than buffers, so a dead stage would block the rest."""
    no_cells, cells = production_configs
    vol = S.synthetic_stack(40, 96, 128, base_seed=3)
    with pytest.raises(OSError, match="sink failed"):
        zd.destripe_volume(vol, _FailingSink(vol.shape), no_cells, cells, None, chunk_planes=4, queue_depth=2)
    out = np.zeros(vol.shape, np.float32)
    with pytest.raises(OSError, match="source failed"):
        zd.destripe_volume(_FailingSource(vol, 20), out, no_cells, cells, None, chunk_planes=4, queue_depth=2)
    with pytest.raises(KeyError):  # engine-side failure: tile missing from tile_config with per-side flats
        flat = np.ones((96, 128), np.float32)
        shadow = dict(retrospective=False, flatfield=[flat, flat], darkfield=np.zeros((96, 128), np.float32), tile_config={})
        zd.destripe_volume(vol, np.zeros(vol.shape, np.uint16), no_cells, cells, shadow, dataset_name="0_0.zarr", chunk_planes=4)
    # and a healthy run after the failures still works
    good = np.zeros(vol.shape, np.float32)
    t = zd.destripe_volume(vol, good, no_cells, cells, None, chunk_planes=8)
    assert t["planes"] == 40 and np.all(good > 0)


def test_destripe_volume_writes_are_cut_on_sink_chunk_boundaries(tmp_path, production_configs):
    """io_threads > 1 with a chunked sink whose Z-chunk is deeper than a sub-range: no two writers may touch
    the same stored chunk (lost updates / torn temp files otherwise)."""
    from aind_smartspim_destripe_b200 import zarr_store as zs

    no_cells, cells = production_configs
    vol = S.synthetic_stack(48, 96, 128, base_seed=5)
    flat, dark = S.synthetic_flat_dark(96, 128)
    shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
    arr = zs.ZarrArray.create(tmp_path / "0", (1, 1) + vol.shape, (1, 1, 32, 64, 64), np.uint16, "default", "/")
    sink = zd._PlanesView(arr, 0, 0)
    assert sink.chunks == (32, 64, 64)
    zd.destripe_volume(vol, sink, no_cells, cells, shadow, chunk_planes=24, io_threads=4, z_range=(8, 48))
    ref = fl.filter_planes(vol[8:], "0_0", no_cells, cells, shadow, 2500)
    np.testing.assert_array_equal(arr[0, 0, 8:], ref)
    assert not list((tmp_path / "0").rglob("*.tmp"))


def test_debug_fetch_reports_the_dispatch_decision(production_configs):
    """DSTR_FETCH_STATS slot 4 is the per-plane cells / no_cells choice the device made (filtering.py:459-467), slots 5 / 6
    the foreground / background means it was made from."""
    no_cells, cells = production_configs
    Z, H, W = 4, 256, 320
    st = S.synthetic_stack(Z, H, W, base_seed=90, cells_every=2)
    shadow = _shadow(H, W)
    eng = E.DestripeEngine(H, W, max_planes=Z)
    eng.set_subchunk(Z)
    fl.filter_planes(st, "0_0", no_cells, cells, shadow, 2500, engine=eng)
    stats = eng.debug_fetch(E.FETCH_STATS, 1, Z)
    seen = set()
    for z in range(Z):
        fg, bg, _ = OF.get_foreground_background_mean(st[z].astype(np.float32))
        expect = bool(fg > bg and fg > 2500)
        seen.add(expect)
        assert bool(stats[z][4]) == expect, (z, stats[z][4:7], fg, bg)
        assert abs(stats[z][5] - fg) <= 1e-3 * max(1.0, abs(fg)) and abs(stats[z][6] - bg) <= 1e-3 * max(1.0, abs(bg))
    assert seen == {True, False}
    eng.close()
