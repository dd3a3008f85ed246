"""Re-hosts the numeric known-answer tests of the reference's own suite
(/root/reference/code/tests/test_filtering.py, line ranges cited per test) against the oracle
restatement, i.e. pins every oracle helper the reference itself pins."""
import numpy as np
import pytest

from oracle import plane_filter as F


def test_sigmoid():  # test_filtering.py:16-26
    assert F.sigmoid(np.array(0)) == pytest.approx(0.5)
    assert F.sigmoid(np.array(-1)) == pytest.approx(1 / (1 + np.exp(1)))
    assert F.sigmoid(np.array(1)) == pytest.approx(1 / (1 + np.exp(-1)))
    data = np.array([-1, 0, 1])
    np.testing.assert_array_almost_equal(F.sigmoid(data), 1 / (1 + np.exp(-data)))


def test_foreground_fraction():  # :28-39
    img = np.array([10, 20, 30, 40, 50])
    z = (img - 30) / 10
    np.testing.assert_array_almost_equal(F.foreground_fraction(img, 30, 10), 1 / (1 + np.exp(-z)))


def test_get_foreground_background_mean():  # :41-70
    img = np.array([10, 20, 400, 500, 600])
    fg, bg, mask = F.get_foreground_background_mean(img, 0.3)
    np.testing.assert_array_equal(mask, [0, 0, 1, 1, 1])
    assert fg == pytest.approx(500.0) and bg == pytest.approx(15.0)


def test_fg_bg_mean_empty_and_single_class():  # :72-114
    fg, bg, mask = F.get_foreground_background_mean(np.array([]), 0.3)
    assert fg == 0.0 and bg == 0.0 and mask.size == 0
    img = np.array([10, 20, 30, 40, 50])
    fg, bg, mask = F.get_foreground_background_mean(img, 1.0)
    assert fg == 0.0 and bg == img.mean()
    np.testing.assert_array_equal(mask, np.zeros_like(img))
    img = np.array([400, 420, 430, 440, 460])
    fg, bg, mask = F.get_foreground_background_mean(img, 0.0)
    assert fg == img.mean() and bg == 0.0
    np.testing.assert_array_equal(mask, np.ones_like(img))


def test_foreground_rule_is_v_ge_384_for_all_uint16():
    v = np.arange(65536, dtype=np.uint16)
    with np.errstate(over="ignore"):
        f = F.foreground_fraction(v.astype(np.float16), 400, 20)
    np.testing.assert_array_equal(f > 0.3, v >= 384)


def test_notch():  # :116-133
    np.testing.assert_array_almost_equal(F.notch(5, 1.0), 1 - np.exp(-(np.arange(5) ** 2) / 2.0))
    assert F.notch(1, 1.0)[0] == pytest.approx(0.0)
    for bad in ((0, 1.0), (-1, 1.0), (5, -1)):
        with pytest.raises(ValueError):
            F.notch(*bad)


def test_gaussian_filter():  # :135-149
    np.testing.assert_array_almost_equal(
        F.gaussian_filter((3, 5), 1.0), np.broadcast_to(F.notch(5, 1.0), (3, 5))
    )
    np.testing.assert_array_equal(F.gaussian_filter((1, 1), 1.0), np.array([[0.0]]))


def test_log_space_fft_filtering_shape_and_sign():  # :151-169
    img = np.linspace(0, 255, 100 * 100, dtype=np.float32).reshape(100, 100)
    out = F.log_space_fft_filtering(img, wavelet="db3", level=1, sigma=64, max_threshold=4)
    assert out.shape == img.shape and np.all(out > 0)


def test_log_space_fft_filtering_small_image():  # :171-180
    img = np.random.default_rng(0).random((4, 4)).astype(np.float32)
    assert F.log_space_fft_filtering(img).shape == (4, 4)


def test_level_zero_is_x_plus_two():
    img = np.random.default_rng(0).integers(0, 4000, (16, 16)).astype(np.float32)
    np.testing.assert_allclose(F.log_space_fft_filtering(img), img + 2.0, rtol=1e-5)


def test_normalize_invert_hemisphere():  # :182-224
    imgs = [np.array([[0, 50], [100, 150]]), np.array([[200, 250], [300, 350]])]
    n = F.normalize_image(imgs)
    assert n.min() >= 1.0 and n.max() <= 2.0
    np.testing.assert_array_equal(F.invert_image(np.array([[1, 2], [3, 4]])), [[3, 2], [1, 0]])
    cfg = {"X1": {"Y1": 0, "Y2": 1}, "X2": {"Y1": 1}}
    flats = [np.zeros((2, 2)), np.ones((2, 2))]
    assert F.get_hemisphere_flatfield("X1_Y1", cfg, flats) is flats[0]
    assert F.get_hemisphere_flatfield("X1_Y2", cfg, flats) is flats[1]
    with pytest.raises(KeyError):
        F.get_hemisphere_flatfield("X3_Y1", cfg, flats)
    with pytest.raises(KeyError):
        F.get_hemisphere_flatfield("X2_Y9", cfg, flats)


def test_flatfield_correction_truncates():  # :226-240
    image_tiles = np.array([[[10, 20], [30, 40]]])
    flat = np.array([[[2, 2], [2, 2]]])
    dark = np.array([[[1, 1], [1, 1]]])
    out = F.flatfield_correction(image_tiles, flat, dark)
    assert out.dtype == np.uint16
    np.testing.assert_array_equal(out, np.array([[[4, 9], [14, 19]]], dtype=np.uint16))
    with pytest.raises(ValueError):
        F.flatfield_correction(image_tiles, flat, dark[:-1])


def test_filter_stripes_branches(production_configs):  # :242-281 (plumbing; here with the real filter)
    no_cells, cells = production_configs
    rng = np.random.default_rng(0)
    dim = rng.integers(90, 200, (64, 64)).astype(np.float32)
    bright = dim.copy()
    bright[:32] = 30000.0
    out_dim = F.filter_stripes(dim, "0_0", no_cells, cells)
    out_bright = F.filter_stripes(bright, "0_0", no_cells, cells)
    np.testing.assert_array_equal(out_dim, F.log_space_fft_filtering(dim, **no_cells))
    np.testing.assert_array_equal(out_bright, F.log_space_fft_filtering(bright, **cells))
    shadow = dict(retrospective=True, flatfield=np.full((64, 64), 2.0, np.float32),
                  darkfield=np.full((70, 70), 10, np.uint16), tile_config=None)
    out_s = F.filter_stripes(dim, "0_0", no_cells, cells, shadow_correction=shadow)
    assert out_s.dtype == np.uint16 and out_s.shape == (64, 64)
    np.testing.assert_array_equal(out_s, np.clip((np.maximum(out_dim - 10, 0)) / 2.0, 0, 65535).astype(np.uint16))


def test_windowed_mean_pyramid_oracle():
    # xarray_multiscale.windowed_mean + preserve_dtype semantics (zarr_destriper.py:399-405)
    from oracle import pyramid as OP

    a = np.arange(2 * 4 * 6, dtype=np.uint16).reshape(2, 4, 6)
    r = OP.windowed_mean(a, (2, 2, 2))
    assert r.shape == (1, 2, 3) and r.dtype == np.uint16
    assert r[0, 0, 0] == (0 + 1 + 6 + 7 + 24 + 25 + 30 + 31) // 8
    b = np.array([[[1, 2], [2, 2]], [[2, 2], [2, 2]]], dtype=np.uint16)  # mean 1.875 -> truncates to 1
    assert OP.windowed_mean(b, (2, 2, 2))[0, 0, 0] == 1
    odd = np.ones((3, 5, 7), np.uint16)
    assert OP.windowed_mean(odd, (2, 2, 2)).shape == (1, 2, 3)
    lv = OP.compute_pyramid(np.ones((1, 1, 8, 8, 8), np.uint16), 3, (1, 1, 2, 2, 2))
    assert [x.shape for x in lv] == [(1, 1, 8, 8, 8), (1, 1, 4, 4, 4), (1, 1, 2, 2, 2)]
