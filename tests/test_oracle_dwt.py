"""Pins oracle/dwt.py (restatement of PyWavelets 1.6.0) with PyWavelets' documented known
answers, perfect reconstruction and the level-shape table of SURVEY.md Appendix A.4."""
import numpy as np
import pytest

from oracle import dwt


def test_db1_documented_answer():
    # PyWavelets docs: pywt.dwt([1, 2, 3, 4], 'db1')
    cA, cD = dwt.dwt_axis(np.array([1.0, 2, 3, 4]), "db1", -1)
    np.testing.assert_allclose(cA, [2.12132034, 4.94974747], atol=1e-8)
    np.testing.assert_allclose(cD, [-0.70710678, -0.70710678], atol=1e-8)


def test_db2_symmetric_documented_answer():
    # PyWavelets docs (DWT and IDWT page): pywt.dwt([3, 7, 1, 1, -2, 5, 4, 6], 'db2'), mode symmetric
    x = np.array([3.0, 7, 1, 1, -2, 5, 4, 6])
    cA, cD = dwt.dwt_axis(x, "db2", -1)
    np.testing.assert_allclose(cA, [5.65685425, 7.39923721, 0.22414387, 3.33677403, 7.77817459], atol=1e-8)
    np.testing.assert_allclose(cD, [-2.44948974, -1.60368225, -4.44140056, -0.41361256, 1.22474487], atol=1e-8)
    np.testing.assert_allclose(dwt.idwt_axis(cA, cD, "db2", -1), x, atol=1e-12)


def test_wavedec_db1_level2_documented_answer():
    # PyWavelets docs: pywt.wavedec([1..8], 'db1', level=2) -> [5, 13], [-2, -2], [-0.7071]*4
    a1, d1 = dwt.dwt_axis(np.arange(1.0, 9.0), "db1", -1)
    a2, d2 = dwt.dwt_axis(a1, "db1", -1)
    np.testing.assert_allclose(a2, [5.0, 13.0], atol=1e-12)
    np.testing.assert_allclose(d2, [-2.0, -2.0], atol=1e-12)
    np.testing.assert_allclose(d1, [-0.70710678] * 4, atol=1e-8)


def test_db3_filter_bank_properties():
    lo, hi, rlo, rhi = dwt.filter_bank("db3")
    assert abs(lo.sum() - np.sqrt(2)) < 1e-10 and abs(hi.sum()) < 1e-10
    assert abs((lo * lo).sum() - 1) < 1e-10 and abs((lo * hi).sum()) < 1e-10
    np.testing.assert_allclose(hi, [-0.3326705529509569, 0.8068915093133388, -0.4598775021193313,
                                    -0.13501102001039084, 0.08544127388224149, 0.035226291882100656])
    np.testing.assert_allclose(rlo, lo[::-1])
    np.testing.assert_allclose(rhi, hi[::-1])


def test_constant_signal():
    cA, cD = dwt.dwt_axis(np.full(40, 3.0), "db3", -1)
    np.testing.assert_allclose(cA, np.sqrt(2) * 3.0, atol=1e-12)
    np.testing.assert_allclose(cD, 0.0, atol=1e-10)


@pytest.mark.parametrize("n", [5, 6, 7, 16, 17, 403, 2000])
def test_perfect_reconstruction_1d(n):
    x = np.random.default_rng(n).standard_normal(n)
    cA, cD = dwt.dwt_axis(x, "db3", -1)
    assert cA.shape == ((n + 5) // 2,)
    r = dwt.idwt_axis(cA, cD, "db3", -1)
    assert r.shape == (2 * cA.size - 4,)
    np.testing.assert_allclose(r[:n], x, atol=1e-10)


def test_short_signal_repeated_reflection():
    # N < F - 1: symmetric extension reflects repeatedly (np.pad symmetric semantics)
    x = np.array([1.0, 2.0, 4.0])
    cA, cD = dwt.dwt_axis(x, "db3", -1)
    assert cA.shape == (4,)
    np.testing.assert_allclose(dwt.idwt_axis(cA, cD, "db3", -1)[:3], x, atol=1e-10)


@pytest.mark.parametrize("shape", [(100, 100), (403, 517), (1600, 2000), (2048, 2048)])
def test_wavedec2_shapes_and_reconstruction(shape):
    x = np.random.default_rng(1).standard_normal(shape)
    coeffs = dwt.wavedec2(x, "db3", level=None)
    L = len(coeffs) - 1
    assert L == min(int(np.floor(np.log2(n / 5.0))) for n in shape)
    h, w = shape
    for lvl in range(1, L + 1):
        h, w = (h + 5) // 2, (w + 5) // 2
        cH, cV, cD = coeffs[L - lvl + 1]
        assert cH.shape == cV.shape == cD.shape == (h, w)
    r = dwt.waverec2(coeffs, "db3")
    assert r.shape[0] in (shape[0], shape[0] + 1) and r.shape[1] in (shape[1], shape[1] + 1)
    np.testing.assert_allclose(r[: shape[0], : shape[1]], x, atol=1e-9)


def test_level_table_survey_a4():
    exp_T = [(802, 1002), (403, 503), (204, 254), (104, 129), (54, 67), (29, 36), (17, 20), (11, 12)]
    exp_P = [(1026, 1026), (515, 515), (260, 260), (132, 132), (68, 68), (36, 36), (20, 20), (12, 12)]
    for shape, exp in (((1600, 2000), exp_T), ((2048, 2048), exp_P)):
        assert dwt.dwtn_max_level(shape, "db3") == 8
        h, w = shape
        for e in exp:
            h, w = dwt.dwt_coeff_len(h, 6), dwt.dwt_coeff_len(w, 6)
            assert (h, w) == e


def test_band_orientation_cH_is_highpass_along_rows_axis():
    # an image constant along axis -1 with variation along axis -2 has cV = cD = 0, cH != 0
    x = np.tile(np.random.default_rng(0).standard_normal((64, 1)), (1, 48))
    cA, (cH, cV, cD) = dwt.dwt2(x, "db3")
    assert np.abs(cV).max() < 1e-10 and np.abs(cD).max() < 1e-10 and np.abs(cH).max() > 0.1


def test_float32_stays_float32_and_level0():
    x = np.random.default_rng(0).random((40, 40)).astype(np.float32)
    c = dwt.wavedec2(x, "db3", level=1)
    assert c[0].dtype == np.float32 and c[1][0].dtype == np.float32
    c0 = dwt.wavedec2(x, "db3", level=0)
    assert len(c0) == 1 and dwt.waverec2(c0, "db3") is not None
    # mixed dtypes promote to float64 in synthesis (filtering.py:221 receives f64 cH')
    r = dwt.waverec2([c[0], (c[1][0].astype(np.float64), c[1][1], c[1][2])], "db3")
    assert r.dtype == np.float64


def test_unknown_wavelet_and_ndim_errors():
    with pytest.raises(ValueError):
        dwt.wavedec2(np.zeros((8, 8)), "nope", level=1)
    with pytest.raises(ValueError):
        dwt.wavedec2(np.zeros(8), "db3", level=1)
