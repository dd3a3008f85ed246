"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares, geometry / table helpers, parameter validation, slab partitioning, slice arithmetic.
No GPU compute is issued here."""
import os
import re

import numpy as np
import pytest
from scipy import fftpack

import aind_smartspim_destripe_b200 as pkg
from aind_smartspim_destripe_b200 import engine as E
from aind_smartspim_destripe_b200 import filtering as fl
from aind_smartspim_destripe_b200 import zarr_destriper as zd
from oracle import dwt as odwt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "dstr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dstr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = E.load_library()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dstr_b200.h but not exported"
    assert sorted(E.EXPORTED_SYMBOLS) == names


def test_max_level_and_level_shapes_match_oracle():
    for shape in [(2048, 2048), (1600, 2000), (100, 100), (403, 517), (4, 4), (5, 5), (11, 4000)]:
        assert E.max_level(*shape) == odwt.dwtn_max_level(shape, "db3")
        h, w = shape
        for lvl in range(1, 9):
            h, w = odwt.dwt_coeff_len(h, 6), odwt.dwt_coeff_len(w, 6)
            assert E.level_shape(shape[0], shape[1], lvl) == (h, w)


@pytest.mark.parametrize("n,s", [(1026, 64.125), (1002, 64.16), (503, 32.24), (67, 4.32), (20, 1.36), (12, 0.88), (13, 0.4)])
def test_notch_kernels_reproduce_packed_rfft_operator(n, s):
    # reference operator: irfft(rfft(x) * g), g = notch(n, s) on the PACKED index (filtering.py:206-215)
    hp, hq = E.notch_kernels(n, s)
    g = fl.notch(n, s)
    M = fftpack.irfft(fftpack.rfft(np.eye(n), axis=-1) * g, axis=-1).T
    t = np.arange(n)[:, None]
    v = np.arange(n)[None, :]
    B = hp[(t - v) % n] + hq[(t + v) % n]
    np.testing.assert_allclose(np.eye(n) - B, M, atol=1e-13)
    x = np.random.default_rng(0).standard_normal(n)
    np.testing.assert_allclose(x - B @ x, fftpack.irfft(fftpack.rfft(x) * g), atol=1e-12)


@pytest.mark.parametrize("n,s", [(1026, 64.125), (1026, 32.06), (1002, 64.16), (503, 32.24), (254, 16.3), (129, 8.3),
                                 (67, 4.3), (36, 2.25), (12, 0.75), (2050, 128.1)])
@pytest.mark.parametrize("eps", [1e-6, 0.0])
def test_device_notch_tables_bound_the_truncation_error(n, s, eps):
    """The even/odd FIR + rank-J tables the device uses (host evaluation of the same float32
    tables) reproduce irfft(rfft(x) * g) within the design tolerance, for hybrid and dense forms."""
    d = E.notch_design(n, s, eps)
    if eps == 0.0:
        assert d["J"] == 0 and d["ntap_e"] >= n
    g = fl.notch(n, s)
    rng = np.random.default_rng(n)
    for trial in range(3):
        x = rng.standard_normal(n) if trial < 2 else np.ones(n)
        if trial == 1:
            x = np.cumsum(x) / 10.0
        ref = x - fftpack.irfft(fftpack.rfft(x) * g)
        y = E.notch_apply_host(x, s, eps)
        assert np.abs(y - ref).max() <= (2e-6 if eps > 0 else 1e-7) * np.abs(x).max()


@pytest.mark.parametrize("n,s", [(1026, 64.125), (1026, 32.06), (1002, 64.16), (515, 32.19), (503, 32.24), (260, 16.25),
                                 (254, 16.3), (132, 8.25), (129, 8.3), (100, 50.0), (640, 3.0)])
def test_tensor_core_tables_reproduce_the_operator(n, s):
    """The fp16 hi/lo Hankel tables, descriptor addressing, banding and pre-scaling of the tcgen05
    row filter (host evaluation of exactly that data path) reproduce irfft(rfft(x) * g)."""
    info = E.notch_umma_info(n, s)
    assert info["eligible"] == 1 and info["smem_bytes"] <= 226 * 1024
    assert info["passes"] * info["outputs_per_pass"] >= info["outputs"] and info["outputs_per_pass"] <= 128
    g = fl.notch(n, s)
    rng = np.random.default_rng(n)
    for trial, thr in enumerate([0.7, 12.0, 3.0, 0.0004]):
        x = rng.standard_normal(n)
        if trial == 1:
            x = np.cumsum(x) / 10.0
        if trial == 2:
            x = np.ones(n)
        x = np.clip(x / np.abs(x).max(), -1, 1) * thr  # |x| <= thr like the in-painted background
        ref = x - fftpack.irfft(fftpack.rfft(x) * g)
        y, d = E.notch_umma_apply_host(x, s, thr)
        assert np.abs(y - ref).max() <= 3e-6 * thr, (trial, d, np.abs(y - ref).max() / thr)


def test_tensor_core_geometry_limits():
    assert E.notch_umma_info(68, 4.0)["eligible"] == 0      # tiny bands stay on the CUDA-core kernel
    assert E.notch_umma_info(2050, 128.0)["eligible"] == 0  # rows beyond 33 x 32 elements
    i = E.notch_umma_info(1026, 64.125)
    assert (i["passes"], i["outputs_per_pass"], i["k_chunks"]) == (5, 112, 33)
    assert i["smem_bytes"] <= 200 * 1024


def test_hybrid_design_is_much_cheaper_than_dense_on_production_bands():
    for n, s in [(1026, 64.125), (1002, 64.16), (515, 32.19), (503, 32.24)]:
        d = E.notch_design(n, s, 1e-6)
        assert d["J"] > 0 and (d["ntap_e"] + d["ntap_o"]) / 2 + 2 * d["J"] < 0.45 * n


def test_foreground_threshold_matches_float16_rule():
    thr = E.foreground_threshold(0.3)
    bits = np.arange(0, 0x7C00, dtype=np.uint16)  # all finite non-negative float16
    h = bits.view(np.float16)
    with np.errstate(over="ignore"):
        rule = fl.foreground_fraction(h, 400, 20) > 0.3
    np.testing.assert_array_equal(rule, h.astype(np.float32) >= thr)
    v = np.arange(65536, dtype=np.uint16)
    np.testing.assert_array_equal(v.astype(np.float16).astype(np.float32) >= thr, v >= 384)
    assert E.foreground_threshold(1.0) == np.inf


def test_float32_foreground_threshold_equals_the_float16_rule():
    # the kernels compare the float32 pixel with dstr_foreground_threshold_f32 instead of rounding to float16
    rng = np.random.default_rng(5)
    for tm in [0.3, 0.0, 0.05, 0.5, 0.7, 0.95]:  # 0.0: the float16 sigmoid underflows below ~178
        h = np.float32(E.foreground_threshold(tm))
        t = np.float32(E.foreground_threshold_f32(tm))
        v = np.concatenate([np.arange(65536, dtype=np.float32),
                            (h + rng.uniform(-2, 2, 20000)).astype(np.float32),
                            np.nextafter(t, np.float32(-np.inf), dtype=np.float32)[None], t[None],
                            np.float32([-1e6, 1e6, np.inf, -np.inf, np.nan])])
        with np.errstate(over="ignore", invalid="ignore"):
            rule = v.astype(np.float16).astype(np.float32) >= h
            np.testing.assert_array_equal(rule, v >= t)
    assert np.isnan(E.foreground_threshold_f32(1.0))  # never foreground


def test_make_params_validation():
    p = E.make_params({"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12})
    assert p.level == -1 and p.sigma == 128.0 and p.max_threshold == 12.0
    assert E.make_params({"level": 3, "sigma": 1, "max_threshold": 2}).level == 3
    with pytest.raises(NotImplementedError):
        E.make_params({"wavelet": "haar", "level": 1, "sigma": 1, "max_threshold": 1})
    with pytest.raises(ValueError):
        E.make_params({"level": 1, "sigma": 0, "max_threshold": 1})
    with pytest.raises(ValueError):
        E.make_params({"level": -2, "sigma": 1, "max_threshold": 1})


def test_host_helpers_known_answers():
    # same known answers as /root/reference/code/tests/test_filtering.py:16-39,116-149,182-224
    assert fl.sigmoid(np.array(0)) == pytest.approx(0.5)
    img = np.array([10, 20, 30, 40, 50])
    np.testing.assert_array_almost_equal(fl.foreground_fraction(img, 30, 10), 1 / (1 + np.exp(-(img - 30) / 10)))
    np.testing.assert_array_almost_equal(fl.notch(5, 1.0), 1 - np.exp(-(np.arange(5) ** 2) / 2.0))
    for bad in ((0, 1.0), (-1, 1.0), (5, -1)):
        with pytest.raises(ValueError):
            fl.notch(*bad)
    np.testing.assert_array_equal(fl.gaussian_filter((1, 1), 1.0), [[0.0]])
    np.testing.assert_array_equal(fl.invert_image(np.array([[1, 2], [3, 4]])), [[3, 2], [1, 0]])
    n = fl.normalize_image([np.array([[0, 50]]), np.array([[200, 350]])])
    assert n.min() == 1.0 and n.max() == 2.0 and n.dtype == np.float16
    cfg = {"X1": {"Y1": 0, "Y2": 1}}
    flats = [np.zeros((2, 2)), np.ones((2, 2))]
    assert fl.get_hemisphere_flatfield("X1_Y2", cfg, flats) is flats[1]
    with pytest.raises(KeyError):
        fl.get_hemisphere_flatfield("X3_Y1", cfg, flats)


def test_flatfield_shape_errors_raise_before_any_device_work():
    image_tiles = np.array([[[10, 20], [30, 40]]])
    flat = np.array([[[2, 2], [2, 2]]])
    dark = np.array([[[1, 1], [1, 1]]])
    with pytest.raises(ValueError):
        fl.flatfield_correction(image_tiles, flat, dark[:-1])
    with pytest.raises(ValueError):
        fl.flatfield_correction(image_tiles, flat[:, :1], dark)


def test_engine_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EngineError):
        E.DestripeEngine(64, 64)
    with pytest.raises(E.EngineError):
        fl.log_space_fft_filtering(np.zeros((64, 64), np.float32), level=1)


def test_z_slab_partition():
    for n, w in [(2000, 8), (2000, 1), (128, 2), (100, 4), (64, 8), (1, 3)]:
        slabs = [zd.z_slab(n, r, w) for r in range(w)]
        assert slabs[0][0] == 0 and slabs[-1][1] == n
        for (a0, a1), (b0, b1) in zip(slabs, slabs[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(a % 64 == 0 for a, _ in slabs if a < n)
        sizes = [b - a for a, b in slabs]
        assert max(sizes) - min(s for s in sizes) <= 64 or n < 64 * w
    with pytest.raises(ValueError):
        zd.z_slab(10, 3, 2)


def test_slice_helpers_and_padding():
    pos, start, stop = zd.recover_global_position(
        (slice(384, 768), slice(0, 1600), slice(0, 2000)), (slice(64, 128), slice(0, 1600), slice(0, 2000))
    )
    assert pos == (slice(448, 512), slice(0, 1600), slice(0, 2000)) and start == (448, 0, 0)
    g, l = zd.unpad_global_coords(pos, (64, 1600, 2000), (0, 0, 0), (1, 1, 2000, 1600, 2000))
    assert g == pos and l == (slice(0, 64), slice(0, 1600), slice(0, 2000))
    with pytest.raises(NotImplementedError):
        zd.unpad_global_coords(pos, (64, 1600, 2000), (0, 2, 2), (1, 1, 2000, 1600, 2000))
    assert zd.pad_array_n_d(np.zeros((2, 3))).shape == (1, 1, 1, 2, 3)
    with pytest.raises(ValueError):
        zd.pad_array_n_d(np.zeros((2, 3)), dim=6)


def test_synthetic_generator_is_seeded_and_streaked():
    a = pkg.synthetic.synthetic_plane(128, 160, seed=3)
    b = pkg.synthetic.synthetic_plane(128, 160, seed=3)
    c = pkg.synthetic.synthetic_plane(128, 160, seed=4)
    assert a.dtype == np.uint16 and np.array_equal(a, b) and not np.array_equal(a, c)
    st = pkg.synthetic.synthetic_stack(4, 96, 96, cells_every=2)
    assert st.shape == (4, 96, 96) and st[1].max() > 20000 and st[0].max() < 20000
    fg = [p[p >= 384].mean() for p in st]
    assert fg[1] > 2500 and fg[3] > 2500 and fg[0] < 2500 and fg[2] < 2500
