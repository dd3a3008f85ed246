"""GPU edge cases: tiny and odd planes, wide rows, non-integer float input, ragged chunk sizes,
explicit levels, concurrent engines."""
import threading

import numpy as np
import pytest

from _parity import REL_TOL, U16_FRACTION, rel_err, u16_agreement
from aind_smartspim_destripe_b200 import engine as E
from aind_smartspim_destripe_b200 import filtering as fl
from aind_smartspim_destripe_b200 import synthetic as S
from oracle import plane_filter as OF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,level", [((4, 4), None), ((8, 8), 1), ((5, 64), None), ((13, 17), 1), ((13, 17), None),
                                         ((31, 200), None), ((64, 33), 2), ((10, 10), 1), ((6, 7), 1)])
def test_tiny_and_odd_planes(shape, level):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    img = rng.integers(50, 4000, shape).astype(np.float32)
    ref = OF.log_space_fft_filtering(img, level=level, sigma=16, max_threshold=4)[: shape[0], : shape[1]]
    out = fl.log_space_fft_filtering(img, level=level, sigma=16, max_threshold=4)
    assert out.shape == shape
    e = rel_err(out, ref)
    print(f"{shape} level {level}: rel {e:.2e}")
    assert e < REL_TOL


def test_wide_rows_use_the_largest_row_kernel():
    # W = 4096 -> level-1 band rows of 2050 coefficients (65 elements per lane)
    img = S.synthetic_plane(96, 4096, seed=3).astype(np.float32)
    conf = dict(level=None, sigma=128, max_threshold=12)
    ref = OF.log_space_fft_filtering(img, **conf)
    out = fl.log_space_fft_filtering(img, **conf)
    frac, mx, _ = u16_agreement(np.clip(out, 0, 65535).astype(np.uint16), np.clip(ref, 0, 65535).astype(np.uint16))
    assert frac >= U16_FRACTION and rel_err(out, ref) < REL_TOL
    with pytest.raises(ValueError):  # rows longer than the kernel supports are refused, not mis-filtered
        fl.log_space_fft_filtering(np.ones((64, 9000), np.float32), level=1, sigma=8, max_threshold=3)


def test_non_integer_float_input():
    rng = np.random.default_rng(11)
    img = (S.synthetic_plane(200, 240, seed=9).astype(np.float32) + rng.random((200, 240)).astype(np.float32) * 0.9)
    conf = dict(level=None, sigma=64, max_threshold=3)
    ref = OF.log_space_fft_filtering(img, **conf)
    out = fl.log_space_fft_filtering(img, **conf)
    assert rel_err(out, ref) < REL_TOL
    fg, bg, _ = fl.get_foreground_background_mean(img)
    rfg, rbg, _ = OF.get_foreground_background_mean(img)
    assert fg == pytest.approx(float(rfg), rel=1e-5) and bg == pytest.approx(float(rbg), rel=1e-5)


@pytest.mark.parametrize("level", [1, 2, 3, 5])
def test_explicit_levels(level):
    img = S.synthetic_plane(256, 320, seed=level).astype(np.float32)
    ref = OF.log_space_fft_filtering(img, level=level, sigma=64, max_threshold=4)
    out = fl.log_space_fft_filtering(img, level=level, sigma=64, max_threshold=4)
    assert rel_err(out, ref) < REL_TOL


@pytest.mark.parametrize("Z", [1, 3, 7, 9])
def test_ragged_chunk_sizes_through_the_host_pipeline(Z):
    st = S.synthetic_stack(Z, 128, 160, base_seed=60)
    eng = E.DestripeEngine(128, 160, max_planes=4)
    p = E.make_params(dict(level=None, sigma=128, max_threshold=12))
    eng.set_subchunk(2)
    out = eng.filter_chunk(st, p, out_dtype=np.uint16)
    for z in range(Z):
        one = eng.filter_chunk(st[z : z + 1], p, out_dtype=np.uint16)[0]
        np.testing.assert_array_equal(out[z], one)
    eng.close()


def test_two_engines_in_two_threads():
    st = S.synthetic_stack(6, 160, 192, base_seed=70)
    p = E.make_params(dict(level=None, sigma=128, max_threshold=12))
    results, errors = {}, []

    def work(k):
        try:
            eng = E.DestripeEngine(160, 192, max_planes=4)
            for _ in range(3):
                results[k] = eng.filter_chunk(st, p, out_dtype=np.uint16)
            eng.close()
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors
    np.testing.assert_array_equal(results[0], results[1])


def test_expm1_flag_and_f32_output():
    img = S.synthetic_plane(128, 160, seed=5)
    eng = E.DestripeEngine(128, 160, max_planes=2)
    p = E.make_params(dict(level=None, sigma=128, max_threshold=12))
    a = eng.filter_chunk(img[None], p, out_dtype=np.float32)
    b = eng.filter_chunk(img[None], p, out_dtype=np.float32, flags=E.FLAG_EXPM1)
    np.testing.assert_allclose(a - b, 2.0, atol=1e-2)  # exp(y) + 1 (reference) vs exp(y) - 1 (corrected)
    eng.close()


def _degenerate_planes(H, W):
    rng = np.random.default_rng(11)
    step = np.full((H, W), 200, np.uint16)
    step[:, W // 2 :] = 4000
    hot = np.full((H, W), 300, np.uint16)
    hot[H // 3, W // 5] = 65535
    rows = (100 + 50 * (np.arange(H) % 7)).astype(np.uint16)[:, None].repeat(W, 1)  # pure horizontal streaks
    return {
        "all zero": np.zeros((H, W), np.uint16),
        "constant 1000": np.full((H, W), 1000, np.uint16),
        "saturated": np.full((H, W), 65535, np.uint16),
        "vertical step": step,
        "one hot pixel": hot,
        "streaks only": rows,
        "full-range noise": rng.integers(0, 65536, (H, W)).astype(np.uint16),
        "binary 0 / 65535": (rng.integers(0, 2, (H, W)) * 65535).astype(np.uint16),
    }


@pytest.mark.parametrize("name", list(_degenerate_planes(8, 8)))
def test_degenerate_planes_match_the_oracle(name, production_configs):
    """Constant, saturated, binary and single-outlier planes: zero-width histograms (threshold_otsu returns the
    value itself), empty masks, all-masked rows, log(1 + 65535) at the top of the range."""
    no_cells, cells = production_configs
    H, W = 192, 256
    img = _degenerate_planes(H, W)[name]
    flat, dark = S.synthetic_flat_dark(H, W)
    shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
    ref = OF.filter_stripes(img.astype(np.float32), "0_0", no_cells, cells, shadow, 2500)
    out = fl.filter_stripes(img, "0_0", no_cells, cells, shadow, 2500)
    assert out.dtype == np.uint16 and np.all(np.isfinite(out.astype(np.float64)))
    frac, mx, exact = u16_agreement(out, ref)
    print(f"{name}: within+-1 {frac:.6f} exact {exact:.4f} max abs {mx}")
    assert frac >= U16_FRACTION
    # and the float path (no shadow dict): finite, within tolerance
    ref_f = OF.log_space_fft_filtering(img.astype(np.float32), **no_cells)
    out_f = fl.log_space_fft_filtering(img.astype(np.float32), **no_cells)
    assert np.all(np.isfinite(out_f))
    e = rel_err(out_f, ref_f)
    print(f"{name}: float rel {e:.2e}")
    assert e < 10 * REL_TOL
