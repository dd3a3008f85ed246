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
