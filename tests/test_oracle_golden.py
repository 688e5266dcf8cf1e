"""The C restatement (oracle/oracle.c) against the golden vectors produced by the reference
itself (tests/golden/make_golden.py) -- bit-exact on every case, CPU only."""
import numpy as np
import pytest

from tests.golden.make_golden import DTYPES, SHIFTS, STRATEGIES


@pytest.mark.parametrize("dt", DTYPES)
def test_translate_all_dtypes(golden, port, dt):
    img = golden[f"tr_{dt}_in"]
    shifts = golden["tr_shifts"]
    assert len(shifts) == len(SHIFTS)
    for si, st in enumerate(STRATEGIES):
        for k, (dx, dy) in enumerate(shifts):
            got = port.translate(img, dx, dy, st, background=1)
            want = golden[f"tr_{dt}_{si}_{k}"]
            assert got.dtype == want.dtype
            np.testing.assert_array_equal(got, want, err_msg=f"{dt} {st!r} dx={dx} dy={dy}")


def test_translate_ir_frame(golden, port):
    f = golden["tr_ir_in"]
    for si, st in enumerate(STRATEGIES):
        np.testing.assert_array_equal(port.translate(f, 1.3, -2.7, st, background=7), golden[f"tr_ir_{si}"])


def test_translate_errors(port):
    img = np.zeros((4, 5), dtype=np.uint16)
    with pytest.raises(RuntimeError):
        port.translate(np.zeros((2, 3, 4), dtype=np.uint16), 1, 1)
    with pytest.raises(RuntimeError):
        port.translate(img, 1, 1, "background", None)
    with pytest.raises(RuntimeError):
        port.translate(img, 1, 1, "no-such-strategy", 0)
    with pytest.raises(RuntimeError):
        port.translate(img.astype(np.float16), 1, 1)


def test_gaussian(golden, port):
    g = golden["ga_in"]
    for k, s in enumerate(golden["ga_sigmas"]):
        np.testing.assert_array_equal(port.gaussian_filter(g, float(s)), golden[f"ga_{k}"])


@pytest.mark.parametrize("k", range(5))
def test_bad_pixels(golden, port, k):
    first, other = golden[f"bp_{k}_first"], golden[f"bp_{k}_other"]
    xy, _thr, clamp = port.bad_pixels_detect(first)
    np.testing.assert_array_equal(xy, golden[f"bp_{k}_xy"])
    np.testing.assert_array_equal(port.bad_pixels_correct_with(xy, clamp, first), golden[f"bp_{k}_first_out"])
    np.testing.assert_array_equal(port.bad_pixels_correct_with(xy, clamp, other), golden[f"bp_{k}_other_out"])
    np.testing.assert_array_equal(port.loader_remove_motion(first, 1.3, -2.7), golden[f"bp_{k}_motion"])


@pytest.mark.parametrize("k", range(5))
def test_bad_pixels_beyond_the_int_product_range(port, k):
    """Saturated / dead pixels far from the median: the reference's int product wraps (x86-64 builds); the compiled
    reference's answers are in tests/golden/bp_extreme_golden.npz and the restatement spells the wrap out."""
    import os

    from tests import bp_extreme_cases as bc

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bp_extreme_golden.npz"))
    first = bc.frames()[k]
    xy, _thr, clamp = port.bad_pixels_detect(first)
    np.testing.assert_array_equal(np.asarray(xy).reshape(-1, 2), g[f"xy_{k}"])
    np.testing.assert_array_equal(port.bad_pixels_correct_with(xy, clamp, first), g[f"first_out_{k}"])
    np.testing.assert_array_equal(port.bad_pixels_correct_with(xy, clamp, bc.second_frame(first, k)), g[f"other_out_{k}"])


def test_find_median_pixel(golden, port):
    f, m = golden["mp_in"], golden["mp_mask"]
    for p, want, want_m in zip(golden["mp_percents"], golden["mp_out"], golden["mp_out_mask"]):
        assert port.find_median_pixel(f, float(p)) == want
        assert port.find_median_pixel(f, float(p), m) == want_m


# ---- pre-coder: pinned by the code it restates and by the writer's identity guarantee ----
def test_split_merge_layouts(port):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 65536, (13, 37), dtype=np.uint16)
    it = rng.integers(0, 256, (13, 37), dtype=np.uint8)
    y, u, v = port.split_444(img, it)
    assert u.shape == (13, 64)  # linesize is 32-byte aligned (h264.cpp:1041)
    np.testing.assert_array_equal(u[:, :37], img & 0xFF)
    np.testing.assert_array_equal(v[:, :37], img >> 8)
    np.testing.assert_array_equal(y[:, :37], it)
    assert (u[:, 37:] == 0xAA).all()  # row padding untouched
    back, it2 = port.merge_444(y, u, v, 37)
    np.testing.assert_array_equal(back, img)
    np.testing.assert_array_equal(it2, it)
    y0, _, _ = port.split_444(img)
    assert (y0[:, :37] == 0).all()
    p = port.split_420(img)
    np.testing.assert_array_equal(p[:13, :37], img & 0xFF)
    np.testing.assert_array_equal(p[13:, :37], img >> 8)
    np.testing.assert_array_equal(port.merge_420(p, 37), img)


def test_key_frame_rule(port):
    k = port.key_frames(130, 50)
    assert list(np.flatnonzero(k)) == [0, 50, 100]
    assert list(np.flatnonzero(port.key_frames(5, 1))) == [0, 1, 2, 3, 4]
    assert list(np.flatnonzero(port.key_frames(7, 3))) == [0, 3, 6]


@pytest.mark.parametrize("delta", [False, True])
def test_precode_round_trip(port, delta):
    from tests.conftest import ir_movie

    mov = ir_movie(23, 20, 28)
    lo, hi = port.precode_movie(mov, gop=5, delta=delta)
    if not delta:
        np.testing.assert_array_equal(lo, mov & 0xFF)
        np.testing.assert_array_equal(hi, mov >> 8)
    else:
        res = mov.copy()
        res[1:] = mov[1:] - mov[:-1]
        res[::5] = mov[::5]
        np.testing.assert_array_equal(lo, res & 0xFF)
        np.testing.assert_array_equal(hi, res >> 8)
    np.testing.assert_array_equal(port.decode_movie(lo, hi, gop=5, delta=delta), mov)


def test_stats(port):
    from tests.conftest import ir_movie

    mov = ir_movie(4, 16, 24)
    lo, hi, hist = port.movie_stats(mov)
    assert lo == mov.min() and hi == mov.max()
    np.testing.assert_array_equal(hist, np.bincount(mov.ravel(), minlength=65536))
    assert port.quantile_from_hist(hist, 0.5) == port.find_median_pixel(mov.reshape(1, -1), 0.5)
    b = port.get_background(mov[0])
    h4 = np.bincount(mov[0].ravel() >> 2, minlength=16384)
    assert b == (int(np.argmax(h4)) << 2) + 1
