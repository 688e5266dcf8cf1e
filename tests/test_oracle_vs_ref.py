"""The C restatement against the reference's own sources compiled into oracle/_ref, live, on
seeded inputs larger and more varied than the golden file.  Skipped where oracle/_ref was not
built (it is built wherever /root/reference exists and then travels with the repo)."""
import numpy as np
import pytest

from tests.conftest import ir_frame
from tests.golden.make_golden import typed_image

STRATS = ["", "noborder", "background", "wrap", "nearest", "constant"]


@pytest.mark.parametrize("dt", ["bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64",
                                "float32", "float64"])
def test_translate(port, ref, dt):
    rng = np.random.default_rng(7)
    img = typed_image(dt, 37, 53, rng)
    for st in STRATS:
        for dx, dy in [(1.3, -2.7), (0, 0), (-0.25, 0.5), (3, 2), (-60.5, 10.25), (0.99999997, 1e-30),
                       (52.999, 36.999), (-1.25, 1.25), (1e9, -1e9)]:
            a = port.translate(img, dx, dy, st, background=1)
            b = ref.translate(img, dx, dy, st, background=1)
            np.testing.assert_array_equal(a, b, err_msg=f"{dt} {st} {dx} {dy}")


def test_translate_full_frame(port, ref):
    f = ir_frame(512, 640, 21)
    for st in ["nearest", "wrap", "background", ""]:
        np.testing.assert_array_equal(port.translate(f, 1.3, -2.7, st, 0), ref.translate(f, 1.3, -2.7, st, 0))


def test_gaussian(port, ref):
    rng = np.random.default_rng(8)
    for sig in [0.3, 0.5, 1.0, 1.7, 2.0, 3.2]:
        img = (rng.random((41, 67)) * 5000).astype(np.float32)
        np.testing.assert_array_equal(port.gaussian_filter(img, sig), ref.gaussian_filter(img, sig))
    f = ir_frame(256, 320, 22).astype(np.float32)
    np.testing.assert_array_equal(port.gaussian_filter(f, 1.0), ref.gaussian_filter(f, 1.0))


@pytest.mark.parametrize("shape", [(64, 80), (33, 47), (5, 5), (7, 4), (128, 160), (3, 9), (1, 1), (2, 2), (256, 320)])
def test_bad_pixels(port, ref, shape):
    h, w = shape
    first = ir_frame(h, w, h * 1000 + w)
    xy, _thr, clamp = port.bad_pixels_detect(first)
    np.testing.assert_array_equal(xy, ref.bad_pixels_list(first))
    hd = ref.bad_pixels_create(first)
    assert hd > 0
    for s in range(3):
        g = ir_frame(h, w, s + 77)
        np.testing.assert_array_equal(port.bad_pixels_correct_with(xy, clamp, g), ref.bad_pixels_correct(hd, g))
    ref.bad_pixels_destroy(hd)
    with pytest.raises(RuntimeError):
        ref.bad_pixels_correct(hd, first)


def test_median_and_motion(port, ref):
    rng = np.random.default_rng(9)
    f = ir_frame(96, 128, 31)
    m = rng.random(f.shape) > 0.4
    for pc in [0.5, 0.1, 0.9, 0.0, 1.0, 0.37]:
        assert port.find_median_pixel(f, pc) == ref.find_median_pixel(f, pc)
        assert port.find_median_pixel(f, pc, m) == ref.find_median_pixel(f, pc, m)
    for sx, sy in [(1.3, -2.7), (0.0, 0.0), (-3.5, 2.25), (100.0, 100.0), (0.5, 0.5)]:
        np.testing.assert_array_equal(port.loader_remove_motion(f, sx, sy), ref.loader_remove_motion(f, sx, sy))
