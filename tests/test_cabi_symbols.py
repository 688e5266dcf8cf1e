"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/librir_b200.h declares (and every reference symbol of signal_processing.h), and --
with no GPU -- every compute entry FAILS LOUDLY instead of falling back to the CPU."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

from librir_b200 import _lib, signal_processing as sp, video_io as vio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_EXPORTS = [  # signal_processing.h:29-94 of the reference
    "translate", "gaussian_filter", "find_median_pixel", "find_median_pixel_mask", "extract_times",
    "resample_time_serie", "bad_pixels_create", "bad_pixels_correct", "bad_pixels_destroy", "label_image",
    "keep_largest_area", "hash_bytes",
]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "librir_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"RIRB_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)


def test_header_declares_reference_interface():
    names = declared_symbols()
    assert len(names) == len(set(names)) and len(names) >= 35
    for n in REFERENCE_EXPORTS:
        assert n in names


def test_library_exports_every_declared_symbol():
    lib = ct.CDLL(_lib.lib_path())
    for n in declared_symbols():
        assert hasattr(lib, n), f"{n} declared in include/librir_b200.h but not exported"
    assert set(declared_symbols()) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_library_matches_reference_glob():
    # librir/low_level/misc.py:115 globs "*signal_processing*.so"
    assert "signal_processing" in os.path.basename(_lib.lib_path())


def test_version_and_host_only_entries():
    lib = _lib.load()
    assert b"sm_100a" in lib.rirb_version()
    key = vio.key_frames(130, 50)
    assert list(np.flatnonzero(key)) == [0, 50, 100]
    assert lib.rirb_key_frames(3, 50, None) == -1 and "bad arguments" in _lib.last_error()


@pytest.mark.skipif(_lib.device_available(), reason="checks behaviour WITHOUT a CUDA device")
def test_no_gpu_means_failure_not_fallback():
    img = np.arange(20 * 24, dtype=np.uint16).reshape(20, 24)
    for call in (
        lambda: sp.translate(img, 1.5, 0.5, "nearest"),
        lambda: sp.gaussian_filter(img, 1.0),
        lambda: sp.find_median_pixel(img),
        lambda: sp.bad_pixels_create(img),
        lambda: vio.split_yuv444(img),
        lambda: vio.precode_movie(img[None]),
        lambda: vio.read_movie(img[None].astype(np.uint8), img[None].astype(np.uint8)),
        lambda: vio.remove_motion(img[None], [0.5], [0.5]),
    ):
        with pytest.raises(RuntimeError) as e:
            call()
        assert "no CPU fallback" in str(e.value) or "no usable CUDA device" in str(e.value)
    # the registration front end and the lossy pre-conditioner refuse to open
    lib = _lib.load()
    assert lib.rirb_ecc_open(448, 358) == 0 and "no CPU fallback" in _lib.last_error()
    from librir_b200 import registration

    with pytest.raises(RuntimeError) as e:
        registration.MaskedRegistratorECC()
    assert "no CPU fallback" in str(e.value)
    assert lib.rirb_lossy_open(64, 48, 45, 6, 2, 5.0, 32, 0, 0) == 0 and "no" in _lib.last_error()


def test_argument_errors_match_reference_conventions():
    img = np.zeros((4, 5), dtype=np.uint16)
    with pytest.raises(RuntimeError):  # test_rir.py:202-211 of the reference
        sp.translate(np.zeros((2, 3, 4), dtype=np.uint16), 1, 1)
    with pytest.raises(RuntimeError):
        sp.translate(img, 1, 1, "background", None)
    with pytest.raises(RuntimeError):
        sp.translate(img.astype(np.float16), 1, 1)
    with pytest.raises(RuntimeError):
        sp.translate(img, 1, 1, "no-such-strategy", 0)
    with pytest.raises(RuntimeError):
        sp.gaussian_filter(np.zeros((2, 3, 4)), 1.0)
    with pytest.raises(RuntimeError):
        sp.find_median_pixel(np.zeros((2, 3, 4)))
    with pytest.raises(RuntimeError):  # test_rir.py:275-277: unknown handle
        sp.bad_pixels_correct(0, img)
    lib = _lib.load()
    assert lib.rirb_bad_pixels_count(12345) == -1


def test_new_entries_argument_errors():
    lib = _lib.load()
    lo = np.zeros((2, 8, 16), np.uint8)
    with pytest.raises(RuntimeError):  # planes of different shapes
        vio.read_movie(lo, np.zeros((2, 8, 8), np.uint8))
    with pytest.raises(RuntimeError):  # one shift per frame
        vio.read_movie(lo, lo, None, 0, 0, [1.0], [1.0])
    assert lib.rirb_loader_read_movie(0, None, None, 1, 16, 8, 0, 0, None, None, 3, None) == -1
    assert lib.rirb_loader_read_movie(77, lo.ctypes.data, lo.ctypes.data, 2, 16, 8, 0, 0, None, None, 3, lo.ctypes.data) == -1
    assert "unknown handle" in _lib.last_error()
    assert lib.rirb_process_movie_host(0, None, 0, 1, 1, 1.0, None, None, b"nearest", 0, 50, 1, 0, None, None, None) == -1
    assert lib.rirb_set_parameter(b"loader_fused", b"0") == 0
    assert lib.rirb_set_parameter(b"bogus", b"1") == -1 and "unknown key" in _lib.last_error()


def test_oracle_reader_chain_composition():
    """oracle.loader_read_image is the composition of the pinned pieces, in readImage's order."""
    from oracle import oracle as O

    port = O.Port()
    rng = np.random.default_rng(2)
    img = rng.integers(0, 16384, (20, 24), dtype=np.uint16)
    lo, hi = (img & 0xFF).astype(np.uint8), (img >> 8).astype(np.uint8)
    np.testing.assert_array_equal(port.loader_read_image(lo, hi), img)
    xy = np.array([[3, 4], [4, 4], [0, 0], [23, 16]], dtype=np.int32)
    got = port.loader_read_image(lo, hi, xy, 100, 10, (1.5, -0.5))
    step = img.copy()
    step[:10] += 100
    step[:17] = port.loader_remove_bad_pixels(step[:17], xy)
    step[:17] = port.loader_remove_motion(step[:17], 1.5, -0.5)
    np.testing.assert_array_equal(got, step)
    np.testing.assert_array_equal(got[17:], img[17:])  # metadata rows: merged only (min_T rows end above them here)


def test_forwarded_entries_fail_cleanly_without_forward_lib():
    if os.environ.get("LIBRIR_B200_FORWARD_LIB"):
        pytest.skip("forward lib configured")
    lib = _lib.load()
    out = np.zeros(4)
    n = ct.c_int(4)
    assert lib.extract_times(None, 0, None, 0, out.ctypes.data_as(ct.c_void_p), ct.byref(n)) == -1
    assert "LIBRIR_B200_FORWARD_LIB" in _lib.last_error()


# ---- the video_io side of the boundary (include/librir_b200_video_io.h, libvideo_io_b200.so) ----
VIDEO_IO_REFERENCE_EXPORTS = [  # video_io.h:30-314 of the reference, the entries the path touches
    "open_camera_file", "video_file_format", "close_camera", "get_image_count", "get_image_time", "get_image_size", "get_filename",
    "supported_calibrations", "calibration_name", "load_image", "enable_bad_pixels", "bad_pixels_enabled",
    "load_motion_correction_file", "enable_motion_correction", "motion_correction_enabled", "get_attribute_count", "get_attribute",
    "get_global_attribute_count", "get_global_attribute", "set_ffmpeg_log_enabled", "h264_open_file", "h264_close_file",
    "h264_set_parameter", "h264_set_global_attributes", "h264_add_image_lossless", "h264_add_image_lossy", "h264_add_loss",
    "h264_get_low_errors", "h264_get_high_errors", "open_video_write", "image_write", "close_video",
    "open_camera_from_memory", "correct_PCR_file", "flip_camera_calibration", "set_global_emissivity", "set_emissivity", "get_emissivity",
    "support_emissivity", "camera_saturate", "calibration_files", "calibrate_inplace", "calibrate_image", "calibrate_image_inplace",
    "get_table_names", "get_table",
]


def video_io_lib_path():
    return os.path.join(os.path.dirname(_lib.lib_path()), "libvideo_io_b200.so")


def test_video_io_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "librir_b200_video_io.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"RIRB_VIO_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)
    assert len(names) == len(set(names))
    for n in VIDEO_IO_REFERENCE_EXPORTS:
        assert n in names, n
    assert "video_io" in os.path.basename(video_io_lib_path())  # librir/low_level/misc.py:117 globs "*video_io*.so"
    lib = ct.CDLL(video_io_lib_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/librir_b200_video_io.h but not exported"


@pytest.mark.skipif(_lib.device_available(), reason="checks behaviour WITHOUT a CUDA device")
def test_video_io_without_gpu_fails_loudly(tmp_path):
    lib = ct.CDLL(video_io_lib_path())
    lib.rirb_video_io_last_error.restype = ct.c_char_p
    assert lib.h264_open_file(str(tmp_path / "x.bin").encode(), 64, 48, 48) == 0
    assert b"no CUDA device" in lib.rirb_video_io_last_error()
    assert lib.open_video_write(str(tmp_path / "y.bin").encode(), 64, 48, 50, 3, 3) == 0
    img = np.zeros((48, 64), dtype=np.uint16)
    assert lib.h264_add_loss(0, img.ctypes.data_as(ct.c_void_p)) == -1
    # method 1 is host code (zstd of the raw image, byte-identical to the reference): it works without a GPU
    w = lib.open_video_write(str(tmp_path / "z.bin").encode(), 64, 48, 50, 1, 3)
    assert w > 0
    assert lib.image_write(w, img.ctypes.data_as(ct.c_void_p), ct.c_int64(5)) == 0
    lib.close_video.restype = ct.c_int64
    assert lib.close_video(w) > 256
    fmt = ct.c_int(0)
    cam = lib.open_camera_file(str(tmp_path / "z.bin").encode(), ct.byref(fmt))
    assert cam > 0 and fmt.value == 4 and lib.get_image_count(cam) == 1
    out = np.ones((48, 64), dtype=np.uint16)
    assert lib.load_image(cam, 0, 0, out.ctypes.data_as(ct.c_void_p)) == 0 and not out.any()
    assert lib.close_camera(cam) == 0


@pytest.mark.parametrize("times", [[0, 20, 40, 60, 80], [30000, 30020, 30040, 30061], [5, 1_000_000_123, 2_000_000_000], [-3, 7, 1000],
                                   [29_000_000_000, 29_020_000_000]])
def test_video_io_reader_applies_the_references_time_conventions(tmp_path, times):
    """The reference's reader rewrites the timestamps of its zstd movie files (WEST acquisition files, ms): origin-relative ms
    -> ns - 10 ms, other non-ns values -> ns relative to the first image, and -32 s when the first lands between 28 s and 32 s
    (IRFileLoader.cpp:354-372, :458-466).  Method-1 files are host code on both sides: no GPU needed."""
    from oracle import refvio as rv

    if not rv.have_ref_vio():
        pytest.skip("oracle/_ref/libs/libvideo_io.so not built")
    lib = ct.CDLL(video_io_lib_path())
    fn = str(tmp_path / "t.bin").encode()
    w = lib.open_video_write(fn, 16, 12, 50, 1, 1)
    assert w > 0
    img = np.arange(12 * 16, dtype=np.uint16).reshape(12, 16)
    for t in times:
        assert lib.image_write(w, img.ctypes.data_as(ct.c_void_p), ct.c_int64(t)) == 0
    lib.close_video.restype = ct.c_int64
    assert lib.close_video(w) > 0
    ref = rv.Camera(fn.decode())
    fmt = ct.c_int(0)
    cam = lib.open_camera_file(fn, ct.byref(fmt))
    assert cam > 0 and fmt.value == 4
    for i in range(len(times)):
        t = ct.c_int64(0)
        assert lib.get_image_time(cam, i, ct.byref(t)) == 0
        assert t.value == ref.image_time(i), (i, t.value, ref.image_time(i))
    lib.close_camera(cam)
    ref.close()


def _raw_movie(kind, n, h, w, stamp):
    """A raw movie file as the reference's reader expects it (IRFileLoader.cpp:124-207): kind 'pcr' (1024-byte header),
    'pcr_enc' (133 bytes of envelope in front of the header) or 'west' (BIN_HEADER + one BIN_TRIGGER, 128 bytes each)."""
    rng = np.random.default_rng(n * 1000 + h)
    mov = rng.integers(0, 16384, (n, h, w), dtype=np.uint16)
    if stamp is not None:  # findTimes: the last 8 bytes of every image, strictly increasing
        for i in range(n):
            mov[i].reshape(-1)[-4:] = np.frombuffer(np.int64(stamp(i)).tobytes(), dtype=np.uint16)
    pcr = np.zeros(256, np.int32)
    pcr[[1, 2, 3, 5, 7, 9, 10, 11]] = [n, w, h, 16, 25, w * h * 2, w, h]
    if kind == "pcr":
        head = pcr.tobytes()
    elif kind == "pcr_enc":
        head = bytes(133) + pcr.tobytes()
    else:
        hdr = bytearray(128)
        hdr[0], hdr[1], hdr[2] = 1, 1, 0
        trig = np.zeros(16, np.int64)
        trig[[0, 1, 2, 9, 10]] = [1234, 25, n, w, h]
        head = bytes(hdr) + trig.tobytes()
    return head + mov.tobytes(), mov


@pytest.mark.parametrize("kind,stamp", [("pcr", None), ("pcr", lambda i: 1000 + 40 * i), ("pcr_enc", None), ("west", lambda i: 30000 + 40 * i),
                                        ("west", None), ("pcr", lambda i: 5_000_000_000 + 40_000_000 * i)])
def test_video_io_reader_opens_raw_movies_like_the_reference(tmp_path, kind, stamp):
    """Raw PCR / encapsulated PCR / uncompressed WEST files (what IRMovie.from_numpy_array writes, and what acquisition
    systems wrote before compression): format code, frame count, size, every image and every timestamp as the compiled
    reference's reader gives them.  Host code on both sides: no GPU needed."""
    from oracle import refvio as rv

    if not rv.have_ref_vio():
        pytest.skip("oracle/_ref/libs/libvideo_io.so not built")
    data, mov = _raw_movie(kind, 7, 24, 40, stamp)
    fn = tmp_path / f"raw.{kind}"
    fn.write_bytes(data)
    lib = ct.CDLL(video_io_lib_path())
    fmt = ct.c_int(0)
    cam = lib.open_camera_file(str(fn).encode(), ct.byref(fmt))
    ref = rv.Camera(str(fn))
    assert cam > 0 and fmt.value == {"pcr": 1, "pcr_enc": 3, "west": 2}[kind]
    assert lib.get_image_count(cam) == ref.count == len(mov)
    img = np.empty(mov.shape[1:], np.uint16)
    for i in range(len(mov)):
        assert lib.load_image(cam, i, 0, img.ctypes.data_as(ct.c_void_p)) == 0
        assert np.array_equal(img, ref.load_image(i)) and np.array_equal(img, mov[i]), i
        t = ct.c_int64(0)
        assert lib.get_image_time(cam, i, ct.byref(t)) == 0 and t.value == ref.image_time(i), (i, t.value, ref.image_time(i))
    lib.close_camera(cam)
    ref.close()
