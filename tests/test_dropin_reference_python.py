"""Drop-in demonstration: the REFERENCE's unmodified Python package (librir/low_level/misc.py loadDlls + the
librir.signal_processing / librir.video_io wrappers) with this repo's two libraries in its libs/ directory.

The package comes from oracle/_ref/pkg (laid out by oracle/build_ref.sh from /root/reference; git-ignored, shipped to the GPU
box by gpurun) -- the tests copy it, swap libs/*signal_processing* and libs/*video_io* for libsignal_processing_b200.so and
libvideo_io_b200.so, and run the reference's own wrappers in a subprocess:

* without a GPU (not marked): the calls must raise RuntimeError -- there is no CPU fallback;
* with a GPU (marked gpu): VALUES are compared with the compiled reference (oracle/_ref): translate, gaussian_filter,
  BadPixels.correct, find_median_pixel; a movie written through IRSaver / h264_add_image_lossless and read back through
  open_camera_file / load_image is the identity (tests/python/test_IRMovie.py:46-49 of the reference), h264_add_loss and
  h264_get_low/high_errors give the compiled saver's numbers."""
import os
import shutil
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "oracle", "_ref", "pkg", "librir")
OUR_LIBS = os.path.join(ROOT, "librir_b200", "libs")

needs_pkg = pytest.mark.skipif(not os.path.isfile(os.path.join(PKG, "__init__.py")),
                               reason="oracle/_ref/pkg not laid out (oracle/build_ref.sh needs /root/reference once)")


def make_site(tmp_path):
    """A copy of the reference's package whose libs/ holds the reference's tools + geometry and OUR two libraries."""
    site = tmp_path / "site"
    shutil.copytree(PKG, site / "librir", ignore=shutil.ignore_patterns("libs", "__pycache__"))
    libs = site / "librir" / "libs"
    libs.mkdir()
    for name in ("libtools.so", "libgeometry.so"):
        shutil.copy(os.path.join(PKG, "libs", name), libs / name)
    for name in ("libsignal_processing_b200.so", "libvideo_io_b200.so"):  # match the *signal_processing* / *video_io* globs
        shutil.copy(os.path.join(OUR_LIBS, name), libs / name)
    return site


def run(site, code, timeout=600):
    env = dict(os.environ, PYTHONPATH=str(site) + os.pathsep + ROOT, LIBRIR_DISABLE_JOBLIB="1")
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=timeout)


@needs_pkg
def test_reference_python_package_loads_our_libraries(tmp_path):
    site = make_site(tmp_path)
    res = run(site, """
        import numpy as np, ctypes as ct
        from librir.low_level import misc
        from librir.signal_processing import rir_signal_processing as rsp
        from librir.signal_processing.BadPixels import BadPixels
        from librir.video_io import rir_video_io as rv
        lib = misc._signal_processing
        lib.rirb_version.restype = ct.c_char_p
        assert b"sm_100a" in lib.rirb_version()
        assert hasattr(misc._video_io, "h264_add_image_lossless") and hasattr(misc._video_io, "open_video_write")
        lib.rirb_device_count.restype = ct.c_int
        img = (np.arange(48 * 64).reshape(48, 64) % 5000 + 7000).astype(np.uint16)
        if lib.rirb_device_count() == 0:
            for call in (lambda: rsp.translate(img, 1.5, -0.5, "nearest"), lambda: BadPixels(img).correct(img),
                         # h264_open_file answers 0 on failure, which the wrapper lets through (rir_video_io.py:507): the add fails
                         lambda: rv.h264_add_image_lossless(rv.h264_open_file("/tmp/x.bin", 64, 48), img, 0)):
                try:
                    call()
                except RuntimeError:
                    continue
                raise SystemExit("expected RuntimeError without a GPU")
            print("DROPIN-OK no-gpu")
        else:
            print("DROPIN-OK gpu")
    """)
    assert "DROPIN-OK" in res.stdout, res.stdout + res.stderr


@needs_pkg
@pytest.mark.gpu
def test_reference_wrappers_give_the_compiled_references_values(tmp_path, ref):
    from tests.conftest import ir_frame, ir_movie

    site = make_site(tmp_path)
    img = ir_frame(96, 160, seed=5)
    mov = ir_movie(6, 96, 160, seed=6)
    mask = (np.random.default_rng(1).random(img.shape) > 0.3).astype(np.uint8)
    np.savez(tmp_path / "in.npz", img=img, mov=mov, mask=mask)
    res = run(site, f"""
        import numpy as np
        from librir.signal_processing import rir_signal_processing as rsp
        from librir.signal_processing.BadPixels import BadPixels
        d = np.load(r"{tmp_path / 'in.npz'}")
        img, mov, mask = d["img"], d["mov"], d["mask"]
        out = dict()
        for k, (dx, dy, st) in enumerate([(1.3, -2.7, "nearest"), (-0.4, 5.25, "background"), (7.5, 0.5, "wrap"), (2.0, 3.0, "")]):
            out[f"tr{{k}}"] = rsp.translate(img, dx, dy, st, 77 if st == "background" else None)
        out["trf"] = rsp.translate(img.astype(np.float32), -1.25, 0.75, "nearest")
        for k, s in enumerate([0.5, 1.0, 2.0]):
            out[f"g{{k}}"] = rsp.gaussian_filter(img, s)
        bp = BadPixels(mov[0])
        out["bp"] = np.stack([bp.correct(f) for f in mov])
        out["med"] = np.array([rsp.find_median_pixel(img, 0.5), rsp.find_median_pixel(img, 0.02), rsp.find_median_pixel(img, 0.9, mask)])
        np.savez(r"{tmp_path / 'out.npz'}", **out)
        print("DONE")
    """)
    assert "DONE" in res.stdout, res.stdout + res.stderr
    got = np.load(tmp_path / "out.npz")
    for k, (dx, dy, st) in enumerate([(1.3, -2.7, "nearest"), (-0.4, 5.25, "background"), (7.5, 0.5, "wrap"), (2.0, 3.0, "")]):
        np.testing.assert_array_equal(got[f"tr{k}"], ref.translate(img, dx, dy, st, 77 if st == "background" else None), err_msg=st)
    want = ref.translate(img.astype(np.float32), -1.25, 0.75, "nearest")
    np.testing.assert_allclose(got["trf"], want, rtol=1e-5, atol=1e-6 * float(np.abs(want).max()))
    for k, s in enumerate([0.5, 1.0, 2.0]):
        want = ref.gaussian_filter(img, s)
        np.testing.assert_allclose(got[f"g{k}"], want, rtol=1e-5, atol=1e-6 * float(np.abs(want).max()), err_msg=f"sigma {s}")
    h = ref.bad_pixels_create(mov[0])
    np.testing.assert_array_equal(got["bp"], np.stack([ref.bad_pixels_correct(h, f) for f in mov]))
    ref.bad_pixels_destroy(h)
    assert list(got["med"]) == [ref.find_median_pixel(img, 0.5), ref.find_median_pixel(img, 0.02), ref.find_median_pixel(img, 0.9, mask)]


@needs_pkg
@pytest.mark.gpu
def test_reference_saver_and_reader_wrappers_round_trip(tmp_path):
    """IRSaver.add_image / rir_video_io.h264_add_image_lossless -> open_camera_file / load_image: the identity, with
    timestamps and attributes, as the reference's own tests ask (tests/python/test_IRMovie.py:46-49, :59-67)."""
    from tests import vio_cases as C

    site = make_site(tmp_path)
    mov = C.movie(130, 64, 80, seed=2)  # two full GOPs of 50 and a partial one
    np.save(tmp_path / "mov.npy", mov)
    res = run(site, f"""
        import numpy as np
        from librir.video_io import rir_video_io as rv
        from librir.video_io.IRSaver import IRSaver
        mov = np.load(r"{tmp_path / 'mov.npy'}")
        fn = r"{tmp_path / 'm.bin'}"
        with IRSaver(fn, mov.shape[2], mov.shape[1]) as s:
            s.set_parameter("compressionLevel", 2)
            s.set_global_attributes({{"Device": "B200", "note": "x" * 3000}})
            for i, f in enumerate(mov):
                s.add_image(f, i * 20_000_000, {{"k": str(i)}})
        cam = rv.open_camera_file(fn)
        assert rv.get_image_count(cam) == len(mov), rv.get_image_count(cam)
        assert tuple(rv.get_image_size(cam)) == mov.shape[1:]
        assert rv.supported_calibrations(cam) == ["Digital Level"]
        for i in (0, 1, 49, 50, 51, 129, 77, 3, 100):
            assert np.array_equal(rv.load_image(cam, i, 0), mov[i]), i
            assert rv.get_attributes(cam)["k"] in (str(i), str(i).encode()), rv.get_attributes(cam)
        assert rv.get_image_time(cam, 7) == 140_000_000
        g = rv.get_global_attributes(cam)
        assert g["Device"] == b"B200" and g["note"] == b"x" * 3000 and g["GOP"] == b"50", sorted(g)
        rv.close_camera(cam)
        # the zstd container shrinks with the pre-coder in front: method 3 < method 2 < method 1 on a smooth movie
        import os
        sizes = dict()
        for method in (1, 2, 3):
            w = rv._video_io.open_video_write((fn + str(method)).encode(), mov.shape[2], mov.shape[1], 50, method, 3)
            assert w > 0
            for i, f in enumerate(mov):
                f = np.ascontiguousarray(f)
                assert rv._video_io.image_write(w, f.ctypes.data_as(rv.ct.c_void_p), rv.ct.c_int64(i)) == 0
            rv._video_io.close_video.restype = rv.ct.c_int64
            sizes[method] = rv._video_io.close_video(w)
            cam = rv.open_camera_file(fn + str(method))
            assert all(np.array_equal(rv.load_image(cam, i, 0), mov[i]) for i in (0, 64, 129, 5))
            rv.close_camera(cam)
        assert sizes[3] < sizes[2] < sizes[1], sizes
        # the reference's IRMovie class on top of the same file: indexing, slicing, timestamps, attributes, the loader's
        # bad-pixel switch and its registration file
        from librir.video_io.IRMovie import IRMovie
        m = IRMovie.from_filename(fn)
        assert len(m) == len(mov) and tuple(m.image_size) == mov.shape[1:] and m.calibrations == ["Digital Level"]
        assert np.array_equal(m[5], mov[5]) and np.array_equal(m[-1], mov[-1]) and np.array_equal(m[10:60:7], mov[10:60:7])
        assert np.allclose(m.timestamps[:4], [0.0, 0.02, 0.04, 0.06]) and m.attributes["Device"] == b"B200"
        assert os.path.basename(m.filename) == "m.bin"
        m.bad_pixels_correction = True
        corrected = m[3]
        assert corrected.shape == mov[3].shape and (corrected != mov[3]).sum() < corrected.size // 50   # a few flagged pixels replaced
        m.bad_pixels_correction = False
        reg = fn + ".regfile"
        with open(reg, "w") as f:
            f.write("\\tx-axis translations\\ty-axis translations\\tConfidence level\\n")
            for i in range(len(mov)):
                f.write(f"{{i}}\\t{{0.25 * (i % 5)}}\\t{{-0.5 * (i % 3)}}\\t1.0\\n")
        m.registration_file = reg
        m.registration = True
        moved = m[7]                                # shifted by (0.5, -0.5): rows [0, h-3), the metadata rows pass through
        assert not np.array_equal(moved, mov[7]) and np.array_equal(moved[-3:], mov[7][-3:])
        assert np.array_equal(m[15], mov[15])       # frame 15: shift (0, 0)
        m.registration = False
        assert np.array_equal(m[7], mov[7])
        m.close()
        print("DONE", sizes)
    """)
    assert "DONE" in res.stdout, res.stdout + res.stderr


@needs_pkg
@pytest.mark.gpu
def test_reference_lossy_wrappers_give_the_compiled_savers_numbers(tmp_path):
    """IRSaver.add_loss / get_low_errors / get_high_errors and add_image_lossy (read back through the reader) against
    tests/golden/vio_golden.npz, which the compiled reference's saver wrote."""
    from tests import vio_cases as C

    site = make_site(tmp_path)
    g = np.load(os.path.join(ROOT, "tests", "golden", "vio_golden.npz"))
    t, h, w, stop = C.LOSSY_SHAPE
    mov = C.lossy_movie()
    np.save(tmp_path / "mov.npy", mov)
    res = run(site, f"""
        import numpy as np
        from librir.video_io import rir_video_io as rv
        from librir.video_io.IRSaver import IRSaver
        mov = np.load(r"{tmp_path / 'mov.npy'}")
        h, w, stop = {h}, {w}, {stop}
        out = dict()
        s = IRSaver(r"{tmp_path / 'loss.bin'}", w, h, stop)
        s.set_parameter("subtractMin", 1)
        out["loss"] = np.stack([s.add_loss(f.copy()) for f in mov])
        out["loss_low"], out["loss_high"] = s.get_low_errors(), s.get_high_errors()
        s.close()
        fn = r"{tmp_path / 'lossy.bin'}"
        s = IRSaver(fn, w, h, stop)
        s.set_parameter("subtractMin", 1)
        for i, f in enumerate(mov):
            s.add_image_lossy(f, i * 1000)
        out["lossy_low"], out["lossy_high"] = s.get_low_errors(), s.get_high_errors()
        s.close()
        cam = rv.open_camera_file(fn)
        ga = rv.get_global_attributes(cam)
        out["min_t"] = np.array([int(ga["MIN_T"]), int(ga["MIN_T_HEIGHT"])])
        # the reader adds MIN_T back (IRFileLoader.cpp:1174-1179): undo it to see what the saver stored
        frames, attrs = [], []
        for i in range(len(mov)):
            f = rv.load_image(cam, i, 0).astype(np.int64)
            f[:stop] -= out["min_t"][0]
            frames.append(f.astype(np.uint16))
            a = rv.get_attributes(cam)
            attrs.append((int(a.get("BackgroundError", -1)), int(a.get("ForegroundError", -1))))
        out["lossy"], out["attrs"] = np.stack(frames), np.array(attrs)
        rv.close_camera(cam)
        np.savez(r"{tmp_path / 'out.npz'}", **out)
        print("DONE")
    """)
    assert "DONE" in res.stdout, res.stdout + res.stderr
    got = np.load(tmp_path / "out.npz")
    key = "lossy_add_loss_sub_min"
    assert np.array_equal(got["loss_low"], g[key + "_low"]) and np.array_equal(got["loss_high"], g[key + "_high"])
    assert np.array_equal(np.array([C.crc(f) for f in got["loss"]], dtype=np.uint32), g[key + "_crc"])
    key = "lossy_add_image_lossy_sub_min"
    assert np.array_equal(got["lossy_low"], g[key + "_low"]) and np.array_equal(got["lossy_high"], g[key + "_high"])
    assert np.array_equal(got["min_t"], g[key + "_min_t"])
    assert np.array_equal(np.array([C.crc(f) for f in got["lossy"]], dtype=np.uint32), g[key + "_crc"])
    assert np.array_equal(got["attrs"][1:, 0], g[key + "_low"][1:]) and np.array_equal(got["attrs"][1:, 1], g[key + "_high"][1:])
