"""Drop-in demonstration: the REFERENCE's unmodified Python package (librir/low_level/misc.py
loadDlls + librir/signal_processing wrappers) loads libsignal_processing_b200.so as its
signal_processing library.  Needs /root/reference (authoring container only; skipped on the GPU
box, where the reference tree does not exist).  On a CPU-only machine the calls must raise
RuntimeError (no CPU fallback); with a GPU they must agree with the compiled reference."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LIBRIR_REFERENCE", "/root/reference")
REF_LIBS = os.path.join(ROOT, "oracle", "_ref", "libs")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src/python/librir")), reason="reference tree not present")
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_LIBS, "libtools.so")), reason="oracle/_ref not built")
def test_reference_python_package_loads_our_library(tmp_path):
    from librir_b200 import _lib

    pkg = tmp_path / "site" / "librir"
    pkg.parent.mkdir()
    # the package itself is the reference's, symlinked file by file (nothing is copied into the repo)
    src = os.path.join(REF, "src/python/librir")
    for dirpath, dirnames, filenames in os.walk(src):
        rel = os.path.relpath(dirpath, src)
        (pkg / rel).mkdir(parents=True, exist_ok=True)
        for fn in filenames:
            os.symlink(os.path.join(dirpath, fn), pkg / rel / fn)
    libs = pkg / "libs"
    libs.mkdir(exist_ok=True)
    os.symlink(os.path.join(REF_LIBS, "libtools.so"), libs / "libtools.so")
    os.symlink(os.path.join(REF_LIBS, "libgeometry.so"), libs / "libgeometry.so")
    os.symlink(_lib.lib_path(), libs / "libsignal_processing.so")  # <- the swap
    stub = tmp_path / "stub.c"
    stub.write_text("int librir_b200_video_io_stub(void) { return 0; }\n")
    subprocess.check_call(["/usr/bin/gcc", "-shared", "-fPIC", str(stub), "-o", str(libs / "libvideo_io.so")])
    code = textwrap.dedent("""
        import numpy as np, ctypes as ct
        from librir.low_level import misc
        from librir.signal_processing import rir_signal_processing as rsp
        from librir.signal_processing.BadPixels import BadPixels
        lib = misc._signal_processing
        lib.rirb_version.restype = ct.c_char_p
        assert b"sm_100a" in lib.rirb_version()
        lib.rirb_device_count.restype = ct.c_int
        img = (np.arange(48 * 64).reshape(48, 64) % 5000 + 7000).astype(np.uint16)
        if lib.rirb_device_count() == 0:
            for call in (lambda: rsp.translate(img, 1.5, -0.5, "nearest"), lambda: BadPixels(img).correct(img)):
                try:
                    call()
                except RuntimeError:
                    continue
                raise SystemExit("expected RuntimeError without a GPU")
            print("DROPIN-OK no-gpu")
        else:
            out = rsp.translate(img, 1.5, -0.5, "nearest")
            assert out.shape == img.shape and out.dtype == img.dtype
            g = rsp.gaussian_filter(img, 1.0)
            assert g.dtype == np.float32
            c = BadPixels(img).correct(img)
            assert c.shape == img.shape
            print("DROPIN-OK gpu")
    """)
    env = dict(os.environ, PYTHONPATH=str(tmp_path / "site"), LIBRIR_DISABLE_JOBLIB="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert "DROPIN-OK" in res.stdout, res.stdout + res.stderr
