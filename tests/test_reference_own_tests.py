"""The reference's OWN Python tests (tests/python/test_*.py of IRFM/librir, unmodified) run on top of the drop-in libraries.

oracle/build_ref.sh lays the files out under oracle/_ref/reftests (git-ignored, shipped to the GPU box by gpurun; nothing of
them is in this repository); this test copies them next to a copy of the reference's package whose libs/ holds the
reference's tools + geometry and THIS repo's libsignal_processing_b200.so + libvideo_io_b200.so, runs pytest on them in a
subprocess and asserts on the outcome.  What is deselected is listed below with the reason."""
import os
import re
import shutil
import subprocess
import sys

import pytest

from tests.test_dropin_reference_python import make_site

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFTESTS = os.path.join(ROOT, "oracle", "_ref", "reftests")
needs_reftests = pytest.mark.skipif(not os.path.isfile(os.path.join(REFTESTS, "tests", "python", "conftest.py")),
                                    reason="oracle/_ref/reftests not laid out (oracle/build_ref.sh needs /root/reference once)")

# test id substring -> why it cannot run on these libraries
OUT_OF_SCOPE = {
    "test_video_file_format": "movies are stored in the zstd container (FILE_FORMAT_ZSTD_COMPRESSED = 4), not as mp4 (5): INTEGRATION.md 1b",
    "test_pcr2h264": "asserts the mp4 file format of the converted movie (the conversion itself runs in test_save_movie_with_pcr2h264)",
    "thermavip": "Thermavip shared memory bridge",
}


def run_reference_tests(tmp_path, files, extra=()):
    site = make_site(tmp_path)
    work = tmp_path / "work"
    shutil.copytree(REFTESTS, work)
    deselect = " and ".join(f"not {k}" for k in OUT_OF_SCOPE)
    # the five host utilities outside the path (extract_times, resample_time_serie, label_image ...) are forwarded to a
    # reference build of the library, as INTEGRATION.md section 1 describes
    env = dict(os.environ, PYTHONPATH=str(site) + os.pathsep + str(work), LIBRIR_DISABLE_JOBLIB="1",
               LIBRIR_B200_FORWARD_LIB=os.path.join(ROOT, "oracle", "_ref", "libs", "libsignal_processing.so"))
    cmd = [sys.executable, "-m", "pytest", "-q", "--no-header", "-p", "no:cacheprovider", "-W", "ignore", "-rfE", "-k", deselect, *extra,
           *[os.path.join("tests", "python", f) for f in files]]
    res = subprocess.run(cmd, cwd=str(work), env=env, capture_output=True, text=True, timeout=1500)
    return res


def summary(res):
    m = re.search(r"(\d+) passed", res.stdout)
    f = re.search(r"(\d+) failed", res.stdout)
    return (int(m.group(1)) if m else 0), (int(f.group(1)) if f else 0)


@needs_reftests
def test_reference_test_rir(tmp_path):
    """tests/python/test_rir.py: the signal_processing / tools wrappers (translate, gaussian_filter, bad pixels, zstd, attributes ...)."""
    res = run_reference_tests(tmp_path, ["test_rir.py"])
    passed, failed = summary(res)
    assert res.returncode == 0 and failed == 0 and passed >= 50, res.stdout[-4000:] + res.stderr[-2000:]


@needs_reftests
def test_reference_test_irmovie_and_video_io(tmp_path):
    """tests/python/test_IRMovie.py and test_video_io.py: IRMovie.from_numpy_array (raw PCR file -> saver -> reader), slicing,
    timestamps, attributes, to_h264 / pcr2h264, the lossless and lossy recorders."""
    res = run_reference_tests(tmp_path, ["test_IRMovie.py", "test_video_io.py", "test_FileAttributes.py"])
    passed, failed = summary(res)
    assert res.returncode == 0 and failed == 0 and passed >= 200, res.stdout[-6000:] + res.stderr[-2000:]


@needs_reftests
def test_reference_test_registration(tmp_path):
    """tests/python/test_registration.py: the reference's MaskedRegistratorECC class (numpy + OpenCV) with its Gaussian, quantile
    and translate calls landing in this repo's library."""
    pytest.importorskip("cv2")
    res = run_reference_tests(tmp_path, ["test_registration.py"])
    passed, failed = summary(res)
    assert res.returncode == 0 and failed == 0 and passed >= 1, res.stdout[-6000:] + res.stderr[-2000:]


@needs_reftests
@pytest.mark.parametrize("script", ["ir_saver.py", "registration.py"])
def test_reference_example_scripts(tmp_path, script):
    """examples/ir_saver.py (lossless and lossy recording, reading back through IRMovie) and examples/registration.py of the
    reference, unmodified, on the drop-in libraries."""
    if script == "registration.py":
        pytest.importorskip("cv2")
    site = make_site(tmp_path)
    work = tmp_path / "work"
    shutil.copytree(os.path.join(REFTESTS, "examples"), work)
    env = dict(os.environ, PYTHONPATH=str(site), LIBRIR_DISABLE_JOBLIB="1",
               LIBRIR_B200_FORWARD_LIB=os.path.join(ROOT, "oracle", "_ref", "libs", "libsignal_processing.so"))
    res = subprocess.run([sys.executable, script], cwd=str(work), env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    if script == "ir_saver.py":
        assert "Number of images:  100" in res.stdout and "Compression factor is" in res.stdout, res.stdout[-2000:]
