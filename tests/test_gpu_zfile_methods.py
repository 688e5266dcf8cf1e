"""The zstd movie file with the pre-coder in front (methods 2 and 3 of video_io.h:298-305, defined by this repo: the reference
never implemented them): round trips, random access, GOPs that straddle calls, device pointers.  Method 1 stays byte-identical
to the reference's files (tests/test_container.py)."""
import os

import numpy as np
import pytest

from tests import vio_cases as C

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from librir_b200 import _lib, tools  # noqa: E402


def to_dev(a):
    return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)


@pytest.mark.parametrize("method,gop", [(2, 50), (3, 50), (3, 7), (3, 1)])
@pytest.mark.parametrize("shape", [(130, 40, 48), (33, 19, 24), (5, 64, 80)])
def test_round_trip_in_pieces(tmp_path, method, gop, shape):
    t, h, w = shape
    mov = C.movie(t, h, w, seed=method * 10 + gop)
    ts = np.arange(t, dtype=np.int64) * 1000 + 5
    fn = str(tmp_path / "m.bin")
    lib = _lib.load()
    z = lib.rirb_z_open_file_write_gop(fn.encode(), w, h, 50, method, 3, gop)
    assert z > 0, _lib.last_error()
    # pieces that do not respect GOP boundaries, host and device pointers mixed
    cuts = [0, 1, 2, min(t, 2 + gop + 3), min(t, 2 * gop + 9), t]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        part = np.ascontiguousarray(mov[a:b])
        src = to_dev(part) if (a % 2) else part
        ptr = src.data_ptr() if hasattr(src, "data_ptr") else part.ctypes.data
        assert lib.rirb_z_write_images(z, ptr, b - a, ts[a:b].ctypes.data, 0) == 0, _lib.last_error()
    assert lib.rirb_z_close_file(z) > 256
    r = tools.ZFileReader(fn)
    assert r.images == t and (r.height, r.width) == (h, w)
    assert np.array_equal(r.timestamps, ts)
    assert np.array_equal(r.read_images(0, t), mov)
    for i in (t - 1, 0, t // 2, min(t - 1, gop), min(t - 1, gop + 1), 1):  # random access: decoding restarts at the GOP's key frame
        assert np.array_equal(r.read_image(i), mov[i]), i
    out = torch.empty((min(t, 11), h, w), dtype=torch.uint16, device="cuda")
    assert lib.rirb_z_read_images(r.handle, t - out.shape[0], out.shape[0], out.data_ptr(), None, 0) == 0
    assert np.array_equal(out.cpu().view(torch.int16).numpy().view(np.uint16), mov[t - out.shape[0]:])
    r.close()


@pytest.mark.parametrize("method", [2, 3])
def test_reads_longer_than_one_pass(tmp_path, method):
    """A read that spans several passes of the reader (a pass is ~64 MB of frames) into DEVICE memory: the next pass's zstd
    threads must not overwrite the pinned planes the previous pass is still uploading (scripts/zfile_bench.py caught it)."""
    t, h, w = 230, 512, 640
    rng = np.random.default_rng(method)
    base = C.movie(10, h, w, seed=method)
    mov = base[rng.integers(0, 10, t)] + rng.integers(0, 4, (t, 1, 1)).astype(np.uint16)
    ts = np.arange(t, dtype=np.int64)
    fn = str(tmp_path / "long.bin")
    with tools.ZFileWriter(fn, w, h, method=method, clevel=1, gop=50) as wr:
        wr.add_images(mov, ts)
    d = torch.empty((t, h, w), dtype=torch.uint16, device="cuda")
    with tools.ZFileReader(fn) as r:
        for _ in range(2):
            d.zero_()
            r.read_images(0, t, out=d)
            assert np.array_equal(d.cpu().view(torch.int16).numpy().view(np.uint16), mov)
        assert np.array_equal(r.read_images(3, t - 7), mov[3:t - 4])  # host destination, not aligned to a pass
    os.remove(fn)


def test_precoder_shrinks_the_file(tmp_path):
    mov = C.movie(200, 128, 160, seed=4)
    sizes = {}
    for method in (1, 2, 3):
        fn = str(tmp_path / f"m{method}.bin")
        lib = _lib.load()
        z = lib.rirb_z_open_file_write_gop(fn.encode(), 160, 128, 50, method, 3, 50)
        ts = np.arange(200, dtype=np.int64)
        assert lib.rirb_z_write_images(z, mov.ctypes.data, 200, ts.ctypes.data, 0) == 0
        sizes[method] = lib.rirb_z_close_file(z)
    assert sizes[3] < sizes[2] < sizes[1], sizes


def test_bad_arguments():
    lib = _lib.load()
    assert lib.rirb_z_open_file_write_gop(b"/tmp/x.bin", 64, 48, 50, 4, 3, 50) == 0 and "method" in _lib.last_error()
    assert lib.rirb_z_open_file_write_gop(b"/tmp/x.bin", 64, 48, 50, 3, 3, 0) == 0
