"""File formats either side of the path: the attribute trailer (FileAttributes.cpp) and
the zstd movie file (ZFile.cpp).  Host code, so most of this runs without a GPU.

Three implementations meet here:
  product    librir_b200.tools (rirb_attrs_* / rirb_z_* in libsignal_processing_b200.so)
  port       oracle/container.py (byte-level restatement)
  reference  oracle/_ref (ZFile.cpp + tools library compiled from /root/reference), when it was built
and the golden files tests/golden/{zfile,attrs}_ref.bin, which the compiled reference wrote.
"""
import os

import numpy as np
import pytest

from librir_b200 import _lib, tools
from oracle import container as oc
from tests import container_cases as cc
from tests.conftest import ir_movie

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not oc.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def read(path):
    with open(path, "rb") as f:
        return f.read()


# ---- the oracle port against the golden files (pins the port) ------------------------------------
def test_port_reads_golden_zfile():
    frames, times, g = oc.read_zfile(os.path.join(GOLD, "zfile_ref.bin"))
    assert np.array_equal(frames, cc.golden_movie())
    assert np.array_equal(times, cc.golden_times())
    assert list(g) == [b"positions"]


def test_port_writes_golden_zfile_bytes(tmp_path):
    p = tmp_path / "port.bin"
    n = oc.write_zfile(p, cc.golden_movie(), cc.golden_times(), rate=50, clevel=2)
    gold = read(os.path.join(GOLD, "zfile_ref.bin"))
    mine = read(p)
    # headers, record framing and trailer layout are zstd-independent; the records match when libzstd is the reference's 1.5.5
    assert mine[:256] == gold[:256]
    assert np.array_equal(oc.read_zfile(p)[0], cc.golden_movie())
    if mine != gold:
        pytest.skip("libzstd here compresses differently from the one that wrote the golden file; content parity only")
    assert n == len(gold) - oc.parse_trailer(gold)[3]


def test_port_trailer_golden():
    g, frames, payload = cc.golden_attrs()
    gold = read(os.path.join(GOLD, "attrs_ref.bin"))
    G, F, T, size = oc.parse_trailer(gold)
    assert G == g and F == frames and np.array_equal(T, cc.golden_times())
    assert gold[: len(gold) - size] == payload
    mine = payload + oc.build_trailer(g, frames, cc.golden_times())
    # the only zstd-dependent bytes are the one compressed value
    assert oc.parse_trailer(mine)[:2] == (g, frames)
    if mine != gold:
        pytest.skip("libzstd here compresses differently from the one that wrote the golden file; content parity only")


# ---- the product against the golden files ------------------------------------------------------------
def test_product_reads_golden_zfile():
    with tools.ZFileReader(os.path.join(GOLD, "zfile_ref.bin")) as r:
        assert (r.width, r.height, len(r)) == (32, 24, 6)
        assert np.array_equal(r.timestamps, cc.golden_times())
        assert np.array_equal(r.read_images(), cc.golden_movie())
        assert np.array_equal(r.read_image(4), cc.golden_movie()[4])
        assert np.array_equal(r.read_images(2, 3), cc.golden_movie()[2:5])
        with pytest.raises(RuntimeError):
            r.read_images(5, 2)


@pytest.mark.parametrize("threads", [1, 4])
def test_product_writes_golden_zfile_bytes(tmp_path, threads):
    p = tmp_path / "prod.bin"
    with tools.ZFileWriter(p, 32, 24, rate=50, clevel=2, threads=threads) as w:
        w.add_images(cc.golden_movie()[:4], cc.golden_times()[:4])
        w.add_image(cc.golden_movie()[4], cc.golden_times()[4])
        w.add_images(cc.golden_movie()[5:], cc.golden_times()[5:])
        n = w.close()
    port = tmp_path / "port.bin"
    assert n == oc.write_zfile(port, cc.golden_movie(), cc.golden_times(), rate=50, clevel=2)
    assert read(p) == read(port)  # same libzstd on both sides: byte-identical files
    if read(port) == read(os.path.join(GOLD, "zfile_ref.bin")):
        assert read(p) == read(os.path.join(GOLD, "zfile_ref.bin"))


def test_product_reads_golden_attrs():
    g, frames, _ = cc.golden_attrs()
    fa = tools.FileAttributes.from_buffer(read(os.path.join(GOLD, "attrs_ref.bin")))
    assert fa.frame_count() == 6
    assert np.array_equal(fa.timestamps, cc.golden_times())
    assert {k.encode("utf8"): v for k, v in fa.attributes.items()} == g  # keys come back as str (ascii, else utf8)
    for i in range(6):
        assert {k.encode(): v for k, v in fa.frame_attributes(i).items()} == frames[i]
    h = fa.handle
    assert tools.attrs_frame_timestamp(h, 3) == cc.golden_times()[3]
    with pytest.raises(RuntimeError):
        tools.attrs_frame_timestamp(h, 6)
    with pytest.raises(RuntimeError):
        tools.attrs_frame_attribute_count(h, 6)
    _lib.load().rirb_attrs_abandon(h)
    fa.handle = 0


def test_product_writes_golden_attrs_bytes(tmp_path):
    g, frames, payload = cc.golden_attrs()
    p = tmp_path / "a.bin"
    p.write_bytes(payload)
    h = tools.attrs_open_file(p)
    assert h > 0 and tools.attrs_image_count(h) == 0
    tools.attrs_set_times(h, cc.golden_times())
    tools.attrs_set_global_attributes(h, g)
    for i, m in enumerate(frames):
        tools.attrs_set_frame_attributes(h, i, m)
    tools.attrs_close(h)
    mine = read(p)
    assert mine == payload + oc.build_trailer(g, frames, cc.golden_times())
    if mine != read(os.path.join(GOLD, "attrs_ref.bin")):
        pytest.skip("libzstd here compresses differently from the one that wrote the golden file; content parity only")


def test_trailer_rewrite_shrink_grow_and_flush(tmp_path):
    """writeIfDirty replaces the trailer in place: growing appends, shrinking truncates, the payload is untouched."""
    p = tmp_path / "a.bin"
    payload = b"\x01\x02" * 400
    p.write_bytes(payload)
    with tools.FileAttributes.from_filename(p) as fa:
        fa.timestamps = [1, 2, 3]
        fa.attributes = {"a": b"x" * 3000, "b": "text"}
        fa.set_frame_attributes(1, {"k": "v"})
    big = read(p)
    assert big.startswith(payload) and oc.parse_trailer(big)[0] == {b"a": b"x" * 3000, b"b": b"text"}
    with tools.FileAttributes.from_filename(p) as fa:
        assert fa.frame_attributes(1) == {"k": b"v"} and list(fa.timestamps) == [1, 2, 3]
        fa.timestamps = [7]
        fa.attributes = {}
        fa.flush()  # nothing set on the handle yet beyond what close() will push: flush must not corrupt the file
        assert read(p).startswith(payload)
    small = read(p)
    assert len(small) < len(big) and small.startswith(payload)
    G, F, T, size = oc.parse_trailer(small)
    assert G == {} and F == [{}] and list(T) == [7] and len(small) == len(payload) + size


def test_attrs_open_quirks(tmp_path):
    """FileAttributes::open creates a missing / too-short file; a file without trailer gets one on close;
    the in-memory variant needs a trailer (FileAttributes.cpp:268-272, 316-372)."""
    p = tmp_path / "new.bin"
    h = tools.attrs_open_file(p)
    assert h > 0 and os.path.exists(p)
    tools.attrs_close(h)
    assert oc.parse_trailer(read(p)) is not None and len(read(p)) == 38
    with pytest.raises(RuntimeError):
        tools.attrs_open_buffer(b"no trailer here, only forty bytes of text....")
    assert tools.attrs_open_file(tmp_path / "no" / "such" / "dir.bin") == 0
    lib = _lib.load()
    assert lib.rirb_attrs_image_count(12345) == -1 and lib.rirb_attrs_flush(12345) == -1


def test_damaged_trailer_is_refused(tmp_path):
    gold = bytearray(read(os.path.join(GOLD, "attrs_ref.bin")))
    size = oc.parse_trailer(bytes(gold))[3]
    start = len(gold) - size
    gold[start:start + 8] = (1 << 40).to_bytes(8, "little")  # global attribute count far beyond the bytes present
    with pytest.raises(RuntimeError):
        tools.attrs_open_buffer(bytes(gold))


def test_zfile_without_trailer_and_truncated(tmp_path):
    """A file whose writer died before close has no trailer and a zero sample count: the records are walked."""
    full = read(os.path.join(GOLD, "zfile_ref.bin"))
    size = oc.parse_trailer(full)[3]
    body = bytearray(full[: len(full) - size])
    body[128 + 16:128 + 24] = (0).to_bytes(8, "little")  # samples = 0
    p = tmp_path / "cut.bin"
    p.write_bytes(bytes(body))
    with tools.ZFileReader(p) as r:
        assert len(r) == 6 and np.array_equal(r.read_images(), cc.golden_movie())
    p.write_bytes(bytes(body[:-5]))  # last record incomplete
    with tools.ZFileReader(p) as r:
        assert len(r) == 5 and np.array_equal(r.read_images(), cc.golden_movie()[:5])
    p.write_bytes(b"not a movie" * 40)
    with pytest.raises(RuntimeError):
        tools.ZFileReader(p)
    if not _lib.device_available():  # methods 2 / 3 pre-code on the GPU; there is no CPU pre-coder in the product
        with pytest.raises(RuntimeError):
            tools.ZFileWriter(tmp_path / "m2.bin", 32, 24, method=2)


@pytest.mark.parametrize("shape,n", [((64, 80), 31), ((512, 640), 12)])
def test_product_port_round_trip(tmp_path, shape, n):
    mov = ir_movie(n, *shape)
    ts = np.arange(n, dtype=np.int64) * 1000 - 5
    p = tmp_path / "m.bin"
    with tools.ZFileWriter(p, shape[1], shape[0], rate=25, clevel=1, threads=0) as w:
        w.add_images(mov, ts)
    frames, times, _ = oc.read_zfile(p)
    assert np.array_equal(frames, mov) and np.array_equal(times, ts)
    q = tmp_path / "port.bin"
    oc.write_zfile(q, mov, ts, rate=25, clevel=1)
    assert read(p) == read(q)
    with tools.ZFileReader(q, threads=3) as r:
        assert np.array_equal(r.read_images(), mov) and np.array_equal(r.timestamps, ts)
    fa = tools.FileAttributes.from_filename(p)  # the trailer of a movie file is an ordinary attribute trailer
    assert fa.frame_count() == n and list(fa.attributes) == ["positions"] and len(fa.attributes["positions"]) == 8 * n
    fa.discard()
    assert read(p) == read(q)  # nothing was changed, so nothing is rewritten


# ---- against the compiled reference, live -------------------------------------------------------------
@needs_ref
def test_reference_reads_product_file(tmp_path):
    mov = ir_movie(140, 48, 64)  # > 125 frames: the positions attribute crosses the 1000-byte compression threshold
    ts = np.arange(140, dtype=np.int64) * 20_000_000
    p = tmp_path / "prod.bin"
    with tools.ZFileWriter(p, 64, 48, rate=50, clevel=2) as w:
        w.add_images(mov, ts)
        size = w.close()
    frames, times = oc.ref_read_zfile(p)
    assert np.array_equal(frames, mov) and np.array_equal(times, ts)
    q = tmp_path / "ref.bin"
    assert oc.ref_write_zfile(q, mov, ts, rate=50, clevel=2) == size
    assert read(p) == read(q)
    with tools.ZFileReader(q) as r:
        assert np.array_equal(r.read_images(), mov) and np.array_equal(r.timestamps, ts)


@needs_ref
def test_reference_and_product_attrs_agree(tmp_path):
    g, frames, payload = cc.golden_attrs()
    a, b = tmp_path / "a.bin", tmp_path / "b.bin"
    a.write_bytes(payload)
    b.write_bytes(payload)
    oc.ref_write_attrs(a, g, frames, cc.golden_times())
    h = tools.attrs_open_file(b)
    tools.attrs_set_times(h, cc.golden_times())
    tools.attrs_set_global_attributes(h, g)
    for i, m in enumerate(frames):
        tools.attrs_set_frame_attributes(h, i, m)
    tools.attrs_close(h)
    assert read(a) == read(b)
    G, F, T = oc.ref_read_attrs(b)
    assert G == g and F == frames and np.array_equal(T, cc.golden_times())
    # the reference edits a trailer the product wrote, the product reads it back
    oc.ref_write_attrs(b, {b"only": b"one"}, [{}, {b"f": b"1"}], [5, 6])
    fa = tools.FileAttributes.from_filename(b)
    assert fa.attributes == {"only": b"one"} and fa.frame_attributes(1) == {"f": b"1"} and list(fa.timestamps) == [5, 6]
    _lib.load().rirb_attrs_abandon(fa.handle)
    fa.handle = 0
    assert read(b).startswith(payload)


# ---- device-resident frames ------------------------------------------------------------------------------
@pytest.mark.gpu
def test_zfile_device_frames(tmp_path):
    import torch

    mov = ir_movie(40, 64, 96)
    ts = np.arange(40, dtype=np.int64)
    d = torch.from_numpy(mov.view(np.int16)).cuda().view(torch.uint16)
    p = tmp_path / "dev.bin"
    with tools.ZFileWriter(p, 96, 64) as w:
        w.add_images(d[:25], ts[:25])
        w.add_images(d[25:], ts[25:])
    assert np.array_equal(oc.read_zfile(p)[0], mov)
    out = torch.zeros_like(d)
    with tools.ZFileReader(p) as r:
        r.read_images(0, 40, out=out)
    assert torch.equal(out.view(torch.int16), d.view(torch.int16))


# ---- randomised: product, port and (when built) the compiled reference on the same random content -------------------
FUZZ_SEED = int(os.environ.get("RIRB_FUZZ_SEED", "0"))
FUZZ_SCALE = max(1, int(os.environ.get("RIRB_FUZZ_SCALE", "1")))


def _random_attrs(rng, nframes):
    def blob(maxlen):
        n = int(rng.choice([0, 1, 7, 999, 1000, 1001, 1500, maxlen]))
        kind = rng.integers(0, 3)
        if kind == 0:
            return bytes(rng.integers(0, 256, n, dtype=np.uint8))            # incompressible
        if kind == 1:
            return (b"abc" * (n // 3 + 1))[:n]                                # compressible
        return bytes(rng.integers(0, 4, n, dtype=np.uint8))                   # compressible, with NULs

    g = {}
    for i in range(int(rng.integers(0, 6))):
        g[b"k%d" % i + bytes(rng.integers(97, 123, int(rng.integers(0, 5)), dtype=np.uint8))] = blob(4000)
    frames = []
    for _ in range(nframes):
        frames.append({b"f%d" % j: blob(1200) for j in range(int(rng.integers(0, 3)))})
    return g, frames


def test_fuzz_trailers_product_vs_port_vs_reference(tmp_path):
    rng = np.random.default_rng(900 + FUZZ_SEED)
    for case in range(25 * FUZZ_SCALE):
        n = int(rng.integers(0, 12))
        times = rng.integers(-2 ** 62, 2 ** 62, n, dtype=np.int64)
        g, frames = _random_attrs(rng, n)
        payload = bytes(rng.integers(0, 256, int(rng.choice([0, 5, 29, 30, 31, 600])), dtype=np.uint8))
        a = tmp_path / f"a{case}.bin"
        a.write_bytes(payload)
        h = tools.attrs_open_file(a)
        assert h > 0
        tools.attrs_set_times(h, times)
        tools.attrs_set_global_attributes(h, g)
        for i, m in enumerate(frames):
            tools.attrs_set_frame_attributes(h, i, m)
        tools.attrs_close(h)
        # FileAttributes::open truncates a file shorter than a trailer (30 bytes) before appending (FileAttributes.cpp:364-371)
        kept = payload if len(payload) >= 30 else b""
        assert read(a) == kept + oc.build_trailer(g, frames, times), f"case {case}"
        G, F, T, _ = oc.parse_trailer(read(a))
        assert G == g and F == frames and np.array_equal(T, times)
        fa = tools.FileAttributes.from_buffer(read(a)) if len(read(a)) else None
        assert fa is not None and fa.frame_count() == n
        assert [fa.frame_attributes(i) for i in range(n)] == [{k.decode(): v for k, v in m.items()} for m in frames]
        _lib.load().rirb_attrs_abandon(fa.handle)
        fa.handle = 0
        if oc.have_ref():
            b = tmp_path / f"b{case}.bin"
            b.write_bytes(payload)
            oc.ref_write_attrs(b, g, frames, times)
            assert read(a) == read(b), f"case {case}: differs from the compiled reference"


def test_fuzz_zfiles_product_vs_port_vs_reference(tmp_path):
    rng = np.random.default_rng(901 + FUZZ_SEED)
    for case in range(10 * FUZZ_SCALE):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 90))
        n = int(rng.integers(0, 150 if case % 5 == 0 else 20))
        kind = rng.integers(0, 3)
        mov = (rng.integers(0, 65536, (n, h, w), dtype=np.uint16) if kind == 0 else
               np.full((n, h, w), int(rng.integers(0, 65536)), np.uint16) if kind == 1 else
               (rng.integers(8000, 8100, (n, h, w)) + np.arange(w)).astype(np.uint16))
        ts = np.sort(rng.integers(0, 2 ** 60, n, dtype=np.int64))
        clevel = int(rng.choice([1, 2, 3, 9]))
        p = tmp_path / f"p{case}.bin"
        with tools.ZFileWriter(p, w, h, rate=int(rng.integers(1, 999)), clevel=clevel, threads=int(rng.choice([0, 1, 3]))) as wr:
            k = int(rng.integers(0, n + 1))
            if k:
                wr.add_images(mov[:k], ts[:k])
            if n - k:
                wr.add_images(mov[k:], ts[k:])
        frames, times, _ = oc.read_zfile(p)
        assert np.array_equal(frames, mov) and np.array_equal(times, ts), f"case {case}"
        with tools.ZFileReader(p, threads=int(rng.choice([0, 1, 5]))) as rd:
            assert len(rd) == n and np.array_equal(rd.timestamps, ts)
            if n:
                a, b = sorted(rng.integers(0, n + 1, 2))
                assert np.array_equal(rd.read_images(a, b - a), mov[a:b])
        if oc.have_ref() and n:
            f2, t2 = oc.ref_read_zfile(p)
            assert np.array_equal(f2, mov) and np.array_equal(t2, ts)
