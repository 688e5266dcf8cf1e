"""Host-side logic that needs no GPU: frame sharding, key-frame rule, regfile format."""
import numpy as np
import pytest

from librir_b200 import movie, video_io as vio


@pytest.mark.parametrize("nframes,world,gop", [(1000, 1, 50), (1000, 8, 50), (100000, 8, 50), (20000, 4, 50),
                                              (1234, 8, 50), (7, 8, 50), (0, 2, 50), (999, 3, 7)])
def test_shards_cover_movie_and_start_on_key_frames(nframes, world, gop):
    shards = [movie.shard_frames(nframes, world, r, gop) for r in range(world)]
    assert shards[0].start == 0 and shards[-1].stop == nframes
    for a, b in zip(shards, shards[1:]):
        assert a.stop == b.start
    key = vio.key_frames(nframes, gop) if nframes else np.zeros(0, np.uint8)
    for s in shards:
        assert s.nframes >= 0
        if s.nframes:
            assert s.start % gop == 0 and key[s.start] == 1
    sizes = [s.nframes for s in shards if s.nframes]
    if nframes >= world * gop:
        assert max(sizes) - min(sizes) <= 2 * gop  # one GOP of imbalance + a short last GOP


def test_shard_frames_rejects_nonsense():
    with pytest.raises(ValueError):
        movie.shard_frames(10, 0, 0)
    with pytest.raises(ValueError):
        movie.shard_frames(10, 2, 2)


def test_key_frames_match_oracle(port):
    for n, g in [(130, 50), (5, 1), (7, 3), (1, 50), (51, 50)]:
        np.testing.assert_array_equal(vio.key_frames(n, g), port.key_frames(n, g))


def test_linesize_is_32_byte_aligned():
    assert [vio.linesize(w) for w in (1, 32, 33, 640, 641)] == [32, 32, 64, 640, 672]


def test_regfile_round_trip(tmp_path):
    x = np.array([0.0, 1.25, -3.5, 100.125])
    y = np.array([0.5, -1.75, 2.0, -0.001])
    f = tmp_path / "shifts.regfile"
    vio.save_translation_file(str(f), x, y)
    lines = open(f).read().splitlines()
    assert lines[0].split("\t")[1:] == ["x-axis translations", "y-axis translations", "Confidence level"]
    assert len(lines) == 5 and len(lines[1].split("\t")) == 4
    gx, gy = vio.load_translation_file(str(f), nframes=4)
    np.testing.assert_allclose(gx, x.astype(np.float32))  # the reference parses float32
    np.testing.assert_allclose(gy, y.astype(np.float32))
    with pytest.raises(RuntimeError):
        vio.load_translation_file(str(f), nframes=5)
    bad = tmp_path / "bad.regfile"
    bad.write_text("h\n1\t2\t3\n")
    with pytest.raises(RuntimeError):
        vio.load_translation_file(str(bad))


def test_precoder_key_frame_bookkeeping_without_gpu(monkeypatch):
    calls = []
    monkeypatch.setattr(vio, "split_yuv444", lambda img, it=None, ls=None: (calls.append(1), 0, 0)[0:0] or (0, 0, 0))
    p = vio.LosslessPrecoder(8, 4, gop=3)
    keys = [p.add_image(np.zeros((4, 8), np.uint16))[0] for _ in range(8)]
    assert keys == [True, False, False, True, False, False, True, False]
    p = vio.LosslessPrecoder(8, 4, gop=3)
    it = np.zeros((4, 8), np.uint8)
    keys = [p.add_image(np.zeros((4, 8), np.uint16), it)[0] for _ in range(9)]  # IT overload uses '>' (h264.cpp:1165)
    assert keys == [True, False, False, False, True, False, False, False, True]
    with pytest.raises(RuntimeError):
        p.add_image(np.zeros((5, 8), np.uint16))


# ---------------------------------------------------------------------------------------------
# host entropy stage (zstd wrappers, tools.cpp:352-376)
# ---------------------------------------------------------------------------------------------
def test_zstd_wrappers_round_trip_and_garbage():
    """Mirrors the reference's own test (tests/python/test_rir.py:47-74): round trip, and garbage raises."""
    from librir_b200 import entropy as e

    rng = np.random.default_rng(0)
    for payload in (b"", b"toto" * 1000, rng.integers(0, 255, 10000, dtype=np.uint8).tobytes()):
        for level in (0, 1, 3, 19):
            c = e.zstd_compress(payload, level)
            assert e.zstd_decompress_bound(c) == len(payload)
            assert e.zstd_decompress(c) == payload
    assert e.zstd_compress_bound(1000) >= 1000
    with pytest.raises(RuntimeError):
        e.zstd_decompress(b"definitely not a zstd frame")
    a = np.arange(5000, dtype=np.uint16)
    assert np.array_equal(np.frombuffer(e.zstd_decompress(e.zstd_compress(a, 3)), np.uint16), a)


def test_zstd_wrappers_match_the_reference_build():
    """Byte-identical to the reference's zstd_compress (libtools of oracle/_ref) at the same level."""
    import ctypes as ct
    import os

    from librir_b200 import entropy as e

    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libs", "libtools.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref not built")
    tools = ct.CDLL(lib)
    tools.zstd_compress.restype = ct.c_longlong
    tools.zstd_compress.argtypes = [ct.c_char_p, ct.c_longlong, ct.c_char_p, ct.c_longlong, ct.c_int]
    tools.zstd_compress_bound.restype = ct.c_longlong
    tools.zstd_compress_bound.argtypes = [ct.c_longlong]
    payload = (np.arange(40000) % 977).astype(np.uint16).tobytes()
    for level in (1, 3, 9):
        cap = tools.zstd_compress_bound(len(payload))
        assert cap == e.zstd_compress_bound(len(payload))
        out = ct.create_string_buffer(cap)
        n = tools.zstd_compress(payload, len(payload), out, cap, level)
        assert n > 0 and out.raw[:n] == e.zstd_compress(payload, level)
