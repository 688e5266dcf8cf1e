"""The restatement (oracle.c) of the rows that live in the reference's video_io library -- writer split + key-frame rule
(a-7), lossy pre-conditioner (f-2), reader chain (a-3 / a-6 / f-1) -- against what the COMPILED reference did:

* tests/golden/vio_golden.npz, written by tests/golden/make_vio_golden.py from oracle/_ref/libs/libvideo_io.so (the
  reference's video_io sources over oracle/libav_stub.c) -- runs everywhere;
* the same library live, on other inputs, where oracle/_ref exists.

Bit-exact everywhere.  One set of pixels is excluded: bad pixels whose whole 3x3 window is flagged, where the reference
reads a stale stack slot (IRFileLoader.cpp:786-795; the golden file carries the mask)."""
import os

import numpy as np
import pytest

from oracle import refvio as rv
from tests import vio_cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vio_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "vio_golden.npz"))


def golden_planes(g, name, k):
    key = f"split_{name}_p{k}"
    if key in g:
        return g[key]
    return np.zeros(tuple(g[key + "_zero"]), dtype=np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# a-7
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", C.SPLIT_CASES, ids=[c[0] for c in C.SPLIT_CASES])
def test_split_and_key_frames_match_the_compiled_writer(port, vio_golden, case):
    name, t, h, w, gop, codec = case
    g = vio_golden
    mov = C.movie(t, h, w, seed=len(name) + t)
    W, H, fmt = (int(v) for v in g[f"split_{name}_dims"])
    pict = g[f"split_{name}_pict"]
    if codec == "h264":
        assert (W, H, fmt) == (w, h, rv.AV_PIX_FMT_YUV444P)
        # key-frame rule, h264.cpp:1050-1064: the writer marks pict_type I (1) on key frames, NONE (0) otherwise
        assert np.array_equal(pict == rv.AV_PICTURE_TYPE_I, port.key_frames(t, gop).astype(bool))
        y, u, v = (golden_planes(g, name, k) for k in range(3))
        assert not y.any()  # Y = 0 without an integration-time image (:1080)
        for i in range(t):
            py, pu, pv = port.split_444(mov[i], linesize=w)
            assert np.array_equal(pu, u[i]) and np.array_equal(pv, v[i]) and not py.any()
            img, _it = port.merge_444(y[i], u[i], v[i], w)
            assert np.array_equal(img, mov[i])
    else:
        # kvazaar branch (:846-862): width and height rounded up to 8, height doubled; pict_type never set
        w8, h8 = -(-w // 8) * 8, -(-h // 8) * 8
        assert (W, H, fmt) == (w8, h8 * 2 if h8 < 2 * h else h8, rv.AV_PIX_FMT_YUV420P)
        assert not pict.any()
        y = golden_planes(g, name, 0)
        assert not golden_planes(g, name, 1).any() and not golden_planes(g, name, 2).any()
        for i in range(t):
            py = port.split_420(mov[i], linesize=W)
            assert np.array_equal(py[:, :w], y[i][: 2 * h, :w])
            assert not y[i][2 * h:].any() and not y[i][:, w:].any()  # the frame is cleared first (:1090)
            assert np.array_equal(port.merge_420(np.ascontiguousarray(y[i][: 2 * h]), w), mov[i])


# ---------------------------------------------------------------------------------------------------------------------
# f-2
# ---------------------------------------------------------------------------------------------------------------------
def run_port_lossy(port, mov, stop, cfg, door, **kw):
    t, h, w = mov.shape
    st = port.lossy_open(w, h, stop, variant=1 if door == "add_loss" else 0, **C.lossy_params(cfg), **kw)
    outs, errs = [], []
    for i in range(t):
        o, e = port.lossy_add(st, mov[i])
        if door == "add_loss":
            o[stop:] = mov[i][stop:]  # addLoss copies back the lossy rows only (:2604)
        outs.append(o)
        errs.append(e)
    port.lossy_close(st)
    return np.stack(outs), np.array(errs)


@pytest.mark.parametrize("door", C.LOSSY_DOORS)
@pytest.mark.parametrize("cfg", C.LOSSY_CONFIGS, ids=[c[0] for c in C.LOSSY_CONFIGS])
def test_lossy_preconditioner_matches_the_compiled_saver(port, vio_golden, cfg, door):
    cname, params = cfg
    g = vio_golden
    t, h, w, stop = C.LOSSY_SHAPE
    mov = C.lossy_movie()
    outs, errs = run_port_lossy(port, mov, stop, params, door)
    key = f"lossy_{door}_{cname}"
    assert np.array_equal(errs[:, 0], g[key + "_low"]) and np.array_equal(errs[:, 1], g[key + "_high"])
    assert np.array_equal(outs[:: C.LOSSY_FULL_EVERY], g[key + "_full"])
    assert np.array_equal(np.array([C.crc(f) for f in outs], dtype=np.uint32), g[key + "_crc"])
    # the bounds really move in these runs, and the window of 40 spreads is full for most of them
    assert len(np.unique(g[key + "_low"])) >= 3


def test_lossy_overlapping_memcpy_is_what_the_compiled_reference_does(port, vio_golden):
    """h264.cpp:2347 shifts the window of spreads with an overlapping memcpy (undefined behaviour).  The compiled reference
    smears one .second value (oracle.c, orc_lossy_add_image); with the intended memmove the bounds differ from frame ~50 on."""
    t, h, w, stop = C.LOSSY_SHAPE
    mov = C.lossy_movie()
    _outs, errs = run_port_lossy(port, mov, stop, {}, "add_image_lossy", memcpy_quirk=False)
    g = vio_golden
    same = (errs[:, 0] == g["lossy_add_image_lossy_default_low"]) & (errs[:, 1] == g["lossy_add_image_lossy_default_high"])
    assert same[:42].all() and not same.all()


# ---------------------------------------------------------------------------------------------------------------------
# a-3 / a-6 / f-1
# ---------------------------------------------------------------------------------------------------------------------
def port_loader_movie(port, mov, codec, min_t, min_th, bp, mo, sx, sy, xy):
    t, h, w = mov.shape
    mth = min_th if min_th else h - 3  # IRFileLoader.cpp:918-921
    out = []
    for i in range(t):
        lo, hi = (mov[i] & 0xFF).astype(np.uint8), (mov[i] >> 8).astype(np.uint8)
        out.append(port.loader_read_image(lo, hi, xy=xy if bp else None, min_T=min_t, min_T_height=mth,
                                          shift=(sx[i], sy[i]) if mo else None))
    return np.stack(out)


@pytest.mark.parametrize("case", C.LOADER_CASES, ids=[c[0] for c in C.LOADER_CASES])
def test_reader_chain_matches_the_compiled_loader(port, vio_golden, case):
    name, t, h, w, codec, min_t, min_th = case
    g = vio_golden
    mov = C.loader_movie(case)
    sx, sy = C.shifts(t, seed=t)
    # the flagged set comes from readImage(0) -- min_T already added -- cropped to h - 3 rows (IRFileLoader.cpp:693-716)
    first = port_loader_movie(port, mov[:1], codec, min_t, min_th, 0, 0, sx, sy, None)[0]
    xy = port.bad_pixels_detect(first[: h - 3])[0]
    assert np.array_equal(xy, g[f"loader_{name}_xy"])
    undefined = np.unpackbits(g[f"loader_{name}_undefined"])[: h * w].reshape(h, w).astype(bool)
    for bp, mo in C.LOADER_MODES:
        got = port_loader_movie(port, mov, codec, min_t, min_th, bp, mo, sx, sy, xy)
        key = f"loader_{name}_{bp}{mo}"
        full = g[key + "_full"]
        if bp and undefined.any():
            # undefined pixels differ, and motion correction spreads them: compare what is defined, without motion
            if not mo:
                assert np.array_equal(got[::10][:, ~undefined], full[:, ~undefined])
            continue
        assert np.array_equal(got[::10], full)
        assert np.array_equal(np.array([C.crc(f) for f in got], dtype=np.uint32), g[key + "_crc"])


# ---------------------------------------------------------------------------------------------------------------------
# live, on other inputs
# ---------------------------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not rv.have_ref_vio(), reason="oracle/_ref/libs/libvideo_io.so not built")


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_live_lossy_both_doors(port, tmp_path, seed):
    rng = np.random.default_rng(seed)
    t, h, w = 220, int(rng.integers(8, 40)), int(rng.integers(8, 60))
    stop = int(rng.integers(max(5, h - 6), h + 1))
    mov = C.movie(t, h, w, seed=50 + seed, jump_at=int(rng.integers(45, 180)), ramp_every=int(rng.integers(3, 9)))
    cfg = dict(lowValueError=int(rng.integers(3, 12)), highValueError=int(rng.integers(0, 4)), runningAverage=int(rng.choice([0, 4, 32, 70])),
               subtractMin=int(rng.integers(0, 2)), removeBadPixels=int(rng.integers(0, 2)), stdFactor=float(rng.choice([1.5, 5.0])))
    for door in C.LOSSY_DOORS:
        s = rv.Saver(tmp_path / "l.bin", w, h, stop, **cfg)
        if door == "add_loss":
            ref = np.stack([s.add_loss(f) for f in mov])
        else:
            for i, f in enumerate(mov):
                s.add_image_lossy(f, i)
        lo_e, hi_e = s.low_errors(), s.high_errors()
        s.close()
        if door != "add_loss":
            d = rv.read_stub_file(tmp_path / "l.bin")
            ref = np.stack([r["planes"][1].astype(np.uint16) | (r["planes"][2].astype(np.uint16) << 8) for r in d["records"]])
        outs, errs = run_port_lossy(port, mov, stop, cfg, door)
        assert np.array_equal(errs[:, 0], lo_e) and np.array_equal(errs[:, 1], hi_e), (door, cfg)
        assert np.array_equal(outs, ref), (door, cfg)


@needs_ref
def test_live_writer_reader_round_trip_is_the_identity(tmp_path):
    """tests/python/test_IRMovie.py:46-49 of the reference, through its own compiled writer and reader."""
    mov = C.movie(60, 24, 40, seed=3)
    for codec in ("h264", "h265"):
        rv.write_lossless(tmp_path / "m.bin", mov, codec=codec, gop=7, timestamps=[i * 20_000_000 for i in range(60)])
        cam = rv.Camera(tmp_path / "m.bin")
        assert cam.count == 60 and (cam.h, cam.w) == (24, 40)
        for i in (0, 1, 2, 59, 30, 31, 5):
            assert np.array_equal(cam.load_image(i), mov[i])
        assert cam.image_time(3) == 60_000_000
        assert cam.global_attributes()["GOP"] == b"7"
        cam.close()
