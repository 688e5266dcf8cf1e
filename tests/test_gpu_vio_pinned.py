"""The CUDA path against what the COMPILED reference's video_io did (rows a-3, a-6, a-7, f-1, f-2), through the C ABI:

* tests/golden/vio_golden.npz (made by tests/golden/make_vio_golden.py from oracle/_ref/libs/libvideo_io.so): the planes
  and key-frame decisions of H264Capture::AddFrame, both doors of the lossy pre-conditioner over 220 frames x 5
  configurations with per-frame CRCs, IRFileLoader::readImage with bad pixels and motion correction;
* the same compiled library LIVE on other inputs (oracle/_ref travels to the GPU box).

Bit-exact everywhere."""
import os

import numpy as np
import pytest

from tests import vio_cases as C
from tests.test_oracle_vs_refvio import golden_planes

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from librir_b200 import _lib, video_io as vio  # noqa: E402
from oracle import refvio as rv  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vio_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "vio_golden.npz"))


def to_dev(a):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(a).cuda()


def to_host(t):
    if isinstance(t, np.ndarray):
        return t
    if t.dtype == torch.uint16:
        return t.cpu().view(torch.int16).numpy().view(np.uint16)
    return t.cpu().numpy()


# ---- a-7 --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", C.SPLIT_CASES, ids=[c[0] for c in C.SPLIT_CASES])
def test_split_planes_and_key_frames_equal_the_compiled_writer(vio_golden, case):
    name, t, h, w, gop, codec = case
    g = vio_golden
    mov = C.movie(t, h, w, seed=len(name) + t)
    W, H, _fmt = (int(v) for v in g[f"split_{name}_dims"])
    if codec == "h264":
        assert np.array_equal(vio.key_frames(t, gop).astype(bool), g[f"split_{name}_pict"] == rv.AV_PICTURE_TYPE_I)
        u, v = golden_planes(g, name, 1), golden_planes(g, name, 2)
        for i in range(t):
            ls = vio.linesize(w)
            py, pu, pv = vio.split_yuv444(mov[i], None, ls)
            assert np.array_equal(pu[:, :w], u[i]) and np.array_equal(pv[:, :w], v[i]) and not py[:, :w].any()
            img, _it = vio.merge_yuv444(py, pu, pv, w)
            assert np.array_equal(img, mov[i])
        # the dense movie entry: lo / hi planes are U / V
        lo, hi = vio.precode_movie(to_dev(mov), gop, False)
        assert np.array_equal(to_host(lo), u) and np.array_equal(to_host(hi), v)
    else:
        y = golden_planes(g, name, 0)
        for i in range(t):
            py = vio.split_yuv420(mov[i], None, W)
            assert np.array_equal(py[:, :w], y[i][: 2 * h, :w]) and not py[:, w:].any()
            assert np.array_equal(vio.merge_yuv420(np.ascontiguousarray(y[i][: 2 * h]), w), mov[i])


# ---- f-2 --------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["one launch per run of frames", "three launches per frame"])
def lossy_driver(request):
    _lib.set_parameter("lossy_run", request.param == "one launch per run of frames")
    yield request.param
    _lib.set_parameter("lossy_run", 1)


def gpu_lossy(mov, stop, cfg, door, chunks, device):
    t, h, w = mov.shape
    pre = vio.LossyPreconditioner(w, h, stop, variant=door, **cfg)
    outs, errs, a = [], [], 0
    for c in chunks:
        part = mov[a:a + c]
        o, e = pre.add_images(to_dev(part) if device else part)
        outs.append(to_host(o))
        errs.append(e)
        a += c
    assert a == t
    return np.concatenate(outs), np.concatenate(errs)


@pytest.mark.parametrize("door", C.LOSSY_DOORS)
@pytest.mark.parametrize("cfg", C.LOSSY_CONFIGS, ids=[c[0] for c in C.LOSSY_CONFIGS])
def test_lossy_preconditioner_equals_the_compiled_saver(vio_golden, cfg, door, lossy_driver):
    cname, params = cfg
    g = vio_golden
    t, h, w, stop = C.LOSSY_SHAPE
    mov = C.lossy_movie()
    # runs of frames that straddle the 40-frame window, the 64-frame launch limit and single-frame calls
    outs, errs = gpu_lossy(mov, stop, params, door, [1, 1, 37, 70, 1, 100, 10], device=True)
    key = f"lossy_{door}_{cname}"
    assert np.array_equal(errs[:, 0], g[key + "_low"]) and np.array_equal(errs[:, 1], g[key + "_high"])
    assert np.array_equal(outs[:: C.LOSSY_FULL_EVERY], g[key + "_full"])
    crcs = np.array([C.crc(f) for f in outs], dtype=np.uint32)
    bad = np.nonzero(crcs != g[key + "_crc"])[0]
    assert bad.size == 0, f"first differing frame {bad[:5]}"


def test_lossy_memmove_switch_changes_the_bounds(vio_golden):
    t, h, w, stop = C.LOSSY_SHAPE
    mov = C.lossy_movie()
    _outs, errs = gpu_lossy(mov, stop, dict(memcpyQuirk=False), "add_image_lossy", [t], device=True)
    g = vio_golden
    same = (errs[:, 0] == g["lossy_add_image_lossy_default_low"]) & (errs[:, 1] == g["lossy_add_image_lossy_default_high"])
    assert same[:42].all() and not same.all()


# ---- a-3 / a-6 / f-1 --------------------------------------------------------------------------------------------------
@pytest.fixture(params=["two passes", "fused"])
def loader_kernel(request):
    _lib.set_parameter("loader_fused", request.param == "fused")
    yield request.param
    _lib.set_parameter("loader_fused", 0)


@pytest.mark.parametrize("case", C.LOADER_CASES, ids=[c[0] for c in C.LOADER_CASES])
def test_reader_chain_equals_the_compiled_loader(vio_golden, case, loader_kernel):
    name, t, h, w, codec, min_t, min_th = case
    g = vio_golden
    mov = C.loader_movie(case)
    sx, sy = C.shifts(t, seed=t)
    lo, hi = (mov & 0xFF).astype(np.uint8), (mov >> 8).astype(np.uint8)
    undefined = np.unpackbits(g[f"loader_{name}_undefined"])[: h * w].reshape(h, w).astype(bool)
    first = vio.read_movie(lo[:1], hi[:1], None, min_t, min_th)[0]  # readImage(0) before bad pixels are on
    bp = vio.LoaderBadPixels(first)
    for use_bp, mo in C.LOADER_MODES:
        got = to_host(vio.read_movie(to_dev(lo), to_dev(hi), bp if use_bp else None, min_t, min_th,
                                     sx.astype(np.float64) if mo else None, sy.astype(np.float64) if mo else None))
        key = f"loader_{name}_{use_bp}{mo}"
        full = g[key + "_full"]
        if use_bp and undefined.any():
            if not mo:
                assert np.array_equal(got[::10][:, ~undefined], full[:, ~undefined])
            continue
        assert np.array_equal(got[::10], full), key
        assert np.array_equal(np.array([C.crc(f) for f in got], dtype=np.uint32), g[key + "_crc"]), key


# ---- live -------------------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not rv.have_ref_vio(), reason="oracle/_ref/libs/libvideo_io.so not built")


@needs_ref
@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15])
def test_live_lossy_against_the_compiled_saver(tmp_path, seed):
    rng = np.random.default_rng(seed)
    t, h, w = 230, int(rng.integers(8, 70)), int(rng.integers(8, 90))
    stop = int(rng.integers(max(5, h - 6), h + 1))
    mov = C.movie(t, h, w, seed=500 + seed, jump_at=int(rng.integers(45, 200)), ramp_every=int(rng.integers(3, 9)))
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
        outs, errs = gpu_lossy(mov, stop, cfg, door, [90, 1, 79, 60], device=bool(seed & 1))
        assert np.array_equal(errs[:, 0], lo_e) and np.array_equal(errs[:, 1], hi_e), (door, cfg)
        assert np.array_equal(outs, ref), (door, cfg)


@needs_ref
@pytest.mark.parametrize("seed", [21, 22])
def test_live_reader_against_the_compiled_loader(tmp_path, seed):
    """A file written by the reference's writer, read by the reference's reader with bad pixels and motion correction on,
    against rirb_loader_read_movie fed with the planes the reference's writer recorded."""
    rng = np.random.default_rng(seed)
    t, h, w = 30, int(rng.integers(12, 80)), int(rng.integers(8, 100))
    codec = ("h264", "h265")[seed & 1]
    if codec == "h265":
        # the reference's reader overruns its integration-time image when the kvazaar padding changes the width
        # (IT[i + y * frame->width], h264.cpp:3046, with frame->width = width rounded up to 8): keep it a multiple of 8
        w = -(-w // 8) * 8
    mov = C.movie(t, h, w, seed=900 + seed, n_bad_frac=4e-3)
    min_t = int(rng.integers(0, 500))
    rv.write_lossless(tmp_path / "m.bin", mov, codec=codec, gop=10, global_attrs={"MIN_T": str(min_t)} if min_t else None)
    sx, sy = C.shifts(t, seed=seed, amp=4.0)
    rv.write_regfile(tmp_path / "m.regfile", sx, sy)
    cam = rv.Camera(tmp_path / "m.bin")
    cam.enable_bad_pixels(True)
    cam.load_motion_correction_file(tmp_path / "m.regfile")
    cam.enable_motion_correction(True)
    ref = np.stack([cam.load_image(i) for i in range(t)])
    cam.close()
    d = rv.read_stub_file(tmp_path / "m.bin")
    if codec == "h264":
        lo = np.stack([r["planes"][1] for r in d["records"]])
        hi = np.stack([r["planes"][2] for r in d["records"]])
    else:
        lo = np.stack([np.ascontiguousarray(r["planes"][0][:h, :w]) for r in d["records"]])
        hi = np.stack([np.ascontiguousarray(r["planes"][0][h:2 * h, :w]) for r in d["records"]])
    first = vio.read_movie(lo[:1], hi[:1], None, min_t, 0)[0]
    bp = vio.LoaderBadPixels(first)
    x, y = vio.load_translation_file(tmp_path / "m.regfile", t)
    got = to_host(vio.read_movie(to_dev(lo), to_dev(hi), bp, min_t, 0, x, y))
    assert np.array_equal(got, ref)
