"""Randomised parity of the remaining entries (bad pixels, reader chain, pre-coder, statistics, lossy pre-conditioner)
against the oracle, through the C ABI.  The committed case counts are small; on the GPU box
``RIRB_FUZZ_SEED=<n> RIRB_FUZZ_SCALE=<k> pytest tests/test_gpu_fuzz.py -m gpu`` runs k times as many cases from other
seeds (the translate / Gaussian fuzzers live in test_gpu_parity.py and take the same variables).
"""
import os

import numpy as np
import pytest

from tests.conftest import ir_frame, ir_movie

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

from librir_b200 import movie  # noqa: E402
from librir_b200 import signal_processing as sp  # noqa: E402
from librir_b200 import video_io as vio  # noqa: E402

SEED = int(os.environ.get("RIRB_FUZZ_SEED", "0"))
SCALE = max(1, int(os.environ.get("RIRB_FUZZ_SCALE", "1")))
WIDTHS = [1, 2, 3, 5, 7, 8, 15, 16, 17, 24, 31, 32, 33, 40, 48, 63, 64, 65, 96, 127, 128, 129, 136, 144, 160, 200, 256, 264, 272]


def to_dev(a):
    return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)


@pytest.fixture(scope="module")
def best():
    from oracle import oracle as O

    return O.best()


def random_frame(rng, h, w):
    """Includes frames with pixels more than 46,340 counts from the median, where the reference's int product wraps
    (the restatement and the library reproduce the x86-64 behaviour, pinned by tests/golden/bp_extreme_golden.npz)."""
    kind = rng.integers(0, 4)
    if kind == 0:
        return rng.integers(0, 65536, (h, w), dtype=np.uint16)
    if kind == 1:
        return rng.integers(7000, 9000, (h, w), dtype=np.uint16)
    f = ir_frame(h, w, int(rng.integers(0, 1 << 30))) if h >= 8 and w >= 8 else rng.integers(0, 16384, (h, w), dtype=np.uint16)
    if kind == 3:  # clusters of stuck pixels, borders included
        n = max(1, h * w // 30)
        f = f.copy()
        f.reshape(-1)[rng.choice(h * w, n, replace=False)] = rng.choice([0, 65535, 16000])
        f[0, :] = rng.choice([0, f[0, 0]])
    return f


def test_fuzz_bad_pixels(port):
    """Detection list + clamp value and the corrected frames, bit for bit, on random shapes and contents."""
    rng = np.random.default_rng(501 + SEED)
    for case in range(30 * SCALE):
        w = int(rng.choice(WIDTHS))
        h = int(rng.integers(1, 90))
        first = random_frame(rng, h, w)
        oxy, _thr, oclamp = port.bad_pixels_detect(first)
        handle = sp.bad_pixels_create(first)
        assert handle > 0
        xy, clamp = sp.bad_pixels_list(handle)
        np.testing.assert_array_equal(xy, oxy, err_msg=f"case {case}: {h}x{w} list")
        assert clamp == oclamp, f"case {case}: {h}x{w} clamp"
        mov = np.stack([random_frame(rng, h, w) for _ in range(3)])
        want = np.stack([port.bad_pixels_correct_with(oxy, oclamp, f) for f in mov])
        np.testing.assert_array_equal(sp.bad_pixels_correct_batch(handle, mov), want, err_msg=f"case {case}: {h}x{w} correct")
        np.testing.assert_array_equal(sp.bad_pixels_correct(handle, mov[1]), want[1])
        sp.bad_pixels_destroy(handle)


def test_fuzz_reader_chain(port):
    """merge -> + min_T -> loader medians -> motion, each stage on or off, random sizes / shifts / offsets."""
    from librir_b200 import _lib

    rng = np.random.default_rng(502 + SEED)
    for case in range(24 * SCALE):
        w = int(rng.choice([w for w in WIDTHS if w >= 3]))
        h = int(rng.integers(6, 100))
        n = int(rng.integers(1, 5))
        mov = np.stack([random_frame(rng, h, w) for _ in range(n)])
        lo, hi = (mov & 0xFF).astype(np.uint8), (mov >> 8).astype(np.uint8)
        big = rng.integers(0, 3) == 0
        sx = rng.uniform(-w - 3, w + 3, n) if big else rng.uniform(-3, 3, n)
        sy = rng.uniform(-h - 3, h + 3, n) if big else rng.uniform(-3, 3, n)
        if rng.integers(0, 3) == 0:
            sx, sy = np.round(sx), np.round(sy * 2) / 2
        use_bp, use_motion = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        min_T = int(rng.choice([0, 273, -7, 40000, 65535]))
        rows = int(rng.integers(0, h + 1))
        bp = vio.LoaderBadPixels(mov[0]) if use_bp else None
        xy = sp.bad_pixels_list(bp.handle)[0] if bp is not None else None
        _lib.set_parameter("loader_fused", int(rng.integers(0, 2)))
        try:
            got = vio.read_movie(lo, hi, bp, min_T, rows, sx if use_motion else None, sy if use_motion else None)
        finally:
            _lib.set_parameter("loader_fused", 0)
        for t in range(n):
            # MIN_T_HEIGHT == 0 is "absent": the loader makes it height - 3 when it opens the file (IRFileLoader.cpp:918-921)
            want = port.loader_read_image(lo[t], hi[t], xy, min_T, rows if rows else h - 3, (sx[t], sy[t]) if use_motion else None)
            np.testing.assert_array_equal(got[t], want, err_msg=f"case {case}: {n}x{h}x{w} bp={use_bp} motion={use_motion} "
                                          f"min_T={min_T} rows={rows} shift=({sx[t]!r},{sy[t]!r})")


def test_fuzz_precoder_and_stats(port):
    """Byte planes (+ temporal delta), their inverse, and the fused statistics, on random shapes / GOPs / offsets."""
    rng = np.random.default_rng(503 + SEED)
    for case in range(24 * SCALE):
        w = int(rng.choice(WIDTHS))
        h = int(rng.integers(1, 70))
        n = int(rng.integers(1, 40))
        gop = int(rng.choice([1, 2, 3, 5, 7, 50]))
        delta = bool(rng.integers(0, 2))
        mov = rng.integers(0, 65536, (n, h, w), dtype=np.uint16) if rng.integers(0, 2) else np.stack(
            [random_frame(rng, h, w) for _ in range(n)])
        first = int(rng.integers(0, 4)) * gop
        lo, hi = vio.precode_movie(mov, gop, delta, first)
        olo, ohi = port.precode_movie(mov, gop, delta)  # the key-frame pattern repeats every gop: any aligned offset gives the same planes
        np.testing.assert_array_equal(lo, olo, err_msg=f"case {case}: {n}x{h}x{w} gop {gop} delta {delta}")
        np.testing.assert_array_equal(hi, ohi)
        np.testing.assert_array_equal(vio.decode_movie(lo, hi, gop, delta, first), mov)
        st = movie.MovieStats("cuda")
        dlo, dhi = vio.precode_movie(to_dev(mov), gop, delta, first, stats=st)
        np.testing.assert_array_equal(dlo.cpu().numpy(), olo)
        np.testing.assert_array_equal(dhi.cpu().numpy(), ohi)
        mn, mx, hist = port.movie_stats(mov)
        assert (st.min(), st.max(), st.count) == (mn, mx, mov.size), f"case {case}"
        np.testing.assert_array_equal(st.histogram(), hist)


def test_fuzz_quantiles(port):
    rng = np.random.default_rng(504 + SEED)
    for case in range(40 * SCALE):
        w = int(rng.choice(WIDTHS))
        h = int(rng.integers(1, 60))
        img = random_frame(rng, h, w)
        pc = float(rng.choice([0.0, 1.0, 0.5, 1e-6, 0.999999, float(np.float32(rng.random()))]))
        assert sp.find_median_pixel(img, pc) == port.find_median_pixel(img, pc), f"case {case}: {h}x{w} {pc}"
        mask = (rng.random((h, w)) < rng.choice([0.0, 0.1, 0.5, 1.0])).astype(np.uint8)
        assert sp.find_median_pixel(img, pc, mask) == port.find_median_pixel(img, pc, mask), f"case {case}: masked {h}x{w} {pc}"


def test_fuzz_lossy_preconditioner(port):
    rng = np.random.default_rng(505 + SEED)
    for case in range(4 * SCALE):
        w = int(rng.choice([16, 40, 64, 96, 136]))
        h = int(rng.integers(8, 60))
        n = int(rng.integers(5, 70))
        stop = int(rng.integers(1, h + 1))
        cfg = dict(low_error=int(rng.integers(0, 12)), high_error=int(rng.integers(0, 6)), std_factor=float(rng.choice([2.0, 5.0, 9.5])),
                   running_average=int(rng.choice([0, 1, 8, 32, 64])), subtract_min=bool(rng.integers(0, 2)),
                   bp_enabled=bool(rng.integers(0, 2)))
        mov = ir_movie(n, h, w, seed=int(rng.integers(0, 1 << 20)), drift=0.7)
        mov = (mov.astype(np.int64) + rng.integers(-40, 40, mov.shape)).clip(0, 16383).astype(np.uint16)
        state = port.lossy_open(w, h, stop, **cfg)
        want, werr = [], []
        for f in mov:
            o, e = port.lossy_add(state, f)
            want.append(o)
            werr.append(e)
        port.lossy_close(state)
        pre = vio.LossyPreconditioner(w, h, stop, lowValueError=cfg["low_error"], highValueError=cfg["high_error"],
                                      stdFactor=cfg["std_factor"], runningAverage=cfg["running_average"], subtractMin=cfg["subtract_min"],
                                      removeBadPixels=cfg["bp_enabled"])
        k = int(rng.integers(1, n))
        from librir_b200 import _lib

        _lib.set_parameter("lossy_run", int(rng.integers(0, 2)))  # the two drivers may alternate on one handle
        out_a, err_a = pre.add_images(mov[:k])
        _lib.set_parameter("lossy_run", int(rng.integers(0, 2)))
        out_b, err_b = pre.add_images(mov[k:])
        _lib.set_parameter("lossy_run", 1)
        np.testing.assert_array_equal(np.concatenate([out_a, out_b]), np.stack(want), err_msg=f"case {case}: {n}x{h}x{w} {cfg} stop {stop}")
        np.testing.assert_array_equal(np.concatenate([err_a, err_b]), np.array(werr))


def test_fuzz_translate_all_dtypes(best, port):
    """The generic translate kernel (every numpy dtype the facade accepts) on random shapes, shifts and strategies,
    bit for bit against the compiled reference (the restatement where the reference's read leaves its buffer)."""
    from tests.golden.make_golden import DTYPES, typed_image
    from tests.test_gpu_parity import reference_reads_past_the_buffer

    rng = np.random.default_rng(506 + SEED)
    specials = [0.0, 1.0, -1.0, 0.5, -0.5, 7.9999995, -7.9999995, 1.0000001, -2.9999998, 1e-20, 0.25, 15.999999]
    for case in range(60 * SCALE):
        dt = DTYPES[int(rng.integers(0, len(DTYPES)))]
        w = int(rng.choice([w for w in WIDTHS if w <= 144]))
        h = int(rng.integers(1, 60))
        img = typed_image(dt, h, w, rng)
        k = rng.integers(0, 3)
        dx = float(rng.choice(specials)) if k == 0 else float(np.float32(rng.uniform(-w - 2, w + 2) if k == 1 else rng.uniform(-3, 3)))
        k = rng.integers(0, 3)
        dy = float(rng.choice(specials)) if k == 0 else float(np.float32(rng.uniform(-h - 2, h + 2) if k == 1 else rng.uniform(-3, 3)))
        st = ["nearest", "background", "wrap", ""][case % 4]
        got, want = sp.translate(img, dx, dy, st, 1), best.translate(img, dx, dy, st, 1)
        undefined = reference_reads_past_the_buffer(h, w, dx, dy)
        if undefined.any():
            want = np.where(undefined, port.translate(img, dx, dy, st, 1), want)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want, err_msg=f"case {case}: {dt} {h}x{w} {st!r} dx={dx!r} dy={dy!r}")


def test_fuzz_ecc_solver(port):
    """rirb_ecc_compute against the numpy restatement of OpenCV's iteration: random window sizes, shifts, warm starts, masks
    and quantile clamps.  Same iteration count, shifts within 2e-5 px, rho within 1e-7 (tests/test_ecc.py explains the bars)."""
    import ctypes as ct

    from librir_b200 import _lib
    from oracle import ecc as oe
    from tests import ecc_cases as ec

    lib = _lib.load()
    rng = np.random.default_rng(507 + SEED)
    wandering = 0
    for case in range(12 * SCALE):
        h, w = int(rng.integers(24, 140)), int(rng.integers(24, 180))
        dx, dy = float(rng.uniform(-4, 4)), float(rng.uniform(-4, 4))
        t, i = ec.small_pair(dx, dy, h=h, w=w, k=int(rng.integers(0, 1000)))
        scale, offset = float(rng.choice([1.0, 5000.0])), float(rng.choice([0.0, 8000.0]))
        t, i = (t * scale + offset).astype(np.float32), (i * np.float32(scale * rng.uniform(0.8, 1.2)) + np.float32(offset)).astype(np.float32)
        mask = None
        if rng.integers(0, 2):
            mask = np.zeros((h, w), np.uint8)
            mask[int(h * 0.1):int(h * 0.9), int(w * 0.15):w] = 1
        thresh = float("inf") if rng.integers(0, 2) else float(np.quantile(t, 0.9))
        tx0, ty0 = (0.0, 0.0) if rng.integers(0, 2) else (float(np.float32(dx + rng.uniform(-1, 1))), float(np.float32(dy + rng.uniform(-1, 1))))
        # the restatement on the clamped, normalised images (the library does both itself)
        t1, i1 = t.copy(), i.copy()
        if np.isfinite(thresh):
            m = (t1 > thresh) | (i1 > thresh)
            t1[m] = thresh
            i1[m] = thresh
        t1 = (t1 - t1.min()) / (t1.max() - t1.min())
        i1 = (i1 - i1.min()) / (i1.max() - i1.min())
        try:
            want = oe.find_transform_ecc_translation(t1, i1, tx0, ty0, mask=mask)
        except oe.ECCError as e:
            want = 1 if "NaN" in str(e) else 2
        hd = lib.rirb_ecc_open(w, h)
        try:
            assert lib.rirb_ecc_set_image(hd, 0, t.ctypes.data_as(ct.c_void_p), w) == 0
            assert lib.rirb_ecc_set_image(hd, 1, i.ctypes.data_as(ct.c_void_p), w) == 0
            if mask is not None:
                assert lib.rirb_ecc_set_mask(hd, 0, mask.ctypes.data_as(ct.c_void_p)) == 0
            shift = np.array([tx0, ty0], dtype=np.float32)
            rho, its = ct.c_double(0), ct.c_int(0)
            st = lib.rirb_ecc_compute(hd, thresh, 1 if mask is not None else 0, 500, 1e-3, shift.ctypes.data_as(ct.c_void_p), ct.byref(rho),
                                      ct.byref(its))
        finally:
            lib.rirb_ecc_close(hd)
        what = f"case {case}: {h}x{w} shift ({dx:.3f},{dy:.3f}) start ({tx0},{ty0}) mask {mask is not None} thresh {thresh}"
        if isinstance(want, int):
            assert st == want, what
            continue
        if want[3] > 50:
            # No convergence (tiny windows, shifts of several pixels): the iteration wanders between neighbouring 1/32-pixel
            # steps of OpenCV's warp until the cap or until the 1e-3 stopping test happens to fire; hundreds of chained
            # updates amplify rounding, so two correct implementations end in different places (cv2 and the restatement do
            # too).  No number can be compared there, but the regime can: the call comes back without a device error, and a
            # solve that reports success wandered as well (not a two-iteration "convergence" on garbage) and ended on
            # finite values with a correlation that is one.
            assert st in (0, 1, 2), what
            if st == 0:
                assert its.value > 25, what + f" -> converged in {its.value} iterations where the restatement needs {want[3]}"
                assert np.isfinite(shift).all() and np.isfinite(rho.value) and abs(rho.value) <= 1.0 + 1e-6, what + f" -> rho {rho.value}, shift {shift}"
            wandering += 1
            continue
        assert st == 0 and its.value == want[3], what + f" -> {st}, {its.value} iterations vs {want[3]}"
        # (a run-away solution -- the window pushed hundreds of pixels out of the image -- is compared relative to its size)
        tol = 2e-5 * max(1.0, abs(want[1]), abs(want[2]))
        assert abs(rho.value - want[0]) < 1e-7 * max(1.0, tol / 2e-5) and abs(shift[0] - want[1]) < tol and abs(shift[1] - want[2]) < tol, \
            what + f" -> rho {rho.value!r} vs {want[0]!r}, shift {shift[0]!r},{shift[1]!r} vs {want[1]!r},{want[2]!r}, {its.value} iterations"
    assert wandering <= max(1, 12 * SCALE // 10), f"{wandering} of {12 * SCALE} cases did not converge"


def test_fuzz_process_movie_host(port):
    """The one-call host path with sub-chunks shrunk (RIRB_HOST_SUB_BYTES) so that small movies rotate through all three
    stream slots: ragged tails, GOP-aligned offsets, pageable and pinned buffers -- against the staged device calls."""
    rng = np.random.default_rng(508 + SEED)
    old = os.environ.get("RIRB_HOST_SUB_BYTES")
    try:
        for case in range(10 * SCALE):
            w = int(rng.choice([16, 24, 40, 64, 96, 128, 136, 33, 7]))
            h = int(rng.integers(3, 70))
            n = int(rng.integers(1, 260))
            gop = int(rng.choice([1, 3, 7, 50]))
            delta = bool(rng.integers(0, 2))
            first = int(rng.integers(0, 5)) * gop
            sigma = float(rng.choice([0.5, 1.0, 1.7]))
            st = str(rng.choice(["nearest", "background", "wrap"]))
            mov = np.stack([random_frame(rng, h, w) for _ in range(min(n, 6))])[rng.integers(0, min(n, 6), n)]
            mov = np.ascontiguousarray(mov)
            dx = rng.uniform(-3, 3, n).astype(np.float32)
            dy = rng.uniform(-3, 3, n).astype(np.float32)
            os.environ["RIRB_HOST_SUB_BYTES"] = str(int(rng.choice([1, 4096, 50000, 300000, 1 << 26])))
            bp = sp.BadPixels(mov[0])
            smoothed = np.empty((n, h, w), np.float32) if rng.integers(0, 2) else None
            frames = mov
            if rng.integers(0, 2):
                frames = torch.from_numpy(mov.view(np.int16)).pin_memory().view(torch.uint16)
            lo, hi = movie.process_movie_host(bp, frames, dx, dy, sigma, st, 321, gop=gop, delta=delta, first_frame=first, smoothed=smoothed)
            corr = bp.correct_batch(mov)
            reg = sp.translate_batch(corr, dx, dy, st, 321)
            wlo, whi = vio.precode_movie(reg, gop=gop, delta=delta, first_frame=first)
            what = f"case {case}: {n}x{h}x{w} gop {gop} delta {delta} first {first} sub {os.environ['RIRB_HOST_SUB_BYTES']}"
            np.testing.assert_array_equal(lo, wlo, err_msg=what)
            np.testing.assert_array_equal(hi, whi, err_msg=what)
            if smoothed is not None:
                np.testing.assert_array_equal(smoothed, sp.gaussian_filter_batch(corr, sigma), err_msg=what)
            t = int(rng.integers(0, n))  # and one frame against the oracle chain
            oxy, _thr, oclamp = port.bad_pixels_detect(mov[0])
            oc_ = port.bad_pixels_correct_with(oxy, oclamp, mov[t])
            np.testing.assert_array_equal(corr[t], oc_, err_msg=what)
            np.testing.assert_array_equal(reg[t], port.translate(oc_, dx[t], dy[t], st, 321), err_msg=what)
    finally:
        if old is None:
            os.environ.pop("RIRB_HOST_SUB_BYTES", None)
        else:
            os.environ["RIRB_HOST_SUB_BYTES"] = old


def test_fuzz_split_merge_and_loader_pieces(port):
    """The writer's / reader's single-frame plane layouts with random line sizes and integration-time images, the in-place
    loader correction and the stand-alone motion step."""
    rng = np.random.default_rng(509 + SEED)
    for case in range(30 * SCALE):
        w = int(rng.choice(WIDTHS))
        h = int(rng.integers(1, 80))
        img = random_frame(rng, h, w)
        it = rng.integers(0, 256, (h, w), dtype=np.uint8) if rng.integers(0, 2) else None
        ls = int(w + rng.choice([0, 1, 7, 32, 63])) if rng.integers(0, 2) else None
        what = f"case {case}: {h}x{w} ls {ls} it {it is not None}"
        for got, want in zip(vio.split_yuv444(img, it, ls), port.split_444(img, it, ls)):
            w_ = want.copy()
            w_[:, w:] = 0  # the oracle marks untouched padding with 0xAA, the wrapper starts from zeros
            np.testing.assert_array_equal(got, w_, err_msg=what)
        y, u, v = vio.split_yuv444(img, it, ls)
        back, it_back = vio.merge_yuv444(y, u, v, w)
        np.testing.assert_array_equal(back, img, err_msg=what)
        if it is not None:
            np.testing.assert_array_equal(it_back, it, err_msg=what)
        want420 = port.split_420(img, ls)
        p = vio.split_yuv420(img, None, ls)
        np.testing.assert_array_equal(p[:, :w], want420[:, :w], err_msg=what)
        np.testing.assert_array_equal(vio.merge_yuv420(p, w), port.merge_420(want420, w), err_msg=what)
        # loader pieces on a stack with 3 metadata rows
        if h >= 6 and w >= 3:
            n = int(rng.integers(1, 4))
            mov = np.stack([random_frame(rng, h, w) for _ in range(n)])
            lbp = vio.LoaderBadPixels(mov[0])
            xy = sp.bad_pixels_list(lbp.handle)[0]
            want = mov.copy()
            for t in range(n):
                want[t, :h - 3] = port.loader_remove_bad_pixels(mov[t, :h - 3], xy)
            got = lbp.remove(mov.copy())
            np.testing.assert_array_equal(got, want, err_msg=what + " loader medians")
            np.testing.assert_array_equal(lbp.remove(to_dev(mov)).cpu().numpy().view(np.uint16), want, err_msg=what)
            sx, sy = rng.uniform(-4, 4, n), rng.uniform(-4, 4, n)
            moved = vio.remove_motion(mov, sx, sy)
            for t in range(n):
                wt = mov[t].copy()
                wt[:h - 3] = port.loader_remove_motion(mov[t, :h - 3], sx[t], sy[t])
                np.testing.assert_array_equal(moved[t], wt, err_msg=what + " motion")


def test_fuzz_batches_equal_single_frames():
    """Batched entries (one launch for a stack, host or device, per-frame shifts) against the per-frame entries."""
    rng = np.random.default_rng(510 + SEED)
    for case in range(12 * SCALE):
        w = int(rng.choice(WIDTHS))
        h = int(rng.integers(1, 100))
        n = int(rng.integers(1, 7))
        mov = np.stack([random_frame(rng, h, w) for _ in range(n)])
        st = str(rng.choice(["nearest", "background", "wrap", ""]))
        dx = rng.uniform(-5, 5, n).astype(np.float32)
        dy = rng.uniform(-5, 5, n).astype(np.float32)
        sigma = float(rng.choice([0.4, 1.0, 1.6, 2.3]))
        what = f"case {case}: {n}x{h}x{w} {st!r} sigma {sigma}"
        want_t = np.stack([sp.translate(mov[t], dx[t], dy[t], st, 77) for t in range(n)])
        np.testing.assert_array_equal(sp.translate_batch(mov, dx, dy, st, 77), want_t, err_msg=what)
        d = to_dev(mov)
        got = sp.translate_batch(d, torch.from_numpy(dx).cuda(), torch.from_numpy(dy).cuda(), st, 77)
        np.testing.assert_array_equal(got.cpu().numpy().view(np.uint16), want_t, err_msg=what + " (device)")
        want_g = np.stack([sp.gaussian_filter(mov[t], sigma) for t in range(n)])
        np.testing.assert_array_equal(sp.gaussian_filter_batch(mov, sigma), want_g, err_msg=what)
        np.testing.assert_array_equal(sp.gaussian_filter_batch(d, sigma).cpu().numpy(), want_g, err_msg=what + " (device)")
        f32 = mov.astype(np.float32)
        np.testing.assert_array_equal(sp.gaussian_filter_batch(f32, sigma), np.stack([sp.gaussian_filter(f, sigma) for f in f32]),
                                      err_msg=what + " (float32)")
