"""Registration front end: MaskedRegistratorECC + OpenCV's findTransformECC (translation model).

  golden     tests/golden/ecc_golden.npz -- outputs of the reference's own class and of cv2 4.13.0
             (tests/golden/make_ecc_golden.py); inputs come from tests/ecc_cases.py
  port       oracle/ecc.py (numpy restatement of ecc.cpp + warpAffine's fixed point, and of the class)
  product    librir_b200.registration / rirb_ecc_* (csrc/ecc.cu)

Tolerances (floating point; the solver stops on |rho - last_rho| < 1e-3, so its output is a smooth function of
the sums): shifts 1e-4 px, rho 1e-6 between port and cv2 (cv2 accumulates its dot products in float blocks);
product against port 2e-5 px / 1e-7 with identical iteration counts.
"""
import os

import numpy as np
import pytest

from oracle import ecc as oe
from oracle import oracle as O
from tests import ecc_cases as ec

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ecc_golden.npz"))
SHIFT_TOL_CV2, RHO_TOL_CV2 = 1e-4, 1e-6
SHIFT_TOL, RHO_TOL = 2e-5, 1e-7
# After the reference image has been replaced, every estimate is chained to the previous ones through a resampled
# reference while the solver restarts from the identity, ~9 px away, and stops after two or three iterations: the
# output is then sensitive to its input (and OpenCV's warp moves in steps of 1/32 pixel), so rounding-level
# differences grow by about x3 per chained reset.  Over the ten resets of the flare sequence the numpy restatement
# ends 3e-3 px from cv2 and the GPU path 1e-2 px; the first five resets stay within 1e-3.
CHAIN_SHIFT_TOL, CHAIN_RHO_TOL = 1e-3, 1e-5
CHAIN_END_TOL = 5e-2


# ---- the port against OpenCV's recorded outputs ---------------------------------------------------------
def test_port_warp_affine_is_bit_exact():
    t, _ = ec.small_pair(1.2, 0.7, k=9)
    m8 = (t > 0.3).astype(np.uint8)
    for k, (tx, ty) in enumerate(GOLD["warp_shifts"]):
        assert np.array_equal(oe.warp_affine_translation(t, tx, ty), GOLD[f"warp_lin_{k}"])
        assert np.array_equal(oe.warp_affine_translation(m8, tx, ty, nearest=True), GOLD[f"warp_nn_{k}"])


def test_port_solver_matches_cv2_golden():
    for k, (dx, dy) in enumerate(ec.SMALL_CASES):
        t, i = ec.small_pair(dx, dy, k=k)
        rho, tx, ty, it = oe.find_transform_ecc_translation(t, i)
        g = GOLD["small_results"][k]
        assert abs(rho - g[0]) < RHO_TOL_CV2 and abs(tx - g[1]) < SHIFT_TOL_CV2 and abs(ty - g[2]) < SHIFT_TOL_CV2, (k, rho, tx, ty, g)
        assert abs(tx - dx) < 0.1 and abs(ty - dy) < 0.1  # and the answer is the shift that was put in
    t, i = ec.small_pair(1.2, 0.7, k=9)
    mask = np.zeros(t.shape, np.uint8)
    mask[10:100, 20:150] = 1
    rho, tx, ty, _ = oe.find_transform_ecc_translation(t, i, mask=mask)
    assert np.allclose([rho, tx, ty], GOLD["small_masked"], atol=SHIFT_TOL_CV2)
    rho, tx, ty, _ = oe.find_transform_ecc_translation(t, i, 1.0, 0.5)
    assert np.allclose([rho, tx, ty], GOLD["small_warm"], atol=SHIFT_TOL_CV2)


def test_port_solver_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    crit = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 500, 1e-3)
    for k, (dx, dy) in enumerate([(0.77, 2.2), (-3.1, -1.4), (5.5, 0.25)]):
        t, i = ec.small_pair(dx, dy, h=90, w=130, k=20 + k)
        cc, M = cv2.findTransformECC(t, i, np.eye(2, 3, dtype=np.float32), cv2.MOTION_TRANSLATION, crit, None, 1)
        rho, tx, ty, _ = oe.find_transform_ecc_translation(t, i)
        assert abs(rho - cc) < RHO_TOL_CV2 and abs(tx - M[0, 2]) < SHIFT_TOL_CV2 and abs(ty - M[1, 2]) < SHIFT_TOL_CV2


def _run(reg, mov, n=None):
    reg.start(mov[0])
    for t in range(1, n or len(mov)):
        reg.compute(mov[t])
    return np.array([reg.x, reg.y, reg.confidences], dtype=np.float64)


def _close(a, g, shift_tol, rho_tol):
    assert a.shape == g.shape
    assert np.max(np.abs(a[:2] - g[:2])) < shift_tol, np.max(np.abs(a[:2] - g[:2]))
    assert np.max(np.abs(a[2] - g[2])) < rho_tol, np.max(np.abs(a[2] - g[2]))


def test_port_class_matches_reference_class_golden():
    """The restated class (oracle Gaussian / quantile / translate + restated ECC) against librir's own class + cv2."""
    mov, sx, sy = ec.movie(30)
    a = _run(oe.MaskedRegistratorECC(O.Port()), mov)
    _close(a, GOLD["seq_default"], SHIFT_TOL_CV2, RHO_TOL_CV2)
    assert np.max(np.abs(a[0] - (sx - sx[0]))) < 0.15 and np.max(np.abs(a[1] - (sy - sy[0]))) < 0.15  # it tracks the camera
    static = np.ones(mov[0].shape, np.uint8)
    static[:, :140] = 0
    a = _run(oe.MaskedRegistratorECC(O.Port(), mask=static, median=0.9), mov, 16)
    _close(a, GOLD["seq_masked"], SHIFT_TOL_CV2, RHO_TOL_CV2)


def test_port_class_reset_rule_and_failures():
    mov2, _, _ = ec.movie(34, flare_from=24, flare=ec.FLARE)
    reg = oe.MaskedRegistratorECC(O.Port())
    a = _run(reg, mov2)
    g = GOLD["seq_reset"]
    _close(a[:, :25], g[:, :25], SHIFT_TOL_CV2, RHO_TOL_CV2)  # up to and including the frame that triggers the first reset
    _close(a[:, :30], g[:, :30], CHAIN_SHIFT_TOL, CHAIN_RHO_TOL)
    _close(a, g, CHAIN_END_TOL, CHAIN_RHO_TOL)
    assert abs(reg.conf_thresh - float(GOLD["seq_reset_thresh"])) < 1e-6
    assert np.array_equal(a[2] < reg.conf_thresh, g[2] < float(GOLD["seq_reset_thresh"]))  # the same frames replace the reference
    assert list(GOLD["fail_raises"]) == [1, 1]
    for img in ec.failing_frames():
        reg = oe.MaskedRegistratorECC(O.Port())
        reg.start(mov2[0])
        with pytest.raises(oe.ECCError):
            reg.compute(img)


# ---- the product ----------------------------------------------------------------------------------------
def _solve(lib, t, i, tx0=0.0, ty0=0.0, mask=None, iters=500):
    import ctypes as ct

    h, w = t.shape
    hd = lib.rirb_ecc_open(w, h)
    assert hd > 0
    try:
        assert lib.rirb_ecc_set_image(hd, 0, t.ctypes.data_as(ct.c_void_p), w) == 0
        assert lib.rirb_ecc_set_image(hd, 1, i.ctypes.data_as(ct.c_void_p), w) == 0
        if mask is not None:
            assert lib.rirb_ecc_set_mask(hd, 0, mask.ctypes.data_as(ct.c_void_p)) == 0
        shift = np.array([tx0, ty0], dtype=np.float32)
        rho, its = ct.c_double(0), ct.c_int(0)
        st = lib.rirb_ecc_compute(hd, float("inf"), 1 if mask is not None else 0, iters, 1e-3, shift.ctypes.data_as(ct.c_void_p),
                                  ct.byref(rho), ct.byref(its))
        return st, rho.value, float(shift[0]), float(shift[1]), its.value
    finally:
        lib.rirb_ecc_close(hd)


@pytest.fixture(params=["one cooperative launch", "launch per iteration"])
def ecc_driver(request):
    from librir_b200 import _lib

    _lib.set_parameter("ecc_fused", request.param == "one cooperative launch")
    yield request.param
    _lib.set_parameter("ecc_fused", 1)


@pytest.mark.gpu
def test_product_solver_matches_port_and_cv2(ecc_driver):
    from librir_b200 import _lib

    lib = _lib.load()
    for k, (dx, dy) in enumerate(ec.SMALL_CASES):
        t, i = ec.small_pair(dx, dy, k=k)  # already normalised: the library's own normalisation is then the identity
        st, rho, tx, ty, it = _solve(lib, t, i)
        prho, ptx, pty, pit = oe.find_transform_ecc_translation(t, i)
        assert st == 0 and it == pit
        assert abs(rho - prho) < RHO_TOL and abs(tx - ptx) < SHIFT_TOL and abs(ty - pty) < SHIFT_TOL, (k, rho, tx, ty, prho, ptx, pty)
        g = GOLD["small_results"][k]
        assert abs(rho - g[0]) < RHO_TOL_CV2 and abs(tx - g[1]) < SHIFT_TOL_CV2 and abs(ty - g[2]) < SHIFT_TOL_CV2
    t, i = ec.small_pair(1.2, 0.7, k=9)
    mask = np.zeros(t.shape, np.uint8)
    mask[10:100, 20:150] = 1
    st, rho, tx, ty, _ = _solve(lib, t, i, mask=mask)
    assert st == 0 and np.allclose([rho, tx, ty], GOLD["small_masked"], atol=SHIFT_TOL_CV2)
    st, rho, tx, ty, _ = _solve(lib, t, i, 1.0, 0.5)
    assert st == 0 and np.allclose([rho, tx, ty], GOLD["small_warm"], atol=SHIFT_TOL_CV2)
    # iteration cap: one iteration only, like criteria (COUNT|EPS, 1, eps)
    st, rho, tx, ty, it = _solve(lib, t, i, iters=1)
    prho, ptx, pty, pit = oe.find_transform_ecc_translation(t, i, iterations=1)
    assert st == 0 and it == pit == 1 and abs(tx - ptx) < SHIFT_TOL and abs(ty - pty) < SHIFT_TOL
    # odd window sizes, large shifts (most of the warped image outside)
    for (h, w, dx, dy) in [(33, 47, 0.4, -0.3), (64, 257, 2.6, 1.1), (120, 160, 30.0, 22.0)]:
        t, i = ec.small_pair(dx, dy, h=h, w=w, k=40)
        st, rho, tx, ty, it = _solve(lib, t, i)
        try:
            prho, ptx, pty, pit = oe.find_transform_ecc_translation(t, i)
        except oe.ECCError as e:
            assert st == (1 if "NaN" in str(e) else 2)
            continue
        assert st == 0 and it == pit and abs(rho - prho) < RHO_TOL and abs(tx - ptx) < SHIFT_TOL and abs(ty - pty) < SHIFT_TOL


@pytest.mark.gpu
def test_product_class_matches_port_and_reference_golden(ecc_driver):
    from librir_b200 import registration as rg

    mov, sx, sy = ec.movie(30)
    reg = rg.MaskedRegistratorECC()
    a = _run(reg, mov)
    port = oe.MaskedRegistratorECC(O.Port())
    b = _run(port, mov)
    _close(a, b, 5e-5, 5e-7)  # the Gaussian in front differs from the oracle's by float rounding (1e-5 relative)
    assert reg.iterations == port.iterations
    _close(a, GOLD["seq_default"], SHIFT_TOL_CV2, RHO_TOL_CV2)
    assert np.array(reg.stabilisation_data).shape == (30, 3)
    # the whole movie in one call (frames in HBM) gives the same numbers as frame-by-frame calls
    import torch

    whole = rg.MaskedRegistratorECC()
    assert whole.compute_movie(torch.from_numpy(mov.view(np.int16)).cuda().view(torch.uint16), max_try=0) == 30
    assert np.array_equal(np.array([whole.x, whole.y, whole.confidences], dtype=np.float64), a) and whole.iterations == reg.iterations
    half = rg.MaskedRegistratorECC()  # and so does any split into several calls, numpy or torch
    half.compute_movie(mov[:7])
    half.compute_movie(mov[7:])
    assert np.array_equal(np.array([half.x, half.y, half.confidences], dtype=np.float64), a)
    static = np.ones(mov[0].shape, np.uint8)
    static[:, :140] = 0
    a = _run(rg.MaskedRegistratorECC(mask=static, median=0.9), mov, 16)
    _close(a, GOLD["seq_masked"], SHIFT_TOL_CV2, RHO_TOL_CV2)


@pytest.mark.gpu
@pytest.mark.parametrize("fh,fv,sigma,crop", [(0.7, 0.7, 0.5, None), (0.97, 0.99, 1.0, None), (1.0, 1.0, 0.5, None), (0.55, 0.8, 2.0, None),
                                              (0.7, 0.7, 0.5, (510, 636))])
def test_windows_and_the_region_the_gaussian_filters(fh, fv, sigma, crop):
    """The Gaussian in front of the solver filters only the region the window needs (window + kernel radius, clipped to the
    frame): windows that touch the frame's borders, wide kernels, and a frame width that is not a multiple of the vector
    width (the layout falls back to whole frames) must all track like the restated class."""
    from librir_b200 import registration as rg
    import torch

    mov, _, _ = ec.movie(12)
    if crop:
        mov = np.ascontiguousarray(mov[:, :crop[0], :crop[1]])
    shape = mov.shape[1:]
    a = _run(oe.MaskedRegistratorECC(O.Port(), window_factorh=fh, window_factorv=fv, sigma=sigma, shape=shape), mov)
    reg = rg.MaskedRegistratorECC(window_factorh=fh, window_factorv=fv, sigma=sigma, shape=shape)
    b = _run(reg, mov)
    _close(b, a, 5e-5, 5e-7)
    whole = rg.MaskedRegistratorECC(window_factorh=fh, window_factorv=fv, sigma=sigma, shape=shape)
    whole.compute_movie(torch.from_numpy(mov.view(np.int16)).cuda().view(torch.uint16), max_try=0)
    assert np.array_equal(np.array([whole.x, whole.y, whole.confidences], dtype=np.float64), b)


@pytest.mark.gpu
def test_product_class_reset_rule_failures_and_device_input():
    import torch

    from librir_b200 import registration as rg

    mov2, _, _ = ec.movie(34, flare_from=24, flare=ec.FLARE)
    reg = rg.MaskedRegistratorECC()
    d = torch.from_numpy(mov2.view(np.int16)).cuda().view(torch.uint16)  # frames already in HBM
    reg.start(d[0])
    for t in range(1, len(mov2)):
        reg.compute(d[t])
    a = np.array([reg.x, reg.y, reg.confidences], dtype=np.float64)
    g = GOLD["seq_reset"]
    _close(a[:, :25], g[:, :25], SHIFT_TOL_CV2, RHO_TOL_CV2)
    _close(a[:, :30], g[:, :30], CHAIN_SHIFT_TOL, CHAIN_RHO_TOL)
    _close(a, g, CHAIN_END_TOL, CHAIN_RHO_TOL)
    assert abs(reg.conf_thresh - float(GOLD["seq_reset_thresh"])) < 1e-6
    assert np.array_equal(a[2] < reg.conf_thresh, g[2] < float(GOLD["seq_reset_thresh"]))
    # the whole movie in one call: solves queued back to back on the device ("ecc_queue", the default) give the same numbers
    # as frame-by-frame calls and as the one-launch-per-frame loop, across the replacements of the reference image
    from librir_b200 import _lib

    runs = {}
    for q in (1, 0):
        _lib.set_parameter("ecc_queue", q)
        whole = rg.MaskedRegistratorECC()
        assert whole.compute_movie(d, max_try=0) == len(mov2)
        runs[q] = np.array([whole.x, whole.y, whole.confidences], dtype=np.float64)
        assert whole.iterations == reg.iterations and whole.conf_thresh == reg.conf_thresh
    _lib.set_parameter("ecc_queue", 1)
    assert np.array_equal(runs[1], a) and np.array_equal(runs[0], a)
    # ... and a failing frame inside a queued run ends it, goes through the retries, and the run resumes behind it
    bad = ec.failing_frames()[0]
    mixed = np.concatenate([mov2[:9], bad[None], mov2[9:14]])
    one = rg.MaskedRegistratorECC()
    one.compute_movie(torch.from_numpy(mixed.view(np.int16)).cuda().view(torch.uint16), max_try=5)
    two = rg.MaskedRegistratorECC()
    two.start(mixed[0])
    for t in range(1, len(mixed)):
        rg.manage_computation_and_tries(mixed[t], two)
    assert np.array_equal(np.array([one.x, one.y, one.confidences]), np.array([two.x, two.y, two.confidences])) and one.median == two.median
    for k, img in enumerate(ec.failing_frames()):
        reg = rg.MaskedRegistratorECC()
        reg.start(mov2[0])
        with pytest.raises(rg.ECCError):
            reg.compute(img)
        n = len(reg.x)
        # retries with the median lowered by 0.01 (values below: the same loop run with OpenCV on the host)
        rg.manage_computation_and_tries(img, reg)
        assert len(reg.x) == n + 1
        if k == 0:  # contrast-inverted: five failures, then the previous estimate is repeated
            assert reg.x[-1] == reg.x[-2] and reg.y[-1] == reg.y[-2] and reg.median == pytest.approx(0.95)
        else:       # flat: the clamp writes the reference's bright structures into both images, the second try converges
            assert abs(reg.x[-1] - -0.0034288063) < 1e-5 and abs(reg.y[-1] - 0.0018962524) < 1e-5 and reg.median == 1
    try:
        import cv2

        assert issubclass(rg.ECCError, cv2.error)
    except ImportError:
        pass
