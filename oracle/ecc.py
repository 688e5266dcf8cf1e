"""oracle/ecc.py -- TEST INFRASTRUCTURE ONLY (never imported by librir_b200).

CPU restatement, in numpy, of the registration front end:

* ``MaskedRegistratorECC``          librir/registration/masked_registration_ecc.py:20-226 (start / compute: Gaussian,
                                    centred crop, quantile clamp, min/max normalisation, ECC with a warm start, the
                                    confidence rule that replaces the reference image)
* ``find_transform_ecc_translation`` the third-party algorithm the reference calls at :165-167,
                                    ``cv2.findTransformECC(template, input, warp, MOTION_TRANSLATION, criteria, mask, 1)``
                                    -- OpenCV 4.13.0 (the version in this image; librir does not pin one), the
                                    published algorithm of Evangelidis & Psarakis (PAMI 2008) as OpenCV implements it:
                                    gradients by the [-0.5 0 0.5] filter with reflect-101 borders, every iteration warps
                                    image, gradients (bilinear) and mask (nearest) with ``warpAffine`` -- whose
                                    coordinates are FIXED POINT, 10 fractional bits rounded to 1/32 pixel, weights from a
                                    32 x 32 float table, constant-zero border -- then zero-means both images under the
                                    warped mask, builds the 2 x 2 Hessian of the Jacobian (= the warped gradients, for a
                                    translation), the correlation rho, lambda, and the update H^-1 J^T (lambda T - I).

Pinned (tests/test_ecc.py): ``warp_affine_translation`` is bit-identical to ``cv2.warpAffine``; the shifts and rho of
``find_transform_ecc_translation`` agree with ``cv2.findTransformECC`` to float rounding on the committed golden
vectors (tests/golden/ecc_golden.npz, made by tests/golden/make_ecc_golden.py with cv2 itself) and, where cv2 is
importable, on live cases.
"""
from __future__ import annotations

import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS


class ECCError(RuntimeError):
    """cv2.error of findTransformECC (NaN, or correlation about to be minimised)."""


def fixed_point_origin(tx: float, ty: float, h: int, w: int):
    """warpAffine's integer source coordinates (5 fractional bits) for M = [[1,0,tx],[0,1,ty]], WARP_INVERSE_MAP."""
    tx = float(np.float32(tx))
    ty = float(np.float32(ty))
    rd = AB_SCALE // INTER_TAB // 2
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    ad = np.rint(x * AB_SCALE).astype(np.int64)
    X0 = np.rint((0.0 * y + tx) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((1.0 * y + ty) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + ad[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + np.zeros(w, dtype=np.int64)[None, :]) >> (AB_BITS - INTER_BITS)
    return X, Y


def _gather(img, yy, xx):
    h, w = img.shape
    ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
    return np.where(ok, img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], img.dtype.type(0))


def warp_affine_translation(img: np.ndarray, tx: float, ty: float, nearest: bool = False) -> np.ndarray:
    """cv2.warpAffine(img, [[1,0,tx],[0,1,ty]], size, INTER_LINEAR|NEAREST + WARP_INVERSE_MAP), border constant 0."""
    h, w = img.shape
    X, Y = fixed_point_origin(tx, ty, h, w)
    if nearest:
        # INTER_NEAREST drops the fractional bits after the same rounding offset is NOT applied: round_delta = AB_SCALE/2
        tx32, ty32 = float(np.float32(tx)), float(np.float32(ty))
        x = np.arange(w, dtype=np.float64)
        y = np.arange(h, dtype=np.float64)
        Xn = (np.rint(np.full(h, tx32) * AB_SCALE).astype(np.int64)[:, None] + AB_SCALE // 2 + np.rint(x * AB_SCALE).astype(np.int64)[None, :]) >> AB_BITS
        Yn = (np.rint((y + ty32) * AB_SCALE).astype(np.int64)[:, None] + AB_SCALE // 2 + np.zeros(w, dtype=np.int64)[None, :]) >> AB_BITS
        return _gather(img, Yn, Xn)
    ix, iy = X >> INTER_BITS, Y >> INTER_BITS
    fx = (X & (INTER_TAB - 1)).astype(np.float32) / np.float32(INTER_TAB)
    fy = (Y & (INTER_TAB - 1)).astype(np.float32) / np.float32(INTER_TAB)
    one = np.float32(1)
    w00, w01, w10, w11 = (one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx
    img = img.astype(np.float32, copy=False)
    return (_gather(img, iy, ix) * w00 + _gather(img, iy, ix + 1) * w01 + _gather(img, iy + 1, ix) * w10 +
            _gather(img, iy + 1, ix + 1) * w11).astype(np.float32)


def gradients(img: np.ndarray):
    """filter2D with [-0.5, 0, 0.5] along x and along y, BORDER_REFLECT_101 (ecc.cpp: dx = (-0.5, 0, 0.5))."""
    p = np.pad(img.astype(np.float32), 1, mode="reflect")
    half = np.float32(0.5)
    gx = (p[1:-1, 2:] * half + p[1:-1, :-2] * -half).astype(np.float32)
    gy = (p[2:, 1:-1] * half + p[:-2, 1:-1] * -half).astype(np.float32)
    return gx, gy


def find_transform_ecc_translation(template, image, tx0=0.0, ty0=0.0, iterations=500, eps=1e-3, mask=None):
    """-> (rho, tx, ty, iterations_done).  Raises ECCError where cv2 raises cv2.error."""
    T = np.asarray(template, dtype=np.float32)
    I = np.asarray(image, dtype=np.float32)
    h, w = T.shape
    pre = np.ones((h, w), np.uint8) if mask is None else (np.asarray(mask) > 0).astype(np.uint8)
    # gaussFiltSize = 1: the blurs are identities; the float mask is scaled by 0.5/0.95 and cast back to uchar (-> 1 or 0)
    pre = np.rint(pre.astype(np.float32) * np.float32(0.5 / 0.95)).astype(np.uint8)  # convertTo(CV_8U) rounds: 0.526 -> 1
    pre_f = pre.astype(np.float32)                                                  # ... and the float mask is rebuilt from it
    gx, gy = gradients(I)
    gx = gx * pre_f
    gy = gy * pre_f
    tx, ty = np.float32(tx0), np.float32(ty0)
    rho, last_rho = -1.0, -float(eps)
    it = 0
    while it < iterations and abs(rho - last_rho) >= eps:
        it += 1
        Iw = warp_affine_translation(I, tx, ty)
        gxw = warp_affine_translation(gx, tx, ty)
        gyw = warp_affine_translation(gy, tx, ty)
        m = warp_affine_translation(pre, tx, ty, nearest=True) != 0
        n = int(m.sum())
        if n == 0:
            raise ECCError("NaN encountered.")
        Iv = Iw[m].astype(np.float64)
        Tv = T[m].astype(np.float64)
        img_mean, tmp_mean = Iv.mean(), Tv.mean()
        img_std = np.sqrt(max(((Iv - img_mean) ** 2).mean(), 0.0))
        tmp_std = np.sqrt(max(((Tv - tmp_mean) ** 2).mean(), 0.0))
        Iz = Iw.copy()
        Iz[m] = (Iw[m] - np.float32(img_mean)).astype(np.float32)  # outside the mask the warped image keeps its values
        Tz = np.zeros_like(T)
        Tz[m] = (T[m] - np.float32(tmp_mean)).astype(np.float32)
        tmp_norm = np.sqrt(n * tmp_std * tmp_std)
        img_norm = np.sqrt(n * img_std * img_std)
        d = np.float64
        H = np.array([[np.sum(gxw.astype(d) ** 2), np.sum(gxw.astype(d) * gyw)],
                      [np.sum(gxw.astype(d) * gyw), np.sum(gyw.astype(d) ** 2)]]).astype(np.float32)
        # cv::invert on a 2 x 2 CV_32F matrix: closed form evaluated in double, stored as float
        det = float(H[0, 0]) * float(H[1, 1]) - float(H[0, 1]) * float(H[1, 0])
        if det == 0.0:
            raise ECCError("NaN encountered.")
        idet = 1.0 / det
        Hinv = np.array([[float(H[1, 1]) * idet, -float(H[0, 1]) * idet], [-float(H[1, 0]) * idet, float(H[0, 0]) * idet]]).astype(np.float32)
        corr = float(np.sum(Tz.astype(d) * Iz))
        last_rho = rho
        rho = corr / (img_norm * tmp_norm) if img_norm * tmp_norm != 0 else float("nan")
        if np.isnan(rho):
            raise ECCError("NaN encountered.")
        ip = np.array([np.sum(gxw.astype(d) * Iz), np.sum(gyw.astype(d) * Iz)]).astype(np.float32)
        tp = np.array([np.sum(gxw.astype(d) * Tz), np.sum(gyw.astype(d) * Tz)]).astype(np.float32)
        iph = (Hinv.astype(d) @ ip.astype(d)).astype(np.float32)  # gemm on CV_32F accumulates in double
        lam_n = img_norm * img_norm - float(np.dot(ip.astype(d), iph.astype(d)))
        lam_d = corr - float(np.dot(tp.astype(d), iph.astype(d)))
        if lam_d <= 0.0:
            raise ECCError("The algorithm stopped before its convergence. The correlation is going to be minimized.")
        lam = lam_n / lam_d
        err = (lam * Tz.astype(d) - Iz).astype(np.float32)
        ep = np.array([np.sum(gxw.astype(d) * err), np.sum(gyw.astype(d) * err)]).astype(np.float32)
        dp = (Hinv.astype(d) @ ep.astype(d)).astype(np.float32)
        tx = np.float32(tx + dp[0])
        ty = np.float32(ty + dp[1])
    return rho, float(tx), float(ty), it


class MaskedRegistratorECC:
    """masked_registration_ecc.py:20-226, restated on top of the oracle's own gaussian_filter / find_median_pixel /
    translate (``backend``: oracle.Port() or oracle.Ref()) and ``find_transform_ecc_translation`` above.
    ``ecc``: the solver, ``find_transform_ecc_translation`` by default (tests swap in cv2 itself to pin this class)."""

    def __init__(self, backend, window_factorh=0.7, window_factorv=0.7, sigma=0.5, mask=None, median=1, ref=None, pre_process=None,
                 ecc=None, shape=(512, 640)):
        self.b = backend
        self.sigma = sigma
        self.x, self.y, self.confidences = [], [], []
        self.ref_img = None
        self.ref = ref
        if ref is not None and pre_process is not None:
            self.ref = pre_process(ref)
        if sigma > 0 and self.ref is not None:
            self.ref = self.b.gaussian_filter(self.ref, sigma)
        self.subW = int(shape[1] * window_factorh)   # the reference hard-codes shape = (512, 640), :77
        self.subH = int(shape[0] * window_factorv)
        self.startX = int((shape[1] - self.subW) / 2)
        self.startY = int((shape[0] - self.subH) / 2)
        self.mask = mask
        self.qmask = None
        self.conf_thresh = None
        self.pre_process = pre_process
        self.median = median
        self.start_mat = (0.0, 0.0)  # warp_matrix[0,2], warp_matrix[1,2]
        self.ecc = ecc or find_transform_ecc_translation
        self.iterations = []

    def _crop(self, img):
        return img[self.startY:self.startY + self.subH, self.startX:self.startX + self.subW]

    def start(self, img):
        if self.pre_process is not None:
            img = self.pre_process(img)
        if self.sigma > 0:
            img = self.b.gaussian_filter(img, self.sigma)
        self.ref_img = self._crop(img)
        if self.mask is not None:
            full = np.asarray(self.mask)
            self.mask = self._crop(full)
            # the reference's wrapper reads the cropped VIEW of a contiguous uint8 mask as if it were compact
            # (rir_signal_processing.py:134-136: astype(copy=False), no ascontiguousarray): w*h consecutive bytes of the full mask
            self.qmask = np.ascontiguousarray(self.mask)
            if full.dtype == np.uint8 and full.flags["C_CONTIGUOUS"] and not self.mask.flags["C_CONTIGUOUS"]:
                s0 = self.startY * full.shape[1] + self.startX
                flat = full.reshape(-1)[s0:s0 + self.subW * self.subH]
                if flat.size == self.subW * self.subH:
                    self.qmask = np.ascontiguousarray(flat.reshape(self.subH, self.subW))
        self.x.append(0)
        self.y.append(0)
        self.confidences.append(1)

    def compute(self, img):
        if self.pre_process is not None:
            img = self.pre_process(img)
        if self.sigma > 0:
            img = self.b.gaussian_filter(img, self.sigma)
        new_im = self._crop(img).copy()
        im1 = np.array(self.ref_img if self.ref is None else self.ref, dtype=np.float32)
        im2 = np.array(new_im, dtype=np.float32)
        mask = self.mask
        if self.median < 1:
            qm = self.qmask if mask is not None else None
            t1 = self.b.find_median_pixel(new_im.astype(np.uint16), self.median, qm)
            t2 = self.b.find_median_pixel(np.asarray(self.ref_img).astype(np.uint16), self.median, qm)
            thresh = max(t1, t2)
            m = (im1 > thresh) | (im2 > thresh)
            im1[m] = thresh
            im2[m] = thresh
        im1 = (im1 - np.min(im1)) / (np.max(im1) - np.min(im1))
        im2 = (im2 - np.min(im2)) / (np.max(im2) - np.min(im2))
        cc, tx, ty, it = self.ecc(im1, im2, self.start_mat[0], self.start_mat[1], 500, 1e-3, mask)
        self.start_mat = (tx, ty)
        self.iterations.append(it)
        self.confidences.append(cc)
        self.x.append(tx)
        self.y.append(ty)
        if len(self.confidences) > 20 and self.ref is None:
            if self.conf_thresh is None:
                self.conf_thresh = np.min(self.confidences) - 2 * np.std(self.confidences)
            if cc < self.conf_thresh:
                self.ref_img = self.b.translate(new_im, -tx, -ty)
                self.start_mat = (0.0, 0.0)
        return [ty, tx]


def cv2_ecc(template, image, tx0, ty0, iterations, eps, mask):
    """The same call signature as find_transform_ecc_translation, answered by OpenCV itself."""
    import cv2

    warp = np.eye(2, 3, dtype=np.float32)
    warp[0, 2], warp[1, 2] = tx0, ty0
    criteria = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, iterations, eps)
    try:
        cc, warp = cv2.findTransformECC(template, image, warp, cv2.MOTION_TRANSLATION, criteria, mask, 1)
    except cv2.error as e:
        raise ECCError(str(e))
    return cc, float(warp[0, 2]), float(warp[1, 2]), -1
