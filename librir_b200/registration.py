"""Host-side mirror of ``librir.registration.masked_registration_ecc``.

``MaskedRegistratorECC`` keeps the reference's interface -- constructor arguments, ``start`` / ``compute``, the ``x`` / ``y`` /
``confidences`` lists, the reference-image reset rule, ``manage_computation_and_tries`` -- but every image operation runs on
the GPU through the C ABI: the Gaussian (``rirb_gaussian_filter*_batch``), the quantile thresholds, the min/max
normalisation and OpenCV's ECC iteration (``rirb_ecc_*``, csrc/ecc.cu).  A frame is uploaded once (or handed over as a torch
CUDA tensor); what comes back per frame is the shift, the correlation and the iteration count.

Where the reference raises ``cv2.error`` (ECC did not converge) this raises ``ECCError``, which derives from ``cv2.error``
when OpenCV is importable, so ``except cv2.error`` keeps working.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np

from . import _lib
from . import signal_processing as sp

try:  # pragma: no cover - depends on the environment
    import cv2 as _cv2

    _ErrBase = _cv2.error
except Exception:  # noqa: BLE001
    _ErrBase = RuntimeError


class ECCError(_ErrBase):
    """findTransformECC's failures: status 1 'NaN encountered.', status 2 'The algorithm stopped before its convergence...'."""


_MESSAGES = {
    1: "NaN encountered.",
    2: "The algorithm stopped before its convergence. The correlation is going to be minimized. "
       "Images may be uncorrelated or non-overlapped",
}


def quantile_mask_like_reference(full_mask, x0, y0, w, h):
    """The pixel set the reference's quantile thresholds are REALLY taken under.

    ``find_median_pixel(new_im, median, mask)`` receives the cropped mask as a non-contiguous view; the wrapper's
    ``mask.astype(np.uint8, copy=False)`` (rir_signal_processing.py:134-136) only compacts it when the dtype changes, so
    for a C-contiguous uint8 full-size mask the C code reads ``w * h`` consecutive bytes of the FULL mask starting at the
    window's first pixel.  Reproduced here so that thresholds (and therefore shifts) equal the reference's; any other
    dtype / layout is compacted by the reference too and gives the plain crop."""
    full = np.asarray(full_mask)
    crop = full[y0:y0 + h, x0:x0 + w]
    if full.dtype == np.uint8 and full.flags["C_CONTIGUOUS"] and not crop.flags["C_CONTIGUOUS"]:
        start = y0 * full.shape[1] + x0
        flat = full.reshape(-1)[start:start + w * h]
        if flat.size == w * h:
            return np.ascontiguousarray(flat.reshape(h, w))
    return np.ascontiguousarray(crop).astype(np.uint8)


def _torch():
    import torch

    return torch


class MaskedRegistratorECC:
    """
    Compute sub-pixels translations in a movie with the ECC algorithm (OpenCV's findTransformECC, translation model).

    Same behaviour as librir's class (masked_registration_ecc.py:20-226): the first image goes to ``start()``, the others
    to ``compute()``; results accumulate in ``x``, ``y`` (translations from the first image) and ``confidences``; the
    reference image is replaced when the confidence drops below min - 2 std of the first 21 confidences.
    """

    def __init__(self, window_factorh=0.7, window_factorv=0.7, sigma=0.5, mask=None, median=1, ref=None, pre_process=None, view=None,
                 shape=(512, 640)):
        self.sigma = sigma
        self.x = []
        self.y = []
        self.confidences = []
        self.iterations = []
        self.window_factorH = window_factorh
        self.window_factorV = window_factorv
        self.subW = int(shape[1] * self.window_factorH)  # the reference hard-codes shape = (512, 640), :77
        self.subH = int(shape[0] * self.window_factorV)
        self.startX = int((shape[1] - self.subW) / 2)
        self.startY = int((shape[0] - self.subH) / 2)
        self.conf_thresh = None
        self.pre_process = pre_process
        self.view = view
        self.median = median
        self.start_mat = np.eye(2, 3, dtype=np.float32)
        self.mask = mask
        self._lib = _lib.load()
        self._h = self._lib.rirb_ecc_open(self.subW, self.subH)
        if self._h <= 0:
            raise RuntimeError(f"MaskedRegistratorECC: {_lib.last_error()}")
        self._fixed_ref = ref is not None
        self._started = False
        if ref is not None:
            if pre_process is not None:
                ref = pre_process(ref)
            g = self._filtered(ref)
            if tuple(g.shape) != (self.subH, self.subW):
                raise RuntimeError("MaskedRegistratorECC: a fixed `ref` must have the size of the registration window")
            self._set(0, g, 0, 0)

    def __del__(self):
        try:
            if getattr(self, "_h", 0):
                self._lib.rirb_ecc_close(self._h)
                self._h = 0
        except Exception:  # noqa: BLE001
            pass

    # ---- device helpers -----------------------------------------------------------------------------
    def _filtered(self, img):
        """img (numpy or torch, any dtype) -> float32 CUDA tensor [h, w], Gaussian-filtered when sigma > 0."""
        torch = _torch()
        if not sp._is_torch(img):
            a = np.ascontiguousarray(img)
            if a.dtype == np.uint16:
                img = torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
            else:
                img = torch.from_numpy(a.astype(np.float32, copy=False)).cuda()
        elif not img.is_cuda:
            img = img.cuda()
        if img.dtype not in (torch.uint16, torch.float32):
            img = img.to(torch.float32)
        if self.sigma > 0:
            return sp.gaussian_filter_batch(img[None], self.sigma)[0]
        return img.to(torch.float32)

    def _set(self, which, g, x0, y0):
        _lib.use_torch_stream()
        ptr = g.data_ptr() + 4 * (y0 * g.shape[1] + x0)
        _lib.check(self._lib.rirb_ecc_set_image(self._h, which, ct.c_void_p(ptr), int(g.shape[1])), "ecc_set_image")

    # ---- the reference's interface ----------------------------------------------------------------------
    def start(self, img):
        if self.pre_process is not None:
            img = self.pre_process(img)
        g = self._filtered(img)
        if not self._fixed_ref:
            self._set(0, g, self.startX, self.startY)
        else:
            self._set(1, g, self.startX, self.startY)
        if self.mask is not None:
            full = np.asarray(self.mask)
            self.mask = full[self.startY:self.startY + self.subH, self.startX:self.startX + self.subW]
            ecc_mask = np.ascontiguousarray(self.mask).astype(np.uint8)
            qmask = quantile_mask_like_reference(full, self.startX, self.startY, self.subW, self.subH)
            _lib.check(self._lib.rirb_ecc_set_mask(self._h, 0, ecc_mask.ctypes.data_as(ct.c_void_p)), "ecc_set_mask")
            _lib.check(self._lib.rirb_ecc_set_mask(self._h, 1, qmask.ctypes.data_as(ct.c_void_p)), "ecc_set_mask")
        self._started = True
        self.x.append(0)
        self.y.append(0)
        self.confidences.append(1)

    def compute(self, img):
        if not self._started:
            raise RuntimeError("MaskedRegistratorECC.compute: call start() with the first image")
        if self.pre_process is not None:
            img = self.pre_process(img)
        g = self._filtered(img)
        self._set(1, g, self.startX, self.startY)
        use_mask = 1 if self.mask is not None else 0
        thresh = float("inf")
        if self.median < 1:
            if self._fixed_ref:
                raise NotImplementedError("median < 1 together with a fixed `ref` image")
            t1 = _lib.check(self._lib.rirb_ecc_quantile(self._h, 1, float(self.median), use_mask), "ecc_quantile")
            t2 = _lib.check(self._lib.rirb_ecc_quantile(self._h, 0, float(self.median), use_mask), "ecc_quantile")
            thresh = float(max(t1, t2))
        shift = np.array([self.start_mat[0, 2], self.start_mat[1, 2]], dtype=np.float32)
        rho, its = ct.c_double(0.0), ct.c_int(0)
        status = self._lib.rirb_ecc_compute(self._h, thresh, use_mask, 500, 1e-3, shift.ctypes.data_as(ct.c_void_p), ct.byref(rho),
                                            ct.byref(its))
        if status < 0:
            raise RuntimeError(f"An error occured while calling 'ecc_compute': {_lib.last_error()}")
        if status > 0:
            raise ECCError(_MESSAGES.get(status, "findTransformECC failed"))
        warp_matrix = np.eye(2, 3, dtype=np.float32)
        warp_matrix[0, 2], warp_matrix[1, 2] = shift[0], shift[1]
        self.start_mat = warp_matrix
        shift = [warp_matrix[1, 2], warp_matrix[0, 2]]
        confidence = rho.value
        self.iterations.append(its.value)
        self.confidences.append(confidence)
        self.x.append(shift[1])
        self.y.append(shift[0])
        if len(self.confidences) > 20 and not self._fixed_ref:
            if self.conf_thresh is None:
                self.conf_thresh = np.min(self.confidences) - 2 * np.std(self.confidences)
            if confidence < self.conf_thresh:
                _lib.check(self._lib.rirb_ecc_reset_reference(self._h, float(-shift[1]), float(-shift[0])), "ecc_reset_reference")
                self.start_mat = np.eye(2, 3, dtype=np.float32)
        return shift

    def append_last_coordinates_and_confidence(self):
        self.x.append(self.x[-1])
        self.y.append(self.y[-1])
        self.confidences.append(self.confidences[-1])

    def return_coordinates_and_confidence_values(self):
        return np.array([self.x, self.y, self.confidences]).T

    @property
    def stabilisation_data(self):
        import pandas as pd

        return pd.DataFrame(data=self.return_coordinates_and_confidence_values(),
                            columns=["x-axis translations", "y-axis translations", "Confidence level"])

    def to_reg_file(self, dest_file):
        self.stabilisation_data.to_csv(dest_file, sep="\t")


def manage_computation_and_tries(img, regis_obj: MaskedRegistratorECC):
    """masked_registration_ecc.py:229-260: up to five tries with the median lowered by 0.01 each time, then the previous
    estimate."""
    nb_try = 0
    max_try = 5
    compute = False
    while nb_try < max_try and not compute:
        try:
            regis_obj.compute(img)
            compute = True
            regis_obj.median = 1 if regis_obj.median < 1 else regis_obj.median
        except ECCError:
            regis_obj.median -= 0.01
            nb_try += 1
    if nb_try >= max_try:
        regis_obj.append_last_coordinates_and_confidence()
    return regis_obj
