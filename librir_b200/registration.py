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
    reference image is replaced when the confidence drops below min - 2 std of the first 21 confidences.  The loop itself
    (warm start, confidence rule, retry rule) lives in the library (``rirb_ecc_track``); ``compute_movie`` hands it a whole
    stack of frames in one call.
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
        self.shape = tuple(shape)
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
        _lib.check(self._lib.rirb_ecc_track_config(self._h, float(sigma), float(median), int(self._fixed_ref)), "ecc_track_config")
        if ref is not None:
            if pre_process is not None:
                ref = pre_process(ref)
            g = self._filtered(ref)
            if tuple(g.shape) != (self.subH, self.subW):
                raise RuntimeError("MaskedRegistratorECC: a fixed `ref` must have the size of the registration window")
            _lib.use_torch_stream()
            _lib.check(self._lib.rirb_ecc_set_image(self._h, 0, ct.c_void_p(g.data_ptr()), int(g.shape[1])), "ecc_set_image")

    def __del__(self):
        try:
            if getattr(self, "_h", 0):
                self._lib.rirb_ecc_close(self._h)
                self._h = 0
        except Exception:  # noqa: BLE001
            pass

    # ---- helpers ----------------------------------------------------------------------------------------
    def _filtered(self, img):
        """img (numpy or torch, any dtype) -> float32 CUDA tensor [h, w], Gaussian-filtered when sigma > 0 (fixed `ref` only)."""
        torch = _torch()
        if not sp._is_torch(img):
            a = np.ascontiguousarray(img)
            img = torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16) if a.dtype == np.uint16 else torch.from_numpy(
                a.astype(np.float32, copy=False)).cuda()
        elif not img.is_cuda:
            img = img.cuda()
        if img.dtype not in (torch.uint16, torch.float32):
            img = img.to(torch.float32)
        if self.sigma > 0:
            return sp.gaussian_filter_batch(img[None], self.sigma)[0]
        return img.to(torch.float32)

    def _frames(self, frames):
        """-> (ctypes pointer, dtype code, n, keep-alive): a stack [n, h, w] (or one frame [h, w]) of uint16 / float32."""
        if sp._is_torch(frames):
            torch = _torch()
            t = frames if frames.dtype in (torch.uint16, torch.float32) else frames.to(torch.float32)
            t = t.contiguous()
            if t.is_cuda:
                _lib.use_torch_stream()
            shape, code, ptr = tuple(t.shape), ("H" if t.dtype == torch.uint16 else "f"), ct.c_void_p(t.data_ptr())
        else:
            t = np.asarray(frames)
            t = np.ascontiguousarray(t if t.dtype in (np.uint16, np.float32) else t.astype(np.float32))
            shape, code, ptr = t.shape, ("H" if t.dtype == np.uint16 else "f"), t.ctypes.data_as(ct.c_void_p)
        if len(shape) == 2:
            shape = (1,) + tuple(shape)
        if len(shape) != 3 or tuple(shape[1:]) != self.shape:
            raise RuntimeError(f"MaskedRegistratorECC: frames must be {self.shape[0]} x {self.shape[1]} images")
        return ptr, ord(code), int(shape[0]), t

    def _track(self, frames, max_try):
        if self.pre_process is not None:
            if sp._is_torch(frames) or np.asarray(frames).ndim == 3:
                frames = np.stack([np.asarray(self.pre_process(f)) for f in frames])
            else:
                frames = self.pre_process(frames)
        if not self._started and self.mask is not None:
            full = np.asarray(self.mask)
            self.mask = full[self.startY:self.startY + self.subH, self.startX:self.startX + self.subW]
            ecc_mask = np.ascontiguousarray(self.mask).astype(np.uint8)
            qmask = quantile_mask_like_reference(full, self.startX, self.startY, self.subW, self.subH)
            _lib.check(self._lib.rirb_ecc_set_mask(self._h, 0, ecc_mask.ctypes.data_as(ct.c_void_p)), "ecc_set_mask")
            _lib.check(self._lib.rirb_ecc_set_mask(self._h, 1, qmask.ctypes.data_as(ct.c_void_p)), "ecc_set_mask")
        ptr, code, n, keep = self._frames(frames)
        x, y, c = np.zeros(n), np.zeros(n), np.zeros(n)
        its = np.zeros(n, dtype=np.int32)
        done = ct.c_longlong(0)
        if self.median != self._lib_median():  # the attribute is public in the reference: honour a value the user changed
            self._set_median(self.median)
        status = self._lib.rirb_ecc_track(self._h, code, ptr, n, self.shape[1], self.shape[0], self.startX, self.startY,
                                          1 if self.mask is not None else 0, int(max_try), x.ctypes.data_as(ct.c_void_p),
                                          y.ctypes.data_as(ct.c_void_p), c.ctypes.data_as(ct.c_void_p), its.ctypes.data_as(ct.c_void_p),
                                          ct.byref(done))
        del keep
        k = done.value
        first = 0
        if not self._started and k > 0:  # start(): integer zeros and a confidence of 1, like the reference's lists
            self.x.append(0)
            self.y.append(0)
            self.confidences.append(1)
            self._started = True
            first = 1
        for i in range(first, k):
            self.x.append(np.float32(x[i]))
            self.y.append(np.float32(y[i]))
            self.confidences.append(float(c[i]))
            self.iterations.append(int(its[i]))
        self._sync_state()
        if status < 0:
            raise RuntimeError(f"An error occured while calling 'ecc_track': {_lib.last_error()}")
        if status > 0:
            raise ECCError(_MESSAGES.get(status, "findTransformECC failed"))
        return k

    def _lib_median(self):
        m = ct.c_double(0.0)
        self._lib.rirb_ecc_track_state(self._h, ct.byref(m), None, None, None)
        return m.value

    def _set_median(self, median):
        _lib.check(self._lib.rirb_ecc_track_set_median(self._h, float(median)), "ecc_track_set_median")

    def _sync_state(self):
        m, th, n = ct.c_double(0.0), ct.c_double(0.0), ct.c_longlong(0)
        st = (ct.c_float * 2)()
        self._lib.rirb_ecc_track_state(self._h, ct.byref(m), ct.byref(th), st, ct.byref(n))
        self.median = 1 if m.value == 1.0 else m.value
        self.conf_thresh = None if np.isnan(th.value) else th.value
        self.start_mat = np.eye(2, 3, dtype=np.float32)
        self.start_mat[0, 2], self.start_mat[1, 2] = st[0], st[1]

    # ---- the reference's interface ----------------------------------------------------------------------
    def start(self, img):
        if self._started:
            raise RuntimeError("MaskedRegistratorECC.start: already started")
        self._track(img, 0)

    def compute(self, img):
        if not self._started:
            raise RuntimeError("MaskedRegistratorECC.compute: call start() with the first image")
        self._track(img, 0)
        return [self.y[-1], self.x[-1]]

    def compute_movie(self, frames, max_try=5):
        """All of ``frames`` ([n, h, w], numpy or torch CUDA) in one call: start() on the first frame if the object is new,
        then manage_computation_and_tries for every other frame (``max_try=0``: plain compute(), raising on a failure)."""
        return self._track(frames, max_try)

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
    estimate; a median < 1 goes back to 1 after a success."""
    regis_obj._track(img, 5)
    return regis_obj
