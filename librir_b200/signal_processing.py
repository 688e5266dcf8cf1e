"""Host-side mirror of ``librir.signal_processing`` for the hot path.

Same names, arguments, defaults and error behaviour as the reference's ctypes wrappers
(src/python/librir/signal_processing/rir_signal_processing.py: translate :23-82,
gaussian_filter :85-113, find_median_pixel :116-147, bad_pixels_* :273-316) and its
``BadPixels`` class (BadPixels.py:16-29) -- but every call lands in the CUDA library.

The ``*_batch`` functions are additive: they take a stack of frames ``[n, h, w]`` as a numpy
array (staged through the GPU) or as a torch CUDA tensor (zero copies, runs on torch's current
stream) and process it in one launch.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np

from . import _lib

_DTYPES = {
    np.dtype(np.bool_): "?",
    np.dtype(np.int8): "b",
    np.dtype(np.uint8): "B",
    np.dtype(np.int16): "h",
    np.dtype(np.uint16): "H",
    np.dtype(np.int32): "i",
    np.dtype(np.uint32): "I",
    np.dtype(np.int64): "l",
    np.dtype(np.uint64): "L",
    np.dtype(np.float32): "f",
    np.dtype(np.float64): "d",
}


def toCharP(s):
    """librir/low_level/misc.py toCharP: str -> bytes."""
    return s.encode("ascii") if isinstance(s, str) else bytes(s)


# ----------------------------------------------------------------------------------------
# buffer helpers: numpy (host) or torch CUDA tensor (device)
# ----------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _torch_np_dtype(t):
    import torch

    return {
        torch.bool: np.dtype(np.bool_), torch.int8: np.dtype(np.int8), torch.uint8: np.dtype(np.uint8),
        torch.int16: np.dtype(np.int16), torch.uint16: np.dtype(np.uint16), torch.int32: np.dtype(np.int32),
        torch.uint32: np.dtype(np.uint32), torch.int64: np.dtype(np.int64), torch.uint64: np.dtype(np.uint64),
        torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64),
    }[t.dtype]


def _ptr(x):
    if _is_torch(x):
        if not x.is_contiguous():
            raise RuntimeError("librir_b200: tensors must be contiguous")
        return ct.c_void_p(x.data_ptr())
    return x.ctypes.data_as(ct.c_void_p)


def _dtype_of(x):
    return _torch_np_dtype(x) if _is_torch(x) else x.dtype


def _empty_like(x, dtype=None):
    if _is_torch(x):
        import torch

        if dtype is None:
            return torch.empty_like(x)
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.uint16): torch.uint16, np.dtype(np.uint8): torch.uint8}[
            np.dtype(dtype)]
        return torch.empty(x.shape, dtype=tdt, device=x.device)
    return np.empty(x.shape, dtype=dtype or x.dtype)


def _prepare_device_call(x):
    if _is_torch(x):
        if not x.is_cuda:
            raise RuntimeError("librir_b200: torch tensors must live on a CUDA device (use numpy for host data)")
        _lib.use_torch_stream()


# ----------------------------------------------------------------------------------------
# translate
# ----------------------------------------------------------------------------------------
def translate(image, dx, dy, strategy=str(), background=None):
    """Sub-pixel shift of one 2-D image by ``(dx, dy)`` with bilinear weights (rir_signal_processing.py:23-82).

    ``strategy`` decides what the pixels without a source get: ``""`` / ``"noborder"`` keep the input's value,
    ``"constant"`` (alias of ``"background"``) writes ``background``, ``"nearest"`` repeats the closest valid
    row / column, ``"wrap"`` continues from the opposite edge.  Result: same shape and dtype as ``image``.
    """
    lib = _lib.load()
    if len(image.shape) != 2:
        raise RuntimeError("translate: wrong input image dimension")
    if strategy == "background" and background is None:
        raise RuntimeError("translate: wrong background value")

    strategy = toCharP(strategy)
    if strategy == b"constant":
        strategy = b"background"
    img = np.ascontiguousarray(image)
    # "noborder" pixels keep the source value; the other strategies write every pixel
    res = np.copy(img, "C") if strategy in (b"", b"noborder") else np.empty_like(img)
    _back = np.zeros((1), dtype=img.dtype)
    if background is not None:
        _back[0] = background
    _tr = np.zeros((2), dtype=np.float32)
    _tr[0] = dx
    _tr[1] = dy
    _dtype = _DTYPES.get(image.dtype, None)
    if _dtype is None:
        raise RuntimeError("An error occured while calling 'translate'")
    r = lib.translate(ord(_dtype), _ptr(img), _ptr(res), img.shape[1], img.shape[0], _tr[0], _tr[1], _ptr(_back), strategy)
    _lib.check(r, "translate")
    return res


def translate_batch(frames, dx, dy, strategy=str(), background=None, out=None):
    """``translate`` on a stack ``[n, h, w]``; ``dx``/``dy`` scalars or per-frame sequences."""
    lib = _lib.load()
    if len(frames.shape) != 3:
        raise RuntimeError("translate_batch: wrong input dimension")
    if strategy == "background" and background is None:
        raise RuntimeError("translate: wrong background value")
    strategy = toCharP(strategy)
    if strategy == b"constant":
        strategy = b"background"
    _prepare_device_call(frames)
    n, h, w = frames.shape
    dt = _dtype_of(frames)
    code = _DTYPES.get(dt)
    if code is None:
        raise RuntimeError("An error occured while calling 'translate'")
    if out is None:
        out = frames.clone() if _is_torch(frames) else np.copy(frames, "C")
    if _is_torch(dx):
        sx, sy, ns = dx, dy, dx.numel()
    else:
        sx = np.ascontiguousarray(np.atleast_1d(dx), dtype=np.float32)
        sy = np.ascontiguousarray(np.atleast_1d(dy), dtype=np.float32)
        ns = sx.size
    _back = np.zeros((1), dtype=dt)
    if background is not None:
        _back[0] = background
    r = lib.rirb_translate_batch(ord(code), _ptr(frames), _ptr(out), w, h, n, _ptr(sx), _ptr(sy), ns, _ptr(_back), strategy)
    _lib.check(r, "translate")
    return out


# ----------------------------------------------------------------------------------------
# gaussian
# ----------------------------------------------------------------------------------------
def gaussian_filter(image, sigma=1.0):
    """Gaussian smoothing of one 2-D image, radius ``max(1, int(2 * sigma))``, taps outside the image dropped and the
    rest renormalised (rir_signal_processing.py:85-113).  The output is float32 whatever the input dtype."""
    lib = _lib.load()
    if len(image.shape) != 2:
        raise RuntimeError("gaussian_filter: wrong input image dimension")
    res = np.empty(image.shape, dtype=np.float32)
    if image.dtype == np.uint16:
        # the uint16 -> float32 conversion of the reference wrapper (rir_signal_processing.py:100) happens
        # in the kernel: half the upload, no host pass (the conversion is exact either way)
        img = np.ascontiguousarray(image)
        r = lib.rirb_gaussian_filter_u16_batch(_ptr(img), _ptr(res), img.shape[1], img.shape[0], 1, sigma)
    else:
        img = np.ascontiguousarray(image, dtype=np.float32)
        r = lib.gaussian_filter(_ptr(img), _ptr(res), img.shape[1], img.shape[0], sigma)
    _lib.check(r, "gaussian_filter")
    return res


def gaussian_filter_batch(frames, sigma=1.0, out=None):
    """Gaussian filter of a stack ``[n, h, w]`` (float32, or uint16 converted on the fly)."""
    lib = _lib.load()
    if len(frames.shape) != 3:
        raise RuntimeError("gaussian_filter_batch: wrong input dimension")
    _prepare_device_call(frames)
    n, h, w = frames.shape
    dt = _dtype_of(frames)
    if out is None:
        out = _empty_like(frames, np.float32)
    if dt == np.uint16:
        r = lib.rirb_gaussian_filter_u16_batch(_ptr(frames), _ptr(out), w, h, n, sigma)
    elif dt == np.float32:
        r = lib.rirb_gaussian_filter_batch(_ptr(frames), _ptr(out), w, h, n, sigma)
    else:
        raise RuntimeError("gaussian_filter_batch: frames must be uint16 or float32")
    _lib.check(r, "gaussian_filter")
    return out


# ----------------------------------------------------------------------------------------
# statistics
# ----------------------------------------------------------------------------------------
def find_median_pixel(image, percent=0.5, mask=None):
    """Quantile of a uint16 image through its histogram: the smallest level whose cumulated count reaches
    ``percent * size`` (optionally over the pixels where ``mask`` is non-zero); rir_signal_processing.py:116-147."""
    lib = _lib.load()
    if len(image.shape) != 2:
        raise RuntimeError("find_median_pixel: wrong input image dimension")
    image = np.ascontiguousarray(image.astype(dtype=np.uint16, copy=False))
    if mask is not None:
        mask = np.ascontiguousarray(mask.astype(dtype=np.uint8, copy=False))
        res = lib.find_median_pixel_mask(_ptr(image), _ptr(mask), image.size, float(percent))
    else:
        res = lib.find_median_pixel(_ptr(image), image.size, float(percent))
    return _lib.check(res, "find_median_pixel")


# ----------------------------------------------------------------------------------------
# bad pixels
# ----------------------------------------------------------------------------------------
def bad_pixels_create(first_image):
    """Run the bad-pixel detection on ``first_image`` and keep the flagged set (and the clamp level) behind a new
    handle, which is what the call returns (rir_signal_processing.py:273-289)."""
    lib = _lib.load()
    if _is_torch(first_image):
        _prepare_device_call(first_image)
        img = first_image
    else:
        img = np.array(first_image, dtype=np.uint16, order="C")
    ret = lib.bad_pixels_create(_ptr(img), img.shape[1], img.shape[0])
    if ret <= 0:
        raise RuntimeError(f"'bad_pixels_create': {_lib.last_error()}")
    return ret


def bad_pixels_correct_gaussian_batch(handle, frames, sigma=1.0, out=None, smoothed=None):
    """``bad_pixels_correct_batch`` followed by ``gaussian_filter_batch`` of the corrected frames, in one pass over the
    movie: returns ``(corrected uint16 [n,h,w], smoothed float32 [n,h,w])``."""
    lib = _lib.load()
    if len(frames.shape) != 3:
        raise RuntimeError("bad_pixels_correct_gaussian_batch: wrong input dimension")
    _prepare_device_call(frames)
    if out is None:
        out = _empty_like(frames)
    if smoothed is None:
        smoothed = _empty_like(frames, np.float32)
    res = lib.rirb_bad_pixels_correct_gaussian_batch(handle, _ptr(frames), _ptr(out), _ptr(smoothed), frames.shape[0], float(sigma))
    _lib.check(res, "bad_pixels_correct_gaussian")
    return out, smoothed


def bad_pixels_destroy(handle):
    """Release a handle obtained from ``bad_pixels_create``."""
    _lib.load().bad_pixels_destroy(handle)


def bad_pixels_correct(handle, img):
    """Median-replace the handle's flagged pixels in ``img`` and clamp the rest; returns a new uint16 image."""
    lib = _lib.load()
    img = np.ascontiguousarray(img, dtype=np.uint16)
    out = np.empty(img.shape, dtype=np.uint16)
    res = lib.bad_pixels_correct(handle, _ptr(img), _ptr(out))
    if res < 0:
        raise RuntimeError("'bad_pixels_correct': unknown error")
    return out


def bad_pixels_correct_batch(handle, frames, out=None):
    """Correct a stack ``[n, h, w]`` of uint16 frames in one launch."""
    lib = _lib.load()
    if len(frames.shape) != 3:
        raise RuntimeError("bad_pixels_correct_batch: wrong input dimension")
    _prepare_device_call(frames)
    if out is None:
        out = _empty_like(frames)
    res = lib.rirb_bad_pixels_correct_batch(handle, _ptr(frames), _ptr(out), frames.shape[0])
    _lib.check(res, "bad_pixels_correct")
    return out


def bad_pixels_list(handle):
    """Raster-ordered ``(x, y)`` list and clamp level held by a handle (introspection)."""
    lib = _lib.load()
    k = _lib.check(lib.rirb_bad_pixels_count(handle), "bad_pixels_count")
    xy = np.zeros((max(k, 1), 2), dtype=np.int32)
    clamp = ct.c_int(0)
    _lib.check(lib.rirb_bad_pixels_get(handle, _ptr(xy), k, ct.byref(clamp)), "bad_pixels_get")
    return xy[:k], clamp.value


class BadPixels:
    """Owner of a bad-pixel handle: detection at construction, ``correct`` per image, release on deletion
    (BadPixels.py:16-29)."""

    def __init__(self, first_image):
        self.handle = bad_pixels_create(first_image)

    def __del__(self):
        try:
            bad_pixels_destroy(self.handle)
        except Exception:
            pass

    def correct(self, img):
        return bad_pixels_correct(self.handle, img)

    def correct_batch(self, frames, out=None):
        return bad_pixels_correct_batch(self.handle, frames, out)

    def correct_gaussian_batch(self, frames, sigma=1.0, out=None, smoothed=None):
        return bad_pixels_correct_gaussian_batch(self.handle, frames, sigma, out, smoothed)


__all__ = [
    "translate",
    "translate_batch",
    "gaussian_filter",
    "gaussian_filter_batch",
    "find_median_pixel",
    "bad_pixels_create",
    "bad_pixels_correct",
    "bad_pixels_correct_batch",
    "bad_pixels_correct_gaussian_batch",
    "bad_pixels_destroy",
    "bad_pixels_list",
    "BadPixels",
]
