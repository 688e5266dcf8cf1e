"""Host entropy stage behind the pre-coder: zstd, as the reference calls it.

OUTSIDE the timed GPU path (BASELINE.json north star): the reference compresses on
the host with zstd v1.5.5 through four thin wrappers in libtools (``zstd_compress_bound``,
``zstd_compress``, ``zstd_decompress_bound``, ``zstd_decompress``: tools.cpp:352-376, Python side
``librir/tools/rir_tools.py:12-78``).  Same four names and meanings here, bound with ctypes to the
system ``libzstd.so.1`` (the library the reference links), plus ``compress_movie`` /
``decompress_movie`` which put the GPU pre-coder in front: frames -> ``rirb_precode_movie`` -> one
zstd frame per GOP and byte plane.  This is the chain the lossless numbers in SURVEY.md 8c quote
(byte planes + temporal delta compress better than raw frames); it is a host utility, not a
container format.
"""
from __future__ import annotations

import ctypes as ct
import ctypes.util

import numpy as np

from . import video_io as vio

_z = None


def _zstd():
    global _z
    if _z is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        z = ct.CDLL(name)
        z.ZSTD_compressBound.restype = ct.c_size_t
        z.ZSTD_compressBound.argtypes = [ct.c_size_t]
        z.ZSTD_compress.restype = ct.c_size_t
        z.ZSTD_compress.argtypes = [ct.c_void_p, ct.c_size_t, ct.c_void_p, ct.c_size_t, ct.c_int]
        z.ZSTD_decompress.restype = ct.c_size_t
        z.ZSTD_decompress.argtypes = [ct.c_void_p, ct.c_size_t, ct.c_void_p, ct.c_size_t]
        z.ZSTD_getFrameContentSize.restype = ct.c_ulonglong
        z.ZSTD_getFrameContentSize.argtypes = [ct.c_void_p, ct.c_size_t]
        z.ZSTD_isError.restype = ct.c_uint
        z.ZSTD_isError.argtypes = [ct.c_size_t]
        _z = z
    return _z


def zstd_compress_bound(size: int) -> int:
    """tools.cpp:352-355."""
    return int(_zstd().ZSTD_compressBound(int(size)))


def zstd_compress(src, level: int = 0) -> bytes:
    """tools.cpp:363-369 / rir_tools.py:21-48: bytes-like in, compressed bytes out; RuntimeError on failure."""
    z = _zstd()
    buf = np.frombuffer(memoryview(src).cast("B"), dtype=np.uint8) if not isinstance(src, np.ndarray) else np.ascontiguousarray(src).view(np.uint8).reshape(-1)
    out = np.empty(zstd_compress_bound(buf.size), dtype=np.uint8)
    ret = z.ZSTD_compress(out.ctypes.data, out.size, buf.ctypes.data, buf.size, int(level))
    if z.ZSTD_isError(ret):
        raise RuntimeError("'zstd_compress': unknown error")
    return out[:ret].tobytes()


def zstd_decompress_bound(src) -> int:
    """tools.cpp:356-362: content size recorded in the frame header, -1 if it is not a zstd frame."""
    z = _zstd()
    b = bytes(src)
    ret = z.ZSTD_getFrameContentSize(b, len(b))
    return -1 if ret >= 0xFFFFFFFFFFFFFFFE else int(ret)


def zstd_decompress(src) -> bytes:
    """tools.cpp:370-376 / rir_tools.py:51-78; RuntimeError on garbage (tests/python/test_rir.py:47-74)."""
    z = _zstd()
    b = bytes(src)
    n = zstd_decompress_bound(b)
    if n < 0:
        raise RuntimeError("'zstd_decompress': unknown error")
    out = np.empty(max(n, 1), dtype=np.uint8)
    ret = z.ZSTD_decompress(out.ctypes.data, n, b, len(b))
    if z.ZSTD_isError(ret) or ret != n:
        raise RuntimeError("'zstd_decompress': unknown error")
    return out[:n].tobytes()


def compress_movie(frames, gop: int = vio.DEFAULT_GOP, delta: bool = True, level: int = 3):
    """uint16 ``[n, h, w]`` -> list of ``(first_frame, nframes, lo_bytes, hi_bytes)`` chunks, one per GOP:
    GPU pre-coder (byte planes, optional temporal delta) then host zstd per plane."""
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    n = frames.shape[0]
    lo, hi = vio.precode_movie(frames, gop=gop, delta=delta, first_frame=0)
    return [(a, min(gop, n - a), zstd_compress(lo[a:a + gop], level), zstd_compress(hi[a:a + gop], level)) for a in range(0, n, gop)]


def decompress_movie(chunks, height: int, width: int, gop: int = vio.DEFAULT_GOP, delta: bool = True) -> np.ndarray:
    """Inverse of :func:`compress_movie`: host zstd, then the GPU inverse pre-coder."""
    n = sum(c[1] for c in chunks)
    lo = np.empty((n, height, width), np.uint8)
    hi = np.empty((n, height, width), np.uint8)
    for a, m, zl, zh in chunks:
        lo[a:a + m] = np.frombuffer(zstd_decompress(zl), np.uint8).reshape(m, height, width)
        hi[a:a + m] = np.frombuffer(zstd_decompress(zh), np.uint8).reshape(m, height, width)
    return vio.decode_movie(lo, hi, gop=gop, delta=delta, first_frame=0)


__all__ = ["zstd_compress_bound", "zstd_compress", "zstd_decompress_bound", "zstd_decompress", "compress_movie", "decompress_movie"]
