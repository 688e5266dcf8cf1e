"""oracle/container.py -- TEST INFRASTRUCTURE ONLY (never imported by librir_b200).

CPU restatement of the two file formats either side of the path, byte level, in
plain Python, plus ctypes doors onto the compiled reference for the same formats:

* attribute trailer   rir::FileAttributes -- FileAttributes.cpp:51-165 (string / map encoding),
                      :454-514 (writeIfDirty: layout of the trailer), :316-372 (open: how it is found)
* zstd movie file     ZFile.cpp:18-46 (headers), :255-296 (open for writing), :483-542 (record),
                      :410-452 (close: sample count + "positions" attribute), :124-253 (open for reading)

Pinned (tests/test_container.py): files written by ``write_zfile`` / ``build_trailer`` are read back
by the compiled reference (oracle/_ref/libs/libref_zfile.so = ZFile.cpp + zfile_shim.cpp, and the
attrs_* entries of libtools.so), files written by the compiled reference are parsed by ``read_zfile`` /
``parse_trailer``, and both agree byte for byte with the committed fixture
tests/golden/zfile_ref.bin, which the compiled reference wrote (tests/golden/make_golden.py).

zstd itself is the third-party dependency of this format (v1.5.5, extra/CMakeLists.txt:20-31; the
image's libzstd.so.1 is that version): compressed bytes are only compared between writers that use
the same library; across versions parity is on the decompressed content.
"""
from __future__ import annotations

import ctypes as ct
import os
import struct

import numpy as np

MAGIC = b"H264ATTRIBUTES"
MIN_COMPRESS = 1000          # MIN_SIZE_FOR_COMRPESSION, FileAttributes.cpp:26
CFLAG = 1 << 63

_z = None


def zstd():
    global _z
    if _z is None:
        z = ct.CDLL("libzstd.so.1")
        z.ZSTD_compressBound.restype = ct.c_size_t
        z.ZSTD_compressBound.argtypes = [ct.c_size_t]
        z.ZSTD_compress.restype = ct.c_size_t
        z.ZSTD_compress.argtypes = [ct.c_void_p, ct.c_size_t, ct.c_void_p, ct.c_size_t, ct.c_int]
        z.ZSTD_decompress.restype = ct.c_size_t
        z.ZSTD_decompress.argtypes = [ct.c_void_p, ct.c_size_t, ct.c_void_p, ct.c_size_t]
        z.ZSTD_isError.restype = ct.c_uint
        z.ZSTD_isError.argtypes = [ct.c_size_t]
        _z = z
    return _z


def zcompress(data: bytes, level: int) -> bytes:
    z = zstd()
    cap = z.ZSTD_compressBound(len(data))
    buf = ct.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), level)
    if z.ZSTD_isError(n):
        raise RuntimeError("zstd compress failed")
    return buf.raw[:n]


def zdecompress(data: bytes, raw_size: int) -> bytes:
    z = zstd()
    buf = ct.create_string_buffer(max(1, raw_size))
    n = z.ZSTD_decompress(buf, raw_size, data, len(data))
    if z.ZSTD_isError(n) or n != raw_size:
        raise RuntimeError("zstd decompress failed")
    return buf.raw[:raw_size]


# ---- attribute trailer ----------------------------------------------------------------------------
def _put_string(v: bytes) -> bytes:
    # FileAttributes.cpp:61-85,107-114: >= 1000 bytes -> zstd level 0, kept only when smaller
    if len(v) >= MIN_COMPRESS:
        c = zcompress(v, 0)
        if len(c) < len(v):
            return struct.pack("<Q", (len(c) + 8) | CFLAG) + struct.pack("<Q", len(v)) + c
    return struct.pack("<Q", len(v)) + v


def _put_map(m: dict) -> bytes:
    out = struct.pack("<Q", len(m))
    for k in sorted(m):  # std::map<std::string,...>: byte-wise key order
        out += _put_string(k) + _put_string(m[k])
    return out


def build_trailer(global_attrs: dict, frame_attrs: list, timestamps) -> bytes:
    """FileAttributes::writeIfDirty, FileAttributes.cpp:461-481.  Keys and values are bytes."""
    assert len(frame_attrs) == len(timestamps)
    s = _put_map(global_attrs)
    for m in frame_attrs:
        s += _put_map(m)
    for t in timestamps:
        s += struct.pack("<q", int(t))
    s += struct.pack("<Q", len(timestamps))
    s += struct.pack("<Q", len(s) + 8 + len(MAGIC))
    return s + MAGIC


class _Cur:
    def __init__(self, b, pos):
        self.b, self.pos = b, pos

    def u64(self):
        (v,) = struct.unpack_from("<Q", self.b, self.pos)
        self.pos += 8
        return v

    def string(self):
        n = self.u64()
        comp, n = bool(n & CFLAG), n & ~CFLAG
        raw = self.b[self.pos:self.pos + n]
        self.pos += n
        if not comp:
            return bytes(raw)
        (size,) = struct.unpack_from("<Q", raw, 0)
        return zdecompress(bytes(raw[8:]), size)

    def _map(self):
        m = {}
        for _ in range(self.u64()):
            k = self.string()
            m[k] = self.string()
        return m


def parse_trailer(data: bytes):
    """-> (global_attrs, frame_attrs, timestamps, trailer_size), or None when `data` does not end with a trailer."""
    if len(data) < 16 + len(MAGIC) or data[-len(MAGIC):] != MAGIC:
        return None
    count, tsize = struct.unpack_from("<QQ", data, len(data) - len(MAGIC) - 16)
    c = _Cur(data, len(data) - tsize)
    g = c._map()
    frames = [c._map() for _ in range(count)]
    times = np.array([struct.unpack_from("<q", data, c.pos + 8 * i)[0] for i in range(count)], dtype=np.int64)
    return g, frames, times, tsize


# ---- zstd movie file ---------------------------------------------------------------------------------
def zfile_headers(width, height, rate, samples, method=1) -> bytes:
    head = bytes([1, 1, method]) + bytes(125)                       # version, triggers, compression (ZFile.cpp:18-27,80-82)
    trig = [0, rate, samples, 0, 1, 1, 0, 3, 1, width, height]      # ZFile.cpp:29-45,275-283 (data_format ends up 3)
    return head + struct.pack("<11Q", *trig) + bytes(128 - 88)


def write_zfile(path, frames: np.ndarray, timestamps, rate=50, clevel=2) -> int:
    """z_open_file_write + z_write_image per frame + z_close_file; returns the bytes of headers + records."""
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    n, h, w = frames.shape
    body = b""
    positions = []
    pos = 256
    for i in range(n):
        c = zcompress(frames[i].tobytes(), clevel)
        positions.append(pos)
        body += struct.pack("<qI", int(timestamps[i]), len(c)) + c
        pos += 12 + len(c)
    trailer = build_trailer({b"positions": struct.pack(f"<{n}q", *positions)}, [{} for _ in range(n)], timestamps)
    with open(path, "wb") as f:
        f.write(zfile_headers(w, h, rate, n) + body + trailer)
    return pos


def read_zfile(path):
    """-> (frames uint16 [n,h,w], timestamps int64 [n], global attributes)."""
    data = open(path, "rb").read()
    if data[0] != 1 or data[1] != 1 or not (1 <= data[2] <= 3):
        raise RuntimeError("not a zstd movie file")
    trig = struct.unpack_from("<11Q", data, 128)
    rate, samples, w, h = trig[1], trig[2], trig[9], trig[10]
    if not (0 < w < 3000 and 0 < h < 3000 and 0 < rate < 1000):
        raise RuntimeError("not a zstd movie file")
    tr = parse_trailer(data)
    g = {}
    if tr is not None and b"positions" in tr[0] and len(tr[0][b"positions"]) == 8 * len(tr[2]):
        g, _, times, _ = tr                                          # ZFile.cpp:163-190
        positions = struct.unpack(f"<{len(times)}q", g[b"positions"])
    else:                                                            # ZFile.cpp:192-249: walk the records
        end = len(data) - (tr[3] if tr is not None else 0)
        positions, times, p = [], [], 256
        while p + 12 <= end and (samples == 0 or len(positions) < samples):
            ts, c = struct.unpack_from("<qI", data, p)
            positions.append(p)
            times.append(ts)
            p += 12 + c
        times = np.array(times, dtype=np.int64)
    frames = np.empty((len(positions), h, w), dtype=np.uint16)
    for i, p in enumerate(positions):
        _, c = struct.unpack_from("<qI", data, p)
        frames[i] = np.frombuffer(zdecompress(data[p + 12:p + 12 + c], w * h * 2), dtype=np.uint16).reshape(h, w)
    return frames, np.asarray(times, dtype=np.int64), g


# ---- the compiled reference -----------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref", "libs")


def have_ref() -> bool:
    return os.path.exists(os.path.join(_REF_DIR, "libref_zfile.so"))


_rz = None
_rt = None


def ref_tools():
    global _rt
    if _rt is None:
        t = ct.CDLL(os.path.join(_REF_DIR, "libtools.so"), mode=ct.RTLD_GLOBAL)
        t.attrs_open_file.argtypes = [ct.c_char_p]
        t.attrs_open_from_memory.argtypes = [ct.c_void_p, ct.c_int64]
        t.attrs_set_times.argtypes = [ct.c_int, ct.c_void_p, ct.c_int]
        t.attrs_timestamps.argtypes = [ct.c_int, ct.c_void_p]
        for n in ("attrs_global_attribute_name", "attrs_global_attribute_value"):
            getattr(t, n).argtypes = [ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]
        for n in ("attrs_frame_attribute_name", "attrs_frame_attribute_value"):
            getattr(t, n).argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]
        t.attrs_set_frame_attributes.argtypes = [ct.c_int, ct.c_int, ct.c_char_p, ct.c_void_p, ct.c_char_p, ct.c_void_p, ct.c_int]
        t.attrs_set_global_attributes.argtypes = [ct.c_int, ct.c_char_p, ct.c_void_p, ct.c_char_p, ct.c_void_p, ct.c_int]
        _rt = t
    return _rt


def ref_zfile():
    global _rz
    if _rz is None:
        ref_tools()
        z = ct.CDLL(os.path.join(_REF_DIR, "libref_zfile.so"))
        z.ref_z_open_file_write.restype = ct.c_void_p
        z.ref_z_open_file_write.argtypes = [ct.c_char_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int]
        z.ref_z_open_file_read.restype = ct.c_void_p
        z.ref_z_open_file_read.argtypes = [ct.c_char_p]
        z.ref_z_close_file.restype = ct.c_ulonglong
        z.ref_z_close_file.argtypes = [ct.c_void_p]
        z.ref_z_image_count.argtypes = [ct.c_void_p]
        z.ref_z_image_size.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p]
        z.ref_z_write_image.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_longlong]
        z.ref_z_read_image.argtypes = [ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_void_p]
        z.ref_z_get_timestamps.argtypes = [ct.c_void_p, ct.c_void_p]
        _rz = z
    return _rz


def ref_write_zfile(path, frames, timestamps, rate=50, clevel=2) -> int:
    z = ref_zfile()
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    n, h, w = frames.shape
    f = z.ref_z_open_file_write(str(path).encode(), w, h, rate, 1, clevel)
    if not f:
        raise RuntimeError("reference z_open_file_write failed")
    for i in range(n):
        if z.ref_z_write_image(f, frames[i].ctypes.data_as(ct.c_void_p), int(timestamps[i])) != 0:
            raise RuntimeError("reference z_write_image failed")
    return int(z.ref_z_close_file(f))


def ref_read_zfile(path):
    z = ref_zfile()
    f = z.ref_z_open_file_read(str(path).encode())
    if not f:
        raise RuntimeError("reference z_open_file_read failed")
    n = z.ref_z_image_count(f)
    w, h = ct.c_int(0), ct.c_int(0)
    z.ref_z_image_size(f, ct.byref(w), ct.byref(h))
    frames = np.empty((n, h.value, w.value), dtype=np.uint16)
    times = np.zeros(n, dtype=np.int64)
    per = np.zeros(n, dtype=np.int64)
    z.ref_z_get_timestamps(f, times.ctypes.data_as(ct.c_void_p))
    for i in range(n):
        t = ct.c_longlong(0)
        if z.ref_z_read_image(f, i, frames[i].ctypes.data_as(ct.c_void_p), ct.byref(t)) != 0:
            raise RuntimeError("reference z_read_image failed")
        per[i] = t.value
    z.ref_z_close_file(f)
    assert np.array_equal(per, times)
    return frames, times


def _blob(fn, *ids) -> bytes:
    n = ct.c_int(0)
    r = fn(*ids, None, ct.byref(n))
    buf = ct.create_string_buffer(max(1, n.value))
    if r == -2:
        r = fn(*ids, buf, ct.byref(n))
    if r < 0:
        raise RuntimeError("reference attrs getter failed")
    return buf.raw[: n.value]


def ref_read_attrs(path):
    """-> (global_attrs, frame_attrs, timestamps) through the compiled reference's attrs_* entries."""
    t = ref_tools()
    h = t.attrs_open_file(str(path).encode())
    if h <= 0:
        raise RuntimeError("reference attrs_open_file failed")
    n = t.attrs_image_count(h)
    times = np.zeros(n, dtype=np.int64)
    t.attrs_timestamps(h, times.ctypes.data_as(ct.c_void_p))
    g = {_blob(t.attrs_global_attribute_name, h, i): _blob(t.attrs_global_attribute_value, h, i)
         for i in range(t.attrs_global_attribute_count(h))}
    frames = [{_blob(t.attrs_frame_attribute_name, h, f, i): _blob(t.attrs_frame_attribute_value, h, f, i)
               for i in range(t.attrs_frame_attribute_count(h, f))} for f in range(n)]
    t.attrs_close(h)  # nothing changed: tableSize != 0, nothing is rewritten
    return g, frames, times


def ref_write_attrs(path, global_attrs: dict, frame_attrs: list, timestamps):
    t = ref_tools()
    h = t.attrs_open_file(str(path).encode())
    if h <= 0:
        raise RuntimeError("reference attrs_open_file failed")
    times = np.ascontiguousarray(timestamps, dtype=np.int64)
    t.attrs_set_times(h, times.ctypes.data_as(ct.c_void_p), len(times))

    def pack(m):
        ks, vs = list(m.keys()), list(m.values())
        kl = np.array([len(k) for k in ks], dtype=np.int32)
        vl = np.array([len(v) for v in vs], dtype=np.int32)
        return b"".join(ks), kl, b"".join(vs), vl

    k, kl, v, vl = pack(global_attrs)
    t.attrs_set_global_attributes(h, k, kl.ctypes.data_as(ct.c_void_p), v, vl.ctypes.data_as(ct.c_void_p), len(global_attrs))
    for i, m in enumerate(frame_attrs):
        k, kl, v, vl = pack(m)
        t.attrs_set_frame_attributes(h, i, k, kl.ctypes.data_as(ct.c_void_p), v, vl.ctypes.data_as(ct.c_void_p), len(m))
    t.attrs_close(h)
