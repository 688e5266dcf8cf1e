"""ctypes front-end of the compiled reference's video_io library (oracle/_ref/libs/libvideo_io.so).

TEST INFRASTRUCTURE ONLY.  ``oracle/build_ref.sh`` compiles the reference's own video_io sources
(h264.cpp, IRFileLoader.cpp, video_io.cpp, ...) where they lie under /root/reference and links them
against ``oracle/libav_stub.c`` -- an identity "codec" + trivial container standing in for
ffmpeg 7.1 / libx264 / kvazaar, which the reference's build fetches from the network.  Everything
above the bitstream therefore RUNS as the reference wrote it:

* the lossless writer's byte-plane split and key-frame rule (H264Capture::AddFrame,
  h264.cpp:1022-1238): the file the writer leaves behind holds the planes it handed to the
  encoder -- ``read_stub_file`` parses it;
* the lossy pre-conditioner, both doors: ``h264_add_image_lossy`` (addImageLossyNoCamera,
  h264.cpp:2253-2424) and ``h264_add_loss`` (addLoss, :2426-2607) with ``h264_get_low/high_errors``;
* the reader's post-decode chain: ``open_camera_file`` -> ``enable_bad_pixels`` ->
  ``load_motion_correction_file`` / ``enable_motion_correction`` -> ``load_image``
  (IRFileLoader::readImage, IRFileLoader.cpp:1148-1247 incl. toArray's merge, += MIN_T,
  removeBadPixels, removeMotion).

Argument conventions follow librir/video_io/rir_video_io.py.
"""
from __future__ import annotations

import ctypes as ct
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref", "libs")
_PATH = os.environ.get("REFVIO_LIB", os.path.join(_REF_DIR, "libvideo_io.so"))

_lib = None


def have_ref_vio() -> bool:
    return os.path.exists(_PATH)


def lib() -> ct.CDLL:
    global _lib
    if _lib is None:
        for dep in ("libtools.so", "libgeometry.so", "libsignal_processing.so"):
            ct.CDLL(os.path.join(_REF_DIR, dep), mode=ct.RTLD_GLOBAL)
        _lib = ct.CDLL(_PATH)
        _lib.h264_close_file.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def _pack_attrs(attrs):
    """The (count, keys, key_lens, values, value_lens) convention of rir_video_io.py:533-625."""
    attrs = attrs or {}
    keys = b"".join(k.encode() if isinstance(k, str) else bytes(k) for k in attrs)
    vals = b"".join(v.encode() if isinstance(v, str) else bytes(v) for v in attrs.values())
    klens = np.array([len(k.encode() if isinstance(k, str) else bytes(k)) for k in attrs] or [0], dtype=np.int32)
    vlens = np.array([len(v.encode() if isinstance(v, str) else bytes(v)) for v in attrs.values()] or [0], dtype=np.int32)
    return len(attrs), keys, klens, vals, vlens


class Saver:
    """h264_open_file ... h264_close_file (video_io.h:222-280)."""

    def __init__(self, filename, width, height, lossy_height=None, **params):
        self.filename = str(filename)
        self.w, self.h = int(width), int(height)
        L = lib()
        self.handle = L.h264_open_file(self.filename.encode(), self.w, self.h, self.h if lossy_height is None else int(lossy_height))
        if self.handle <= 0:
            raise RuntimeError("h264_open_file failed")
        for k, v in params.items():
            self.set_parameter(k, v)

    def set_parameter(self, key, value):
        if lib().h264_set_parameter(self.handle, str(key).encode(), str(value).encode()) < 0:
            raise RuntimeError(f"h264_set_parameter({key}) failed")

    def set_global_attributes(self, attrs):
        n, keys, klens, vals, vlens = _pack_attrs(attrs)
        if lib().h264_set_global_attributes(self.handle, n, keys, _p(klens), vals, _p(vlens)) < 0:
            raise RuntimeError("h264_set_global_attributes failed")

    def _add(self, fn, image, timestamp, attrs):
        img = np.ascontiguousarray(image, dtype=np.uint16)
        assert img.shape == (self.h, self.w)
        n, keys, klens, vals, vlens = _pack_attrs(attrs)
        f = getattr(lib(), fn)
        f.argtypes = [ct.c_int, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_char_p, ct.c_void_p, ct.c_char_p, ct.c_void_p]
        if f(self.handle, _p(img), int(timestamp), n, keys, _p(klens), vals, _p(vlens)) < 0:
            raise RuntimeError(fn + " failed")

    def add_image_lossless(self, image, timestamp, attrs=None):
        self._add("h264_add_image_lossless", image, timestamp, attrs)

    def add_image_lossy(self, image, timestamp, attrs=None):
        self._add("h264_add_image_lossy", image, timestamp, attrs)

    def add_loss(self, image):
        """H264_Saver::addLoss: returns the pre-conditioned frame (the call works in place)."""
        img = np.array(image, dtype=np.uint16, order="C")
        f = lib().h264_add_loss
        f.argtypes = [ct.c_int, ct.c_void_p]
        if f(self.handle, _p(img)) < 0:
            raise RuntimeError("h264_add_loss failed")
        return img

    def _errors(self, fn):
        f = getattr(lib(), fn)
        f.argtypes = [ct.c_int, ct.c_void_p, ct.c_void_p]
        size = ct.c_int(0)
        dummy = np.zeros(1, dtype=np.uint16)
        r = f(self.handle, _p(dummy), ct.byref(size))
        if r == 0:
            return dummy[: size.value].copy()
        out = np.zeros(size.value, dtype=np.uint16)
        if f(self.handle, _p(out), ct.byref(size)) < 0:
            raise RuntimeError(fn + " failed")
        return out

    def low_errors(self):
        return self._errors("h264_get_low_errors")

    def high_errors(self):
        return self._errors("h264_get_high_errors")

    def close(self):
        if self.handle:
            lib().h264_close_file(self.handle)
            self.handle = 0


class Camera:
    """open_camera_file ... close_camera (video_io.h:30-152)."""

    def __init__(self, filename):
        L = lib()
        fmt = ct.c_int(0)
        self.handle = L.open_camera_file(str(filename).encode(), ct.byref(fmt))
        if self.handle <= 0:
            raise RuntimeError("open_camera_file failed")
        self.file_format = fmt.value
        w, h = ct.c_int(0), ct.c_int(0)
        L.get_image_size(self.handle, ct.byref(w), ct.byref(h))
        self.w, self.h = w.value, h.value
        self.count = L.get_image_count(self.handle)

    def enable_bad_pixels(self, enable=True):
        if lib().enable_bad_pixels(self.handle, int(enable)) < 0:
            raise RuntimeError("enable_bad_pixels failed")

    def load_motion_correction_file(self, filename):
        if lib().load_motion_correction_file(self.handle, str(filename).encode()) < 0:
            raise RuntimeError("load_motion_correction_file failed")

    def enable_motion_correction(self, enable=True):
        if lib().enable_motion_correction(self.handle, int(enable)) < 0:
            raise RuntimeError("enable_motion_correction failed")

    def load_image(self, pos, calibration=0):
        out = np.zeros((self.h, self.w), dtype=np.uint16)
        f = lib().load_image
        f.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_void_p]
        if f(self.handle, int(pos), int(calibration), _p(out)) < 0:
            raise RuntimeError("load_image failed")
        return out

    def image_time(self, pos):
        t = ct.c_int64(0)
        lib().get_image_time(self.handle, int(pos), ct.byref(t))
        return t.value

    def _attrs(self, count_fn, get_fn):
        L = lib()
        n = getattr(L, count_fn)(self.handle)
        out = {}
        for i in range(max(n, 0)):
            k = ct.create_string_buffer(4096)
            v = ct.create_string_buffer(1 << 20)
            kl, vl = ct.c_int(4096), ct.c_int(1 << 20)
            if getattr(L, get_fn)(self.handle, i, k, ct.byref(kl), v, ct.byref(vl)) == 0:
                out[k.raw[: kl.value].decode()] = v.raw[: vl.value]
        return out

    def global_attributes(self):
        return self._attrs("get_global_attribute_count", "get_global_attribute")

    def attributes(self):
        """Attributes of the frame read last."""
        return self._attrs("get_attribute_count", "get_attribute")

    def close(self):
        if self.handle:
            lib().close_camera(self.handle)
            self.handle = 0


def write_regfile(path, shift_x, shift_y):
    """The motion-correction file IRFileLoader::loadTranslationFile reads (IRFileLoader.cpp:822-847):
    one header line, then one tab-separated row of 4 columns per frame; columns 1 and 2 are x and y."""
    with open(path, "w") as f:
        f.write("frame\tx\ty\tconfidence\n")
        for i, (x, y) in enumerate(zip(shift_x, shift_y)):
            f.write(f"{i}\t{float(x)!r}\t{float(y)!r}\t1\n")


# ---------------------------------------------------------------------------------------------------
# the stub container (layout documented at the top of oracle/libav_stub.c)
# ---------------------------------------------------------------------------------------------------
AV_PIX_FMT_YUV420P = 0
AV_PIX_FMT_YUV444P = 5
AV_PICTURE_TYPE_I = 1
AV_PKT_FLAG_KEY = 1


def read_stub_file(path):
    """Parse a file written through the stub: header fields, per-record flags / pict_type / pts / dts and the
    three planes of every frame exactly as AddFrame filled them (rows un-padded)."""
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"RIR1ftyp", "not a stub container"
    codec_id, w, h, fmt, fps = struct.unpack_from("<5i", data, 8)
    (nb,) = struct.unpack_from("<q", data, 28)
    (clen,) = struct.unpack_from("<i", data, 36)
    comment = data[40: 40 + clen].decode()
    pos = 40 + clen
    if fmt == AV_PIX_FMT_YUV420P:
        dims = [(h, w), ((h + 1) // 2, (w + 1) // 2), ((h + 1) // 2, (w + 1) // 2)]
    else:
        dims = [(h, w)] * 3
    recs = []
    while True:
        (magic,) = struct.unpack_from("<I", data, pos)
        if magic != 0x52454346:
            assert magic == 0x444E4546, "damaged stub container"
            break
        flags, pict = struct.unpack_from("<2i", data, pos + 4)
        pts, dts, dur, ppos = struct.unpack_from("<4q", data, pos + 12)
        (size,) = struct.unpack_from("<I", data, pos + 44)
        o = pos + 48
        planes = []
        for ph, pw in dims:
            planes.append(np.frombuffer(data, dtype=np.uint8, count=ph * pw, offset=o).reshape(ph, pw))
            o += ph * pw
        assert o == pos + 48 + size
        recs.append(dict(flags=flags, pict_type=pict, pts=pts, dts=dts, duration=dur, pos=ppos, planes=planes))
        pos = o
    assert len(recs) == nb
    return dict(codec_id=codec_id, width=w, height=h, pix_fmt=fmt, fps=fps, comment=comment, records=recs, video_end=pos + 4)


def write_lossless(path, movie, codec="h264", gop=50, timestamps=None, global_attrs=None, frame_attrs=None, **params):
    """movie[t][h][w] through h264_open_file / h264_add_image_lossless / h264_close_file."""
    t, h, w = movie.shape
    s = Saver(path, w, h, h, codec=codec, GOP=gop, **params)
    if global_attrs:
        s.set_global_attributes(global_attrs)
    for i in range(t):
        s.add_image_lossless(movie[i], i if timestamps is None else timestamps[i], frame_attrs[i] if frame_attrs else None)
    s.close()
