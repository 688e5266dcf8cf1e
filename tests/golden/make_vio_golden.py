"""Generate tests/golden/vio_golden.npz from the REFERENCE ITSELF: oracle/_ref/libs/libvideo_io.so, i.e. the reference's
own video_io sources compiled by oracle/build_ref.sh over oracle/libav_stub.c (identity codec).  Run in the authoring
container, where /root/reference exists:

    python tests/golden/make_vio_golden.py

What is recorded (cases in tests/vio_cases.py; inputs are hash-built there, so only answers are stored):
  * split_<case>_*   : the planes H264Capture::AddFrame handed to the encoder and the pict_type it set, per frame
                       (h264_open_file / h264_add_image_lossless / h264_close_file)
  * lossy_<door>_<config>_* : per-frame CRC-32 of the frames the lossy pre-conditioner produced, every 20th frame in
                       full, and h264_get_low_errors / h264_get_high_errors (h264_add_image_lossy and h264_add_loss)
  * loader_<case>_<bp><motion>_* : per-frame CRC-32 (+ sampled full frames) of load_image after open_camera_file /
                       enable_bad_pixels / load_motion_correction_file / enable_motion_correction, the flagged-pixel
                       list of the handle and the mask of pixels whose 3x3 window is entirely flagged (undefined in the
                       reference: it reads a stale stack slot, IRFileLoader.cpp:786-795)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import refvio as rv  # noqa: E402
from tests import vio_cases as C  # noqa: E402


def u16_of(rec, codec, h, w):
    p = rec["planes"]
    if codec == "h264":
        return p[1].astype(np.uint16) | (p[2].astype(np.uint16) << 8)
    return p[0][:h, :w].astype(np.uint16) | (p[0][h:2 * h, :w].astype(np.uint16) << 8)


def main():
    assert rv.have_ref_vio(), "oracle/_ref/libs/libvideo_io.so missing: run oracle/build_ref.sh"
    out = {}
    tmp = tempfile.mkdtemp(prefix="vio_golden_")
    fn = os.path.join(tmp, "m.bin")
    # ---- a-7: split + key-frame rule -----------------------------------------------------------------------------
    for name, t, h, w, gop, codec in C.SPLIT_CASES:
        mov = C.movie(t, h, w, seed=len(name) + t)
        rv.write_lossless(fn, mov, codec=codec, gop=gop)
        d = rv.read_stub_file(fn)
        out[f"split_{name}_dims"] = np.array([d["width"], d["height"], d["pix_fmt"]], dtype=np.int32)
        out[f"split_{name}_pict"] = np.array([r["pict_type"] for r in d["records"]], dtype=np.int8)
        out[f"split_{name}_flags"] = np.array([r["flags"] for r in d["records"]], dtype=np.int8)
        for k in range(3):
            pl = np.stack([r["planes"][k] for r in d["records"]])
            if pl.any():
                out[f"split_{name}_p{k}"] = pl
            else:
                out[f"split_{name}_p{k}_zero"] = np.array(pl.shape, dtype=np.int32)
        assert d["comment"] == f"size:{w}x{h}"
    # ---- f-2: lossy pre-conditioner, both doors ------------------------------------------------------------------
    t, h, w, stop = C.LOSSY_SHAPE
    mov = C.lossy_movie()
    for door in C.LOSSY_DOORS:
        for cname, cfg in C.LOSSY_CONFIGS:
            s = rv.Saver(fn, w, h, stop, **cfg)
            if door == "add_loss":
                frames = np.stack([s.add_loss(mov[i]) for i in range(t)])
            else:
                for i in range(t):
                    s.add_image_lossy(mov[i], i * 1000)
            lo_e, hi_e = s.low_errors(), s.high_errors()
            s.close()
            if door != "add_loss":
                d = rv.read_stub_file(fn)
                frames = np.stack([u16_of(r, "h264", h, w) for r in d["records"]])
                cam = rv.Camera(fn)
                ga = cam.global_attributes()
                per_frame = []
                for i in range(t):
                    cam.load_image(i)
                    a = cam.attributes()
                    per_frame.append((int(a.get("BackgroundError", b"-1")), int(a.get("ForegroundError", b"-1"))))
                cam.close()
                # the per-frame attributes the saver stores are the same numbers (frame 0 has none)
                assert all(per_frame[i] == (lo_e[i], hi_e[i]) for i in range(1, t)), "attributes differ from get_*_errors"
                out[f"lossy_{door}_{cname}_min_t"] = np.array([int(ga.get("MIN_T", b"0")), int(ga.get("MIN_T_HEIGHT", b"0"))], dtype=np.int32)
            key = f"lossy_{door}_{cname}"
            out[key + "_crc"] = np.array([C.crc(f) for f in frames], dtype=np.uint32)
            out[key + "_full"] = frames[:: C.LOSSY_FULL_EVERY].copy()
            out[key + "_low"] = lo_e.astype(np.uint16)
            out[key + "_high"] = hi_e.astype(np.uint16)
    # ---- a-3 / a-6 / f-1: reader ---------------------------------------------------------------------------------
    port = O.Port()
    reg = os.path.join(tmp, "m.regfile")
    for case in C.LOADER_CASES:
        name, t, h, w, codec, min_t, min_th = case
        mov = C.loader_movie(case)
        ga = {}
        if min_t:
            ga["MIN_T"] = str(min_t)
        if min_th:
            ga["MIN_T_HEIGHT"] = str(min_th)
        rv.write_lossless(fn, mov, codec=codec, gop=10, global_attrs=ga)
        sx, sy = C.shifts(t, seed=t)
        rv.write_regfile(reg, sx, sy)
        for bp, mo in C.LOADER_MODES:
            cam = rv.Camera(fn)
            assert (cam.w, cam.h, cam.count) == (w, h, t)
            if bp:
                cam.enable_bad_pixels(True)
            if mo:
                cam.load_motion_correction_file(reg)
                cam.enable_motion_correction(True)
            order = list(range(t)) if not (bp and mo) else list(np.argsort(C.hash_noise((1, t), 77)[0], kind="stable"))  # seeks
            got = {}
            for i in order:
                got[int(i)] = cam.load_image(int(i))
            cam.close()
            frames = np.stack([got[i] for i in range(t)])
            key = f"loader_{name}_{bp}{mo}"
            out[key + "_crc"] = np.array([C.crc(f) for f in frames], dtype=np.uint32)
            out[key + "_full"] = frames[::10].copy()
        # the flagged set (setBadPixelsEnabled: badPixels(readImage(0), w, h - 3, 5), IRFileLoader.cpp:693-716) is not
        # exported by the reference; what IS pinned is its effect above.  Recorded here from the restatement (itself pinned
        # against the compiled badPixels<u16> by tests/test_oracle_vs_ref.py) so that the all-flagged windows can be masked.
        cam = rv.Camera(fn)
        first = cam.load_image(0)
        cam.close()
        xy = port.bad_pixels_detect(first[: h - 3])[0]
        bad = np.zeros((h - 3, w), dtype=bool)
        bad[xy[:, 1], xy[:, 0]] = True
        undefined = np.zeros((h, w), dtype=bool)
        for x, y in xy:
            x0 = 0 if x == 0 else (w - 3 if x == w - 1 else x - 1)
            y0 = 0 if y == 0 else (h - 6 if y == h - 4 else y - 1)
            undefined[y, x] = bad[y0:y0 + 3, x0:x0 + 3].all()
        out[f"loader_{name}_xy"] = xy.astype(np.int32)
        out[f"loader_{name}_undefined"] = np.packbits(undefined)
    path = os.path.join(ROOT, "tests", "golden", "vio_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
