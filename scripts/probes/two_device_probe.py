"""Two-GPU probe: a handle created on device 0 used from device 1 and back (prints the library's message at each step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from librir_b200 import _lib, signal_processing as sp
from tests.conftest import ir_frame

lib = _lib.load()
print("devices", lib.rirb_device_count())
img = ir_frame(64, 96, 5)
print("set 0", lib.rirb_set_device(0))
bp = sp.BadPixels(img)
want = bp.correct(img)
print("set 1", lib.rirb_set_device(1))
out = np.empty_like(img)
r = lib.bad_pixels_correct(bp.handle, sp._ptr(img), sp._ptr(out))
print("correct with dev-0 handle on dev 1:", r, _lib.last_error())
h1 = lib.bad_pixels_create(sp._ptr(img), 96, 64)
print("create on dev 1:", h1, _lib.last_error() if h1 <= 0 else "")
r = lib.bad_pixels_correct(h1, sp._ptr(img), sp._ptr(out))
print("correct on dev 1:", r, _lib.last_error() if r < 0 else "", np.array_equal(out, want))
print("set 0", lib.rirb_set_device(0))
r = lib.bad_pixels_correct(bp.handle, sp._ptr(img), sp._ptr(out))
print("correct with dev-0 handle back on dev 0:", r, _lib.last_error() if r < 0 else "", np.array_equal(out, want))
