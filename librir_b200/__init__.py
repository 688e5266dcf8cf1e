"""librir_b200 -- a B200-native (sm_100a CUDA) implementation of librir's per-frame hot path
on uint16 infrared movies: bad-pixel median correction, Gaussian filtering, sub-pixel
translation and the lossless writer's byte-plane (+ delta) pre-coder.

Layout mirrors the slice of the reference it replaces:

* ``librir_b200.signal_processing``  same functions/classes as ``librir.signal_processing``
  (``translate``, ``gaussian_filter``, ``find_median_pixel``, ``BadPixels``, ``bad_pixels_*``);
* ``librir_b200.video_io``           the writer's pre-coder (``H264Capture::AddFrame`` split,
  ``VideoGrabber::toArray`` merge) and the loader's per-frame hooks;
* ``librir_b200.movie``              device-resident movie shards: batched launches,
  frame-range sharding over ranks, NCCL all-reduce of min/max/histogram.

Everything computes in ``libs/libsignal_processing_b200.so`` (C ABI: ``include/librir_b200.h``).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
