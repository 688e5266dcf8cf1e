#!/usr/bin/env python
"""examples/pipeline_demo.py -- the per-frame path end to end on a synthetic IR movie (needs a B200).

    python examples/pipeline_demo.py [nframes]

1. the reference-shaped calls (numpy in, numpy out: what librir.signal_processing users write);
2. the same path device-resident, one launch per stage for the whole movie;
3. the lossless chain: GPU pre-coder -> host zstd -> back, bit-exact;
4. the reader's post-decode chain and the saver's lossy pre-conditioner;
5. registration: shifts estimated on the GPU (ECC), then applied -- a registered movie without leaving HBM.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from librir_b200 import entropy, movie, registration, signal_processing as sp, video_io as vio  # noqa: E402
from tests.conftest import ir_movie  # noqa: E402  (synthetic movie generator)


def main():
    import torch

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    h, w = 512, 640
    mov = ir_movie(n, h, w)
    rng = np.random.default_rng(0)
    dx, dy = rng.uniform(-3, 3, n).astype(np.float32), rng.uniform(-3, 3, n).astype(np.float32)

    # 1. frame by frame, like librir.signal_processing
    bp = sp.BadPixels(mov[0])
    t0 = time.perf_counter()
    for t in range(10):
        c = bp.correct(mov[t])
        g = sp.gaussian_filter(c, 1.0)
        r = sp.translate(c, dx[t], dy[t], "nearest")
    print(f"per-frame host calls: {10 / (time.perf_counter() - t0):8.0f} frames/s (each call crosses PCIe twice)")

    # 2. device-resident movie, batched launches
    d = torch.from_numpy(mov.view(np.int16)).cuda().view(torch.uint16)
    ddx, ddy = torch.from_numpy(dx).cuda(), torch.from_numpy(dy).cuda()
    stats = movie.MovieStats("cuda")
    c, r = torch.empty_like(d), torch.empty_like(d)  # reused buffers: cudaMalloc is the slow part of a first call
    g = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    planes = (torch.empty((n, h, w), dtype=torch.uint8, device="cuda"), torch.empty((n, h, w), dtype=torch.uint8, device="cuda"))
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bp.correct_batch(d, out=c)
        sp.gaussian_filter_batch(c, 1.0, out=g)
        sp.translate_batch(c, ddx, ddy, "nearest", 0, out=r)
        lo, hi = vio.precode_movie(r, gop=50, delta=True, stats=stats, out=planes)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
    print(f"device-resident batch: {n / el:8.0f} frames/s; registered movie min/max {stats.min()}/{stats.max()}, "
          f"median {stats.quantile(0.5)}, background {stats.background()}")

    # 3. lossless chain with the host entropy stage
    chunks = entropy.compress_movie(mov[:100], gop=50, delta=True, level=3)
    back = entropy.decompress_movie(chunks, h, w, gop=50, delta=True)
    size = sum(len(c[2]) + len(c[3]) for c in chunks)
    assert np.array_equal(back, mov[:100])
    print(f"lossless chain: {mov[:100].nbytes / size:.2f}x smaller than raw, bit-exact round trip")

    # 4. reader chain and lossy pre-conditioner
    lo8, hi8 = vio.precode_movie(mov[:50], delta=False)
    lbp = vio.LoaderBadPixels(mov[0])
    frames = vio.read_movie(lo8, hi8, lbp, min_T=0, min_T_height=0, shifts_x=dx[:50].astype(np.float64), shifts_y=dy[:50].astype(np.float64))
    pre = vio.LossyPreconditioner(w, h, h - 3)
    out, errors = pre.add_images(mov[:50])
    print(f"reader chain -> {frames.shape}; lossy pre-conditioner froze {float((out != mov[:50]).mean()) * 100:.1f} % of the pixels, "
          f"error bounds of the last frame {tuple(errors[-1])}")

    # 5. estimate the camera motion, then undo it
    from tests import ecc_cases

    shaky, sx, sy = ecc_cases.movie(40)
    ds = torch.from_numpy(shaky.view(np.int16)).cuda().view(torch.uint16)
    reg = registration.MaskedRegistratorECC()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reg.compute_movie(ds)
    el = time.perf_counter() - t0
    ex = torch.tensor(np.array(reg.x, dtype=np.float32)).cuda()
    ey = torch.tensor(np.array(reg.y, dtype=np.float32)).cuda()
    steady = sp.translate_batch(ds, -ex, -ey, "nearest", 0)
    err = max(np.max(np.abs(np.array(reg.x) - (sx - sx[0]))), np.max(np.abs(np.array(reg.y) - (sy - sy[0]))))
    resid = float((steady[1:, 100:400, 100:500].float() - steady[0, 100:400, 100:500].float()).abs().mean())
    print(f"registration: {len(shaky) / el:.0f} frames/s, largest error against the true camera path {err:.3f} px, "
          f"mean |frame - first frame| after undoing it {resid:.1f} counts")


if __name__ == "__main__":
    main()
