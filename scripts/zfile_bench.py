#!/usr/bin/env python
"""scripts/zfile_bench.py -- write / read throughput of the zstd movie file: the product's
thread-pooled container (rirb_z_write_images / rirb_z_read_images) next to the compiled reference's ZFile
(oracle/_ref, one frame per call on one thread -- what z_write_image / z_read_image do).  Host work; the GPU
only appears in the `device` rows, where the frames start / end in HBM.  Evidence for profiles/.

    python scripts/zfile_bench.py [--frames 400] > gpurun_out/zfile_bench.jsonl
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--clevel", type=int, default=2)
    args = ap.parse_args()
    from librir_b200 import tools
    from oracle import container as oc  # checker / baseline only
    from tests.conftest import ir_movie

    n, h, w = args.frames, 512, 640
    mov = ir_movie(n, h, w)
    ts = np.arange(n, dtype=np.int64) * 20_000_000
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="zfile_bench_")
    rows = []

    def rec(name, impl, threads, secs, extra=None):
        row = {"op": name, "impl": impl, "threads": threads, "frames": n, "frame": [w, h], "seconds": round(secs, 4),
               "frames_per_s": round(n / secs, 1), "MB_per_s": round(mov.nbytes / secs / 1e6, 1)}
        row.update(extra or {})
        rows.append(row)
        print(json.dumps(row), flush=True)

    have_ref = oc.have_ref()
    if have_ref:
        p = os.path.join(tmp, "ref.bin")
        t0 = time.perf_counter()
        oc.ref_write_zfile(p, mov, ts, clevel=args.clevel)
        rec("write", "reference", 1, time.perf_counter() - t0, {"file_MB": round(os.path.getsize(p) / 1e6, 1)})
        t0 = time.perf_counter()
        frames, _ = oc.ref_read_zfile(p)
        rec("read", "reference", 1, time.perf_counter() - t0)
        assert np.array_equal(frames, mov)
    for threads in (1, 0):
        p = os.path.join(tmp, f"prod{threads}.bin")
        t0 = time.perf_counter()
        with tools.ZFileWriter(p, w, h, clevel=args.clevel, threads=threads) as wr:
            wr.add_images(mov, ts)
        rec("write", "product", threads or cores, time.perf_counter() - t0, {"file_MB": round(os.path.getsize(p) / 1e6, 1)})
        if have_ref:
            assert open(p, "rb").read() == open(os.path.join(tmp, "ref.bin"), "rb").read(), "files differ from the reference's"
        t0 = time.perf_counter()
        with tools.ZFileReader(p, threads=threads) as rd:
            frames = rd.read_images()
        rec("read", "product", threads or cores, time.perf_counter() - t0)
        assert np.array_equal(frames, mov)
    try:
        import torch

        if torch.cuda.is_available():
            d = torch.from_numpy(mov.view(np.int16)).cuda().view(torch.uint16)
            p = os.path.join(tmp, "dev.bin")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with tools.ZFileWriter(p, w, h, clevel=args.clevel, threads=0) as wr:
                wr.add_images(d, ts)
            rec("write", "product, frames in HBM", cores, time.perf_counter() - t0)
            out = torch.empty_like(d)
            t0 = time.perf_counter()
            with tools.ZFileReader(p, threads=0) as rd:
                rd.read_images(0, n, out=out)
            torch.cuda.synchronize()
            rec("read", "product, frames to HBM", cores, time.perf_counter() - t0)
            assert torch.equal(out.view(torch.int16), d.view(torch.int16))
            # methods 2 / 3 (video_io.h:298-305): the GPU pre-coder in front of zstd; host frames and frames in HBM
            for method in (1, 2, 3):
                for where, src in (("host", mov), ("HBM", d)):
                    p = os.path.join(tmp, f"m{method}{where}.bin")
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    with tools.ZFileWriter(p, w, h, method=method, clevel=args.clevel, threads=0, gop=50) as wr:
                        wr.add_images(src, ts)
                    rec("write", f"product, method {method}, frames in {where}", cores, time.perf_counter() - t0,
                        {"file_MB": round(os.path.getsize(p) / 1e6, 1)})
                    t0 = time.perf_counter()
                    with tools.ZFileReader(p, threads=0) as rd:
                        if where == "host":
                            back = rd.read_images()
                        else:
                            rd.read_images(0, n, out=out)
                    torch.cuda.synchronize()
                    rec("read", f"product, method {method}, frames to {where}", cores, time.perf_counter() - t0)
                    if where == "host":
                        assert np.array_equal(back, mov)
                    else:
                        assert torch.equal(out.view(torch.int16), d.view(torch.int16))
    except ImportError:
        pass
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
