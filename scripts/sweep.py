#!/usr/bin/env python
"""scripts/sweep.py -- frame-size sweep (BASELINE.json configs[4]): achieved
HBM GB/s of every kernel of the path, alone, against the measured copy peak, for frame sizes
320x256 .. 2048x2048.  Device-resident inputs, CUDA events, 3 warm-ups, median of `--reps`.

    python scripts/sweep.py [--gb 2.0] [--reps 7] > gpurun_out/sweep.jsonl

Under torchrun every rank sweeps its own GPU (weak scaling: the kernels share nothing) and rank 0
prints the max-over-ranks time per cell; one JSON line per (size, kernel) plus a markdown table on
stderr.  Not a bench.py line: evidence for profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = [(320, 256), (640, 512), (1024, 1024), (1280, 1024), (2048, 2048)]
BYTES_PER_PX = {"bp_correct": 4, "bp_correct_gaussian_fused": 8, "gaussian_u16_f32": 6, "gaussian_f32_f32": 8, "translate_u16": 4, "loader_motion_u16": 4, "loader_merge_minT_bp": 4, "loader_read_chain": 8,
                "precode_split": 4, "precode_delta_split": 4, "decode_delta_merge": 4, "stats_minmax_hist": 2}


def main():
    # the reader chain is merge pass + motion pass (8 B/px) by default, ONE pass over HBM (2 B/px in, 2 B/px out)
    # with RIRB_LOADER_FUSED=1
    BYTES_PER_PX["loader_read_chain"] = 4 if os.environ.get("RIRB_LOADER_FUSED", "0").startswith("1") else 8  # library default: two passes
    import torch
    import torch.distributed as dist

    from bench import hbm_peak
    from librir_b200 import movie, signal_processing as sp, video_io as vio

    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0, help="input bytes per kernel launch (GB) -- far beyond the 126 MB L2")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--only", default="", help="comma-separated kernel names to run (default: all)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = hbm_peak()
    rows = []
    for w, h in SIZES:
        npx = w * h
        n = max(50, int(args.gb * 1e9 / (npx * 2)) // 50 * 50)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        y = torch.arange(h, device=dev, dtype=torch.float32).view(1, h, 1)
        x = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, w)
        bg = 8000 + 2000 * torch.exp(-(((x - w / 2) / (0.23 * w)) ** 2) - ((y - h / 2) / (0.23 * h)) ** 2)
        frames = torch.empty((n, h, w), dtype=torch.uint16, device=dev)
        for a in range(0, n, 128):
            b = min(n, a + 128)
            f = (bg + 3.0 * torch.randn((b - a, h, w), generator=g, device=dev)).clamp_(0, 16383).to(torch.int16)
            frames.view(torch.int16)[a:b] = f
        bad = torch.randperm(npx, device=dev)[: max(2, round(1e-3 * npx))]
        fv = frames.view(torch.int16).view(n, -1)
        fv[:, bad[: len(bad) // 2]] = 0
        fv[:, bad[len(bad) // 2:]] = 16000
        dx = torch.rand(n, generator=g, device=dev) * 6 - 3
        dy = torch.rand(n, generator=g, device=dev) * 6 - 3
        bp = sp.BadPixels(frames[0])
        out16 = torch.empty_like(frames)
        out32 = torch.empty((n, h, w), dtype=torch.float32, device=dev)
        f32 = None
        lo = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
        hi = torch.empty_like(lo)
        stats = movie.MovieStats(dev)
        sx = dx.double().cpu().numpy()
        sy = dy.double().cpu().numpy()
        lbp = vio.LoaderBadPixels(frames[0].cpu().view(torch.int16).numpy().view("uint16"))
        vio.precode_movie(frames, 50, False, 0, out=(lo, hi))  # the planes the loader kernels read

        def run_f32():
            nonlocal f32
            if f32 is None:
                f32 = frames.view(torch.int16).to(torch.float32)
            sp.gaussian_filter_batch(f32, 1.0, out=out32)

        cases = {
            "bp_correct": lambda: bp.correct_batch(frames, out=out16),
            "bp_correct_gaussian_fused": lambda: bp.correct_gaussian_batch(frames, 1.0, out=out16, smoothed=out32),
            "gaussian_u16_f32": lambda: sp.gaussian_filter_batch(frames, 1.0, out=out32),
            "gaussian_f32_f32": run_f32,
            "translate_u16": lambda: sp.translate_batch(frames, dx, dy, "nearest", background=0, out=out16),
            "loader_motion_u16": lambda: vio.remove_motion(frames, sx, sy, meta_rows=3, out=out16),
            "loader_merge_minT_bp": lambda: vio.read_movie(lo, hi, lbp, 273, h - 3, None, None, out=out16),
            "loader_read_chain": lambda: vio.read_movie(lo, hi, lbp, 273, h - 3, sx, sy, out=out16),
            "precode_split": lambda: vio.precode_movie(frames, 50, False, 0, out=(lo, hi)),
            "precode_delta_split": lambda: vio.precode_movie(frames, 50, True, 0, out=(lo, hi)),
            "decode_delta_merge": lambda: vio.decode_movie(lo, hi, 50, True, 0, out=out16),
            "stats_minmax_hist": lambda: stats.update(frames),
        }
        only = [x for x in args.only.split(",") if x]
        for name, fn in cases.items():
            if only and name not in only:
                continue
            for _ in range(3):
                fn()
            times = []
            for _ in range(args.reps):  # four back-to-back launches per timed region, as bench.py's `configs` cells
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _i in range(4):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / 4)
            ms = statistics.median(times)
            t = torch.tensor([ms, min(times), max(times)], dtype=torch.float64, device=dev)
            per_rank = None
            if world > 1:
                allt = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(allt, t)
                per_rank = [[round(float(v), 4) for v in a] for a in allt]  # [median, min, max] of every rank
                ms = max(a[0] for a in per_rank)
            gbs = BYTES_PER_PX[name] * npx * n / (ms * 1e-3) / 1e9
            row = {"frame": [w, h], "frames_per_launch": n, "kernel": name, "ms": ms, "achieved_gbs": gbs, "peak_gbs": peak,
                   "frac": gbs / peak, "frames_per_s_per_gpu": n / (ms * 1e-3), "n_gpus": world, "peak_source": peak_src}
            if per_rank is not None:
                row["per_rank_ms_median_min_max"] = per_rank
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
        del frames, out16, out32, lo, hi, f32, bp, lbp
        torch.cuda.empty_cache()
    if rank == 0:
        names = list(BYTES_PER_PX)
        print("| frame | " + " | ".join(names) + " |", file=sys.stderr)
        print("|---|" + "---|" * len(names), file=sys.stderr)
        for w, h in SIZES:
            cells = []
            for nm in names:
                r = [x for x in rows if x["frame"] == [w, h] and x["kernel"] == nm]
                cells.append(f"{r[0]['achieved_gbs']:.0f} ({r[0]['frac']:.2f})" if r else "-")
            print(f"| {w}x{h} | " + " | ".join(cells) + " |", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
