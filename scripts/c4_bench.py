#!/usr/bin/env python
"""scripts/c4_bench.py -- BASELINE.json configs[3]: 1024x1024 uint16 movie, registration-style sub-pixel
translate with per-frame shifts + min/max/histogram statistics + their NCCL all-reduce, frame-sharded.
A step = one 1,000-frame chunk per GPU through translate -> stats -> all-reduce (the all-reduce is INSIDE the
timed step here: the config names it).  One JSON line on rank 0; evidence for profiles/, not the bench.py line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    from bench import hbm_peak
    from librir_b200 import movie, signal_processing as sp

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = H = 1024
    chunk, steps, warmup = 1000, 10, 3
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.empty((chunk, H, W), dtype=torch.uint16, device=dev)
    for a in range(0, chunk, 100):
        frames.view(torch.int16)[a:a + 100] = (9000 + 3.0 * torch.randn((100, H, W), generator=g, device=dev)).clamp_(0, 16383).to(torch.int16)
    dx = torch.rand(chunk, generator=g, device=dev) * 6 - 3
    dy = torch.rand(chunk, generator=g, device=dev) * 6 - 3
    out = torch.empty_like(frames)
    stats = movie.MovieStats(dev)

    def step():
        sp.translate_batch(frames, dx, dy, "nearest", background=0, out=out)
        stats.update(out)
        stats.all_reduce()

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    if rank == 0:
        peak, src = hbm_peak()
        gbs = 6 * W * H * chunk / (ms * 1e-3) / 1e9  # translate 4 B/px + stats 2 B/px
        print(json.dumps({"workload": "C4: 1024x1024 u16, translate (per-frame shifts, nearest) + stats + NCCL all-reduce per 1,000-frame chunk",
                          "n_gpus": world, "frames_per_step_per_gpu": chunk, "ms_per_step": ms, "value": chunk * world / (ms * 1e-3),
                          "unit": "frames/s", "achieved_gbs_per_gpu": gbs, "frac_of_peak": gbs / peak, "peak_source": src}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
