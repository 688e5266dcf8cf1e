#!/usr/bin/env python
"""scripts/probes/hostlink_probe.py -- what the HOST side of the box can feed N GPUs at once.

Run under torchrun with N ranks (one per GPU), or alone.  Every rank measures, with all ranks active at the same time
(barriers around each phase), pinned-memory H2D, D2H and both-at-once bandwidth on its own GPU, plus a plain host memcpy
between two pinned buffers; rank 0 prints one JSON line with per-rank and aggregate GB/s, the NUMA node of every GPU and
the CPUs each rank may run on.  `--bind` pins a rank's CPU affinity (and so the first touch of its pinned buffers) to the
CPUs that sysfs lists for its GPU's NUMA node before anything is allocated.  Evidence for profiles/, not a bench line."""
import argparse
import json
import os
import sys
import time


def gpu_numa(bus_id):
    try:
        p = f"/sys/bus/pci/devices/{bus_id.lower()}/numa_node"
        return int(open(p).read().strip())
    except Exception:
        return None


def numa_cpus(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = []
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus += list(range(int(a), int(b) + 1))
            else:
                cpus.append(int(part))
        return cpus
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--bind", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    props = torch.cuda.get_device_properties(local)
    bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0" if hasattr(props, "pci_bus_id") else None
    node = gpu_numa(bus) if bus else None
    bound = None
    if args.bind and node is not None and node >= 0:
        cpus = numa_cpus(node)
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)
                bound = len(cpus)
            except Exception:
                bound = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(h2d, d2h, memcpy=False):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            if memcpy:
                h_out.copy_(h_in)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return args.reps * n / dt / 1e9

    run(True, True)
    res = {"h2d": run(True, False), "d2h": run(False, True), "both_per_direction": run(True, True), "host_memcpy": run(False, False, True)}
    vals = torch.tensor([res["h2d"], res["d2h"], res["both_per_direction"], res["host_memcpy"], float(node if node is not None else -9),
                         float(len(os.sched_getaffinity(0)))], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        rows = [[float(x) for x in v.cpu()] for v in allv]
        out = {"probe": "hostlink", "ranks": world, "mb_per_copy": args.mb, "bind": bool(args.bind), "cpus_bound_rank0": bound,
               "host_cpus": os.cpu_count(),
               "per_rank": [{"h2d": round(r[0], 1), "d2h": round(r[1], 1), "both_per_direction": round(r[2], 1), "host_memcpy": round(r[3], 1),
                             "gpu_numa_node": int(r[4]), "cpus_allowed": int(r[5])} for r in rows],
               "aggregate": {"h2d": round(sum(r[0] for r in rows), 1), "d2h": round(sum(r[1] for r in rows), 1),
                             "both_per_direction": round(sum(r[2] for r in rows), 1), "host_memcpy": round(sum(r[3] for r in rows), 1)}}
        try:
            out["numa_nodes"] = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except Exception:
            pass
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
