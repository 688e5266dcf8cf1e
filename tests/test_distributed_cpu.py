"""world_size-2 gloo test of the N>1 host path: GOP-aligned shards + the statistics all-reduce
(min / max / histogram), checked against whole-movie statistics from the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from librir_b200 import movie
    from oracle import oracle as O
    from tests.conftest import ir_movie

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mov = ir_movie(23, 16, 24)
        port = O.Port()
        shard = movie.shard_frames(len(mov), world, rank, gop=5)
        stats = movie.MovieStats("cpu")
        if shard.nframes:
            lo, hi, hist = port.movie_stats(mov[shard.start:shard.stop])  # stand-in for the CUDA kernel's output
            stats.minmax[0], stats.minmax[1] = int(lo), int(hi)
            stats.hist.copy_(torch.from_numpy(hist.astype(np.int64)))
            stats.count = int(mov[shard.start:shard.stop].size)
            stats._fresh = False
        stats.all_reduce()
        glo, ghi, ghist = port.movie_stats(mov)
        assert stats.min() == glo and stats.max() == ghi and stats.count == mov.size
        assert np.array_equal(stats.histogram(), ghist)
        assert stats.background() == port.get_background(mov.reshape(1, -1))
        # frame 0 reaches every rank
        f0 = torch.from_numpy(mov[0].view(np.int16).copy()) if rank == 0 else torch.zeros(mov[0].shape, dtype=torch.int16)
        movie.broadcast_first_frame(f0, src=0)
        assert np.array_equal(f0.numpy().view(np.uint16), mov[0])
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stats_all_reduce_gloo(tmp_path, world):
    import torch.multiprocessing as mp

    port_no = 29500 + (os.getpid() % 400) + world
    mp.spawn(_worker, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").exists()
