"""How the histogram kernels behave when many pixels share a value (same-address shared-memory atomics): fused
pre-coder + statistics and the stand-alone statistics kernel on (a) the bench's synthetic movie, (b) a constant movie,
(c) a two-valued movie, (d) uniform noise.  2,000 frames of 640x512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from librir_b200 import movie, video_io as vio

n, h, w = 2000, 512, 640
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
y = torch.arange(h, device=dev, dtype=torch.float32).view(1, h, 1)
x = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, w)
bg = 8000 + 2000 * torch.exp(-(((x - w / 2) / (0.23 * w)) ** 2) - ((y - h / 2) / (0.23 * h)) ** 2)
cases = {}
f = torch.empty((n, h, w), dtype=torch.int16, device=dev)
for a in range(0, n, 100):
    f[a:a + 100] = (bg + 3.0 * torch.randn((100, h, w), generator=g, device=dev)).clamp_(0, 16383).to(torch.int16)
cases["bench movie"] = f.view(torch.uint16)
cases["constant 8000"] = torch.full((n, h, w), 8000, dtype=torch.int16, device=dev).view(torch.uint16)
two = torch.full((n, h, w), 8000, dtype=torch.int16, device=dev)
two[:, :, ::2] = 8001
cases["two values"] = two.view(torch.uint16)
cases["uniform noise"] = torch.randint(0, 16384, (n, h, w), generator=g, device=dev, dtype=torch.int16).view(torch.uint16)
lo = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
hi = torch.empty_like(lo)
for name, mov in cases.items():
    res = []
    for label, fn in (("precode+stats fused", lambda st: vio.precode_movie(mov, 50, True, 0, out=(lo, hi), stats=st)),
                      ("stats alone", lambda st: st.update(mov)),
                      ("precode alone", lambda st: vio.precode_movie(mov, 50, True, 0, out=(lo, hi)))):
        st = movie.MovieStats(dev)
        for _ in range(3):
            fn(st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn(st)
        e1.record()
        torch.cuda.synchronize()
        res.append(f"{label} {e0.elapsed_time(e1) / 5:.3f} ms")
    print(f"{name:15s}: " + ", ".join(res))
