import torch, time
n = 1<<30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    return reps*n/dt/1e9
for _ in range(2): run(True, True, 1)
print("H2D only GB/s", run(True, False)); print("D2H only GB/s", run(False, True)); print("both, GB/s per direction", run(True, True))
# chunked 64MB copies
c = 64<<20
def run_chunks(reps=3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps):
        for a in range(0, n, c):
            with torch.cuda.stream(s1): d_in[a:a+c].copy_(h_in[a:a+c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[a:a+c].copy_(d_out[a:a+c], non_blocking=True)
    torch.cuda.synchronize(); return reps*n/(time.perf_counter()-t0)/1e9
print("both, 64MB chunks", run_chunks())
