import sys, os, torch, statistics
sys.path.insert(0, os.getcwd())
from bench import synth_movie_torch
from librir_b200 import signal_processing as sp
dev=torch.device("cuda",0)
fr=synth_movie_torch(3000,0,dev); out=torch.empty((3000,512,640),dtype=torch.float32,device=dev)
for sigma in (0.5,1.0,1.5,2.0,2.4):
    for _ in range(3): sp.gaussian_filter_batch(fr,sigma,out=out)
    ts=[]
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); sp.gaussian_filter_batch(fr,sigma,out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms=statistics.median(ts); gbs=6*640*512*3000/(ms*1e-3)/1e9
    print(f"sigma {sigma} radius {max(1,int(2*sigma))}: {ms:.3f} ms  {gbs:.0f} GB/s  frac {gbs/6534.8:.2f}")
