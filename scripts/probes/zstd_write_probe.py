import sys, time, os, numpy as np
sys.path.insert(0,'/root/repo')
from librir_b200 import tools
from oracle import container as oc
from tests.conftest import ir_movie
n=300; mov=ir_movie(n,512,640); ts=np.arange(n,dtype=np.int64)
for d in ('/dev/shm','/tmp'):
    for rep in range(2):
        t0=time.perf_counter()
        with tools.ZFileWriter(d+'/p.bin',640,512,threads=1) as w: w.add_images(mov,ts)
        b=time.perf_counter()-t0
        t0=time.perf_counter(); oc.ref_write_zfile(d+'/r.bin',mov,ts); a=time.perf_counter()-t0
        t0=time.perf_counter()
        with tools.ZFileWriter(d+'/p.bin',640,512,threads=1) as w:
            for i in range(n): w.add_image(mov[i],ts[i])
        c=time.perf_counter()-t0
        print(f"{d}: ref {n/a:.0f} fps  product batch {n/b:.0f} fps  product per-frame {n/c:.0f} fps", flush=True)
# compress only, python ctypes, same lib
z=oc.zstd()
import ctypes as ct
cap=z.ZSTD_compressBound(mov[0].nbytes); buf=ct.create_string_buffer(cap)
t0=time.perf_counter()
for i in range(n): z.ZSTD_compress(buf,cap,mov[i].ctypes.data_as(ct.c_void_p),mov[i].nbytes,2)
print(f"ZSTD_compress only: {n/(time.perf_counter()-t0):.0f} fps")
z.ZSTD_createCCtx.restype=ct.c_void_p; z.ZSTD_compressCCtx.argtypes=[ct.c_void_p,ct.c_void_p,ct.c_size_t,ct.c_void_p,ct.c_size_t,ct.c_int]; z.ZSTD_compressCCtx.restype=ct.c_size_t
cctx=z.ZSTD_createCCtx()
t0=time.perf_counter()
for i in range(n): z.ZSTD_compressCCtx(cctx,buf,cap,mov[i].ctypes.data_as(ct.c_void_p),mov[i].nbytes,2)
print(f"ZSTD_compressCCtx only: {n/(time.perf_counter()-t0):.0f} fps")
os.remove('/dev/shm/p.bin'); os.remove('/dev/shm/r.bin')
