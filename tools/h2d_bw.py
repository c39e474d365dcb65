"""GPU box: host->device bandwidth of the 157 MB score tensor from pinned memory (torch pin_memory, cudaHostAlloc
write-combined / default; whole and in chunks): the ceiling of bench.py's e2e leg."""
import ctypes, numpy as np, torch, time
from cuda import cudart
n = 1024*9600*4
nbytes = n*4
def alloc(flags):
    err, ptr = cudart.cudaHostAlloc(nbytes, flags)
    assert err == cudart.cudaError_t.cudaSuccess, err
    buf = (ctypes.c_float * n).from_address(ptr)
    a = np.ctypeslib.as_array(buf)
    return torch.from_numpy(a)
x_def = torch.empty(n, dtype=torch.float32).pin_memory()
x_wc = alloc(cudart.cudaHostAllocWriteCombined)
x_pl = alloc(cudart.cudaHostAllocDefault)
print('pinned?', x_def.is_pinned(), x_wc.is_pinned(), x_pl.is_pinned())
d = torch.empty(n, dtype=torch.float32, device='cuda')
def t(x, name):
    for _ in range(3): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    best=1e9
    for _ in range(10):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    print(f"{name}: {best:.3f} ms = {nbytes/best/1e6:.1f} GB/s")
t(x_def,'torch pin_memory'); t(x_wc,'cudaHostAlloc write-combined'); t(x_pl,'cudaHostAlloc default')
# chunked 4 x
for k in (2,4,8):
    parts = x_def.chunk(k); dparts = d.chunk(k)
    torch.cuda.synchronize(); best=1e9
    for _ in range(10):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for p,q in zip(parts,dparts): q.copy_(p, non_blocking=True)
        e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    print(f"{k} chunks: {best:.3f} ms = {nbytes/best/1e6:.1f} GB/s")
