"""Developer aid (GPU box): pure-write and copy bandwidth of torch fill/copy kernels, for the encode roofline."""
import torch
n = 1024 * 9600 * 8  # floats written by one encode launch (labels + offsets, B = 1024)
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n // 2, dtype=torch.float32, device="cuda")
c = torch.empty(n // 2, dtype=torch.float32, device="cuda")
def timeit(fn, reps=20, inner=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / inner)
    return best
t = timeit(lambda: a.zero_()); print(f"zero_ {n*4/1e6:.0f} MB: {t:.1f} us = {n*4/t/1e6:.2f} TB/s written")
t = timeit(lambda: a.fill_(1.0)); print(f"fill_ {n*4/1e6:.0f} MB: {t:.1f} us = {n*4/t/1e6:.2f} TB/s written")
t = timeit(lambda: c.copy_(b)); print(f"copy_ {n*2/1e6:.0f} MB -> {n*2/1e6:.0f} MB: {t:.1f} us = {n*4/t/1e6:.2f} TB/s read+write")
big = torch.empty(1 << 29, dtype=torch.float32, device="cuda")
t = timeit(lambda: big.zero_(), reps=5, inner=3); print(f"zero_ {big.numel()*4/1e6:.0f} MB: {t:.1f} us = {big.numel()*4/t/1e6:.2f} TB/s written")
