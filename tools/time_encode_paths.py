"""Developer aid (GPU box): time ssdx_encode at the bench shape for every encode driver (ssdx_set_option), and print the
per-phase cycle breakdown of the grid fused kernel when the debug build exists (python build.py --debug-timers).
usage: [SSDX_LIB=...] python tools/time_encode_paths.py [B] [reps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
sys.argv = ["bench.py"]
import bench
anchors, labels, boxes, counts = bench.make_inputs(B, 1995)
if os.environ.get("ZERO_GT"):  # all-background images: the output stream alone
    counts[:] = 0
coder = ssdseglib.datacoder.DataEncoderDecoder(4, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
gb, gl, gc = torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda()
f32 = dict(dtype=torch.float32, device="cuda:0")
o_labels, o_offsets = torch.empty((B, 9600, 4), **f32), torch.empty((B, 9600, 4), **f32)
ws = torch.empty(max(sx.lib.ssdx_encode_workspace_bytes(9600, B, 16), 256), dtype=torch.uint8, device="cuda:0")
std = sx.std_array((0.1, 0.1, 0.2, 0.2))
sp = torch.cuda.current_stream().cuda_stream
h = coder._anchors.handle
fn = lambda: sx.check(sx.lib.ssdx_encode(h, gb.data_ptr(), gl.data_ptr(), gc.data_ptr(), B, 16, 4, 0.5, std,
                                         o_labels.data_ptr(), o_offsets.data_ptr(), None, ws.data_ptr(), ws.numel(), sp))
INNER = 10
for path in ("grid_fused", "grid_split", "tile_fused", "tile_tiled"):
    sx.set_option("encode_path", path)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(INNER):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / INNER)
    ts.sort()
    gbs = B * 307520 / (ts[len(ts) // 2] * 1e-6) / 1e9
    print(f"{os.path.basename(os.environ.get('SSDX_LIB', 'libssdx.so')):24s} encode {path:11s} B={B}: median "
          f"{ts[len(ts)//2]:7.1f} us  best {ts[0]:7.1f} us  ({gbs:6.0f} GB/s algorithmic)", flush=True)
sx.set_option("encode_path", "auto")
if hasattr(sx.lib, "ssdx_debug_phase_cycles_encode_grid"):
    out = (ctypes.c_ulonglong * 32)()
    sx.lib.ssdx_debug_phase_cycles_encode_grid.argtypes = [ctypes.c_void_p]
    sx.set_option("encode_path", "grid_fused")
    sx.lib.ssdx_debug_phase_cycles_encode_grid(out)
    fn(); torch.cuda.synchronize()
    sx.lib.ssdx_debug_phase_cycles_encode_grid(out)
    v = list(out)
    names = [(0, "setup (per CTA)"), (6, "stage GT"), (1, "pass 0"), (2, "pass 1"), (3, "closed form"),
             (4, "wait background"), (5, "emit + reset")]
    nimg = max(v[24], 1); tot = sum(v[i] for i, _ in names)
    print(f"grid fused: images {v[24]}  cycles/image {tot/nimg:.0f}")
    for i, n in names:
        print(f"  {n:16s} {v[i]/nimg:9.0f} cycles/image  {100*v[i]/max(tot,1):5.1f}%")
