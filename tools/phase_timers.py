"""Developer aid (GPU box): per-phase cycle breakdown of ssdx_encode_fused_kernel and ssdx_nms_fast_kernel using the
debug build (lib/libssdx_dbg.so, compiled by `python build.py --debug-timers` with -DSSDX_PHASE_TIMERS)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")
os.environ["SSDX_LIB"] = os.path.join(PKG, "lib", "libssdx_dbg.so")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
sys.argv = ["bench.py"]
import bench
B = 1024
anchors, labels, boxes, counts = bench.make_inputs(B, 1995)
coder = ssdseglib.datacoder.DataEncoderDecoder(4, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
out = (ctypes.c_ulonglong * 32)()
sx.lib.ssdx_debug_phase_cycles.argtypes = [ctypes.c_void_p]
sx.lib.ssdx_debug_phase_cycles_encode.argtypes = [ctypes.c_void_p]
gb, gl, gc = torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda()
for it in range(2):
    sx.lib.ssdx_debug_phase_cycles_encode(out)
    onehot, off = coder.encode_batch(gb, gl, gc); torch.cuda.synchronize()
sx.lib.ssdx_debug_phase_cycles_encode(out)
v = list(out)
names = ["init", "seed", "match (warp 0)", "wait warps+bulk", "closed form", "compact+patch"]
ncta = max(v[24], 1); tot = sum(v[16:22])
print(f"encode fused: CTAs {v[24]}  cycles/CTA {tot/ncta:.0f}")
for n, c in zip(names, v[16:22]):
    print(f"  {n:16s} {c/ncta:9.0f} cycles/CTA  {100*c/max(tot,1):5.1f}%")
print(f"  per image: surviving (tile,GT) {v[25]/ncta:.0f}  slow-path entries {v[26]/ncta:.0f}  divisions {v[27]/ncta:.0f}  GT-side updates {v[28]/ncta:.0f}")
scores = bench.make_scores_torch(onehot, 4242)
yxyx = coder._decode(off, 0)
nms = ssdseglib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
for it in range(2):
    sx.lib.ssdx_debug_phase_cycles(out)
    nms(yxyx, scores); torch.cuda.synchronize()
sx.lib.ssdx_debug_phase_cycles(out)
v = list(out)
names = ["list load", "sort", "box gather", "pair tests", "block resolve", "refill scans", "tail"]
ncta = max(v[8], 1)
tot = sum(v[:7])
print(f"nms fast: CTAs {v[8]}  refill scans {v[9]}  candidates consumed/CTA {v[10]/ncta:.0f}  cycles/CTA {tot/ncta:.0f}")
for n, c in zip(names, v[:7]):
    print(f"  {n:14s} {c/ncta:9.0f} cycles/CTA  {100*c/max(tot,1):5.1f}%")
