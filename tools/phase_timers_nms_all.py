"""Developer aid (GPU box): per-phase cycle breakdown of ssdx_nms_fast_all_kernel (debug build, build.py --debug-timers)
on random predictions (BASELINE config 1 data).  usage: python tools/phase_timers_nms_all.py [B]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")
os.environ["SSDX_LIB"] = os.path.join(PKG, "lib", "libssdx_dbg.so")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
off, sc = orc.synth_predictions(np.random.default_rng(1994), B, 9600, 4)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1)
dec = ssdseglib.layers.DecodeBoxesCentroidsOffsets(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], 0.1, 0.1, 0.2, 0.2)
yx = dec(torch.from_numpy(off).cuda()); scd = torch.from_numpy(sc).cuda()
nms = ssdseglib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
out = (ctypes.c_ulonglong * 32)()
sx.lib.ssdx_debug_phase_cycles.argtypes = [ctypes.c_void_p]
for it in range(2):
    sx.lib.ssdx_debug_phase_cycles(out)
    nms(yx, scd); torch.cuda.synchronize()
sx.lib.ssdx_debug_phase_cycles(out)
v = list(out)
names = {10: "list load", 11: "sort", 12: "box gather", 21: "kept-box tests", 22: "tiles (warp 0)", 13: "pair tests: barrier",
         14: "block resolve", 15: "refill scans", 24: "kept boxes copied", 23: "loop exit: code", 19: "loop exit: barrier", 20: "rows staged", 16: "rows stored"}
n = max(v[18], 1); tot = sum(v[k] for k in names)
print(f"nms fast_all, B={B}: images {v[18]}  cycles/image {tot/n:.0f} ({tot/n/1965:.1f} us)")
for k, nm in names.items():
    print(f"  {nm:20s} {v[k]/n:9.0f} cycles/image  {100*v[k]/max(tot,1):5.1f}%")
