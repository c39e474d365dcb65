import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
B, A, C, K = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 9600, 4, 200
off, sc = orc.synth_predictions(np.random.default_rng(1994), B, A, C)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1)
dec = ssdseglib.layers.DecodeBoxesCentroidsOffsets(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], 0.1, 0.1, 0.2, 0.2)
yx = dec(torch.from_numpy(off).cuda()); scd = torch.from_numpy(sc).cuda()
out = torch.empty((B, K, 6), device="cuda"); val = torch.empty(B, dtype=torch.int32, device="cuda")
ws = torch.empty(int(sx.lib.ssdx_nms_workspace_bytes(B, A, C, K, K)), dtype=torch.uint8, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
fn = lambda: sx.check(sx.lib.ssdx_nms(yx.data_ptr(), scd.data_ptr(), B, A, C, K, K, 0.45, 0.01, out.data_ptr(), val.data_ptr(), None, ws.data_ptr(), ws.numel(), 0, sp))
for path in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("auto", "blocked", "perclass", "pcta")):
    sx.set_option("nms_path", path)
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 100)
    ts.sort(); print(f"B={B} nms_path={path}: {ts[len(ts)//2]:.1f} us")
if len(sys.argv) > 3:  # the two phases on their own (ssdx_nms_select / ssdx_nms_finish)
    sx.set_option("nms_path", "auto")
    f1 = lambda: sx.check(sx.lib.ssdx_nms_select(scd.data_ptr(), None, B, A, C, K, K, 0.01, ws.data_ptr(), ws.numel(), 0, sp))
    f2 = lambda: sx.check(sx.lib.ssdx_nms_finish(yx.data_ptr(), scd.data_ptr(), None, B, A, C, K, K, 0.45, 0.01, out.data_ptr(), val.data_ptr(), None, ws.data_ptr(), ws.numel(), 0, sp))
    for name, f in (("select", f1), ("finish", f2)):
        for _ in range(5): f()
        torch.cuda.synchronize(); ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): f()
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 100)
        ts.sort(); print(f"B={B} {name} alone: {ts[len(ts)//2]:.1f} us")
