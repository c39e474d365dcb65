"""Developer aid (GPU box): does running the two NMS kernels of two half-batches on two streams pay?
select(h1); [finish(h1) on a side stream || select(h2)]; finish(h2).  usage: python tools/proto_nms_halves.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
B, A, C, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 9600, 4, 200
off, sc = orc.synth_predictions(np.random.default_rng(1994), B, A, C)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1)
dec = ssdseglib.layers.DecodeBoxesCentroidsOffsets(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], 0.1, 0.1, 0.2, 0.2)
yx = dec(torch.from_numpy(off).cuda()); scd = torch.from_numpy(sc).cuda()
out = torch.empty((B, K, 6), device="cuda"); val = torch.empty(B, dtype=torch.int32, device="cuda")
ref = torch.empty_like(out)
h = B // 2
wsb = int(sx.lib.ssdx_nms_workspace_bytes(B, A, C, K, K))
ws = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(2)]
main, side = torch.cuda.Stream(), torch.cuda.Stream()
e1, e2 = torch.cuda.Event(), torch.cuda.Event()

def sel(lo, n, w, st):
    sx.check(sx.lib.ssdx_nms_select(scd[lo:].data_ptr(), None, n, A, C, K, K, 0.01, w.data_ptr(), w.numel(), 0, st.cuda_stream))
def fin(lo, n, w, st, o):
    sx.check(sx.lib.ssdx_nms_finish(yx[lo:].data_ptr(), scd[lo:].data_ptr(), None, n, A, C, K, K, 0.45, 0.01, o[lo:].data_ptr(),
                                    val[lo:].data_ptr(), None, w.data_ptr(), w.numel(), 0, st.cuda_stream))
def whole(o):
    sx.check(sx.lib.ssdx_nms(yx.data_ptr(), scd.data_ptr(), B, A, C, K, K, 0.45, 0.01, o.data_ptr(), val.data_ptr(), None,
                             ws[0].data_ptr(), ws[0].numel(), 0, main.cuda_stream))
def halves(o):
    sel(0, h, ws[0], main); e1.record(main)
    side.wait_event(e1); fin(0, h, ws[0], side, o); e2.record(side)
    sel(h, B - h, ws[1], main); fin(h, B - h, ws[1], main, o)
    main.wait_event(e2)
def serial_halves(o):
    sel(0, h, ws[0], main); fin(0, h, ws[0], main, o); sel(h, B - h, ws[1], main); fin(h, B - h, ws[1], main, o)

with torch.cuda.stream(main):
    whole(ref); halves(out); torch.cuda.synchronize()
    print("halves == whole:", bool(torch.equal(ref, out)))
    for name, f in (("one call", whole), ("two halves, two streams", halves), ("two halves, one stream", serial_halves)):
        for _ in range(5): f(out)
        torch.cuda.synchronize(); ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main)
            for _ in range(10): f(out)
            b.record(main); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 100)
        ts.sort(); print(f"B={B} {name}: {ts[len(ts)//2]:.1f} us")
