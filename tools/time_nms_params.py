"""GPU box: NMS stage of the round-trip workload at batch 1024 for several (max_per_class, max_total, iou, score) settings,
including the reference notebook's own (4, 10, 0.025, 0.725; 03-...ipynb cell 23)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
sys.argv = ["bench.py"]
import bench
B = 1024
anchors, labels, boxes, counts = bench.make_inputs(B, 1995)
coder = ssdseglib.datacoder.DataEncoderDecoder(4, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
gb, gl, gc = torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda()
onehot, off = coder.encode_batch(gb, gl, gc)
scores = bench.make_scores_torch(onehot, 4242)
yxyx = coder._decode(off, 0)
logits = torch.log(scores.clamp_min(1e-30)).contiguous()
sp = torch.cuda.current_stream().cuda_stream
SETTINGS = ((200, 200, 0.45, 0.01), (4, 10, 0.025, 0.725), (100, 200, 0.45, 0.01), (20, 50, 0.3, 0.3),
            (50, 20, 0.3, 0.3), (100, 100, 0.45, 0.01))
only = os.environ.get("NMS_SETTING")  # index into SETTINGS (profiling one setting under ncu)
for mpc, mtot, iou, thr in (SETTINGS if only is None else (SETTINGS[int(only)],)):
    out = torch.empty((B, mtot, 6), device="cuda")
    ws = torch.empty(max(sx.lib.ssdx_nms_workspace_bytes(B, 9600, 4, mpc, mtot), 256), dtype=torch.uint8, device="cuda")
    fn = lambda: sx.check(sx.lib.ssdx_nms(yxyx.data_ptr(), scores.data_ptr(), B, 9600, 4, mpc, mtot, iou, thr, out.data_ptr(),
                                          None, None, ws.data_ptr(), ws.numel(), 0, sp))
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 100)
    ts.sort()
    # the same call on raw logits (soft-max inside the kernels that read the scores: ssdx_nms_logits)
    wl = torch.empty(max(sx.lib.ssdx_nms_logits_workspace_bytes(B, 9600, 4, mpc, mtot), 256), dtype=torch.uint8, device="cuda")
    fl = lambda: sx.check(sx.lib.ssdx_nms_logits(yxyx.data_ptr(), logits.data_ptr(), None, B, 9600, 4, mpc, mtot, iou, thr,
                                                 out.data_ptr(), None, None, wl.data_ptr(), wl.numel(), 0, sp))
    for _ in range(3): fl()
    torch.cuda.synchronize(); tl = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fl()
        e1.record(); torch.cuda.synchronize(); tl.append(e0.elapsed_time(e1) * 100)
    tl.sort()
    print(f"max_per_class={mpc:4d} max_total={mtot:4d} iou={iou} score_thr={thr}: {ts[len(ts)//2]:7.1f} us per 1024 images "
          f"({'merged' if mpc >= mtot else 'per-class'} pipeline); on logits {tl[len(tl)//2]:7.1f} us")
