"""GPU box: device-timed segmentation-presence reduction (layers.SegmentationSuppression) and the
fused vs materialised suppression in front of the NMS.  CUDA events, C ABI, inputs resident."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np, torch
from ssdseglib import _ssdx as sx
PEAK = 6542.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
sp = torch.cuda.current_stream().cuda_stream
def timeit(fn, inner=10, reps=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort(); return ts[len(ts) // 2]
B, H, W, C = 64, 480, 640, 4
g = torch.Generator(device="cuda"); g.manual_seed(1)
mask = torch.rand((B, H, W, C), generator=g, device="cuda")
pres = torch.empty(4, device="cuda"); ws = torch.empty(256, dtype=torch.uint8, device="cuda")
def run(m):
    sx.check(sx.lib.ssdx_seg_class_presence(m.data_ptr(), B * H * W, C, 4, pres.data_ptr(), ws.data_ptr(), 256, 0, sp))
us = timeit(lambda: run(mask)); print(json.dumps({"case": "presence, every class appears (early exit once all 4 are seen)", "images": B, "us": round(us, 1), "presence": pres.tolist()}))
absent = mask.clone(); absent[..., 2] = -1.0   # class 2 never wins: the whole tensor has to be read
us = timeit(lambda: run(absent)); nbytes = B * H * W * C * 4
print(json.dumps({"case": "presence, one class absent (full pass)", "images": B, "us": round(us, 1), "presence": pres.tolist(),
                  "GBps": nbytes / us / 1e3, "frac_of_measured_hbm_peak": nbytes / us / 1e3 / PEAK}))
scores = torch.rand((1024, 9600, 4), generator=g, device="cuda"); out = torch.empty_like(scores)
us = timeit(lambda: sx.check(sx.lib.ssdx_scale_scores(scores.data_ptr(), 1024 * 9600, 4, pres.data_ptr(), out.data_ptr(), 0, sp)))
print(json.dumps({"case": "materialised product (B,A,C) * presence, B=1024", "us": round(us, 1), "GBps": 2 * scores.numel() * 4 / us / 1e3,
                  "note": "the fused path (ssdx_nms_scaled) avoids this pass: 4 extra multiplies per float4 inside the NMS select"}))

# ---- metrics.jaccard_iou_bounding_boxes at batch 1024 (two (B,A,4) streams in, (B,) out)
import ssdseglib
from oracle import ssd_oracle as orc
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
coder = ssdseglib.datacoder.DataEncoderDecoder(4, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
labels, boxes, counts = orc.synth_gt(np.random.default_rng(3), 1024, 16)
_, yt = coder.encode_batch(torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda())
yp = yt + 0.5 * torch.randn(yt.shape, generator=g, device="cuda")
outm = torch.empty(1024, device="cuda")
stdm = sx.std_array((0.1, 0.1, 0.2, 0.2))
us = timeit(lambda: sx.check(sx.lib.ssdx_metric_iou_boxes(coder._anchors.handle, yt.data_ptr(), yp.data_ptr(), 1024, stdm, outm.data_ptr(), sp)))
nb = 2 * yt.numel() * 4
print(json.dumps({"case": "metrics.jaccard_iou_bounding_boxes, B=1024", "us": round(us, 1), "GBps": nb / us / 1e3, "frac_of_measured_hbm_peak": nb / us / 1e3 / PEAK, "mean_iou": float(outm.nanmean())}))
