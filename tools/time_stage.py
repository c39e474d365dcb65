"""Developer aid (GPU box): time one stage of the round-trip workload with CUDA events for a given library build.
usage: SSDX_LIB=<path to libssdx*.so> python tools/time_stage.py [encode|decode|nms] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
stage = sys.argv[1] if len(sys.argv) > 1 else "encode"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
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
nms = ssdseglib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
dev = 0
f32 = dict(dtype=torch.float32, device="cuda:0")
o_labels, o_offsets = torch.empty((B, 9600, 4), **f32), torch.empty((B, 9600, 4), **f32)
o_boxes, o_det = torch.empty((B, 9600, 4), **f32), torch.empty((B, 200, 6), **f32)
o_valid = torch.empty((B,), dtype=torch.int32, device="cuda:0")
ws_enc = torch.empty(max(sx.lib.ssdx_encode_workspace_bytes(9600, B, 16), 256), dtype=torch.uint8, device="cuda:0")
ws_nms = torch.empty(max(sx.lib.ssdx_nms_workspace_bytes(B, 9600, 4, 200, 200), 256), dtype=torch.uint8, device="cuda:0")
std = sx.std_array((0.1, 0.1, 0.2, 0.2))
sp = torch.cuda.current_stream().cuda_stream
h = coder._anchors.handle
fn = {
    "encode": lambda: sx.check(sx.lib.ssdx_encode(h, gb.data_ptr(), gl.data_ptr(), gc.data_ptr(), B, 16, 4, 0.5, std,
                                                  o_labels.data_ptr(), o_offsets.data_ptr(), None, ws_enc.data_ptr(),
                                                  ws_enc.numel(), sp)),
    "decode": lambda: sx.check(sx.lib.ssdx_decode(h, off.data_ptr(), B, std, 0, o_boxes.data_ptr(), sp)),
    "nms": lambda: sx.check(sx.lib.ssdx_nms(yxyx.data_ptr(), scores.data_ptr(), B, 9600, 4, 200, 200, 0.45, 0.01,
                                            o_det.data_ptr(), o_valid.data_ptr(), None, ws_nms.data_ptr(),
                                            ws_nms.numel(), dev, sp)),
}[stage]
INNER = 10
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
print(f"{os.path.basename(os.environ.get('SSDX_LIB', 'libssdx.so')):28s} {stage}: median {ts[len(ts)//2]:7.1f} us  best {ts[0]:7.1f} us  per launch (B={B}, C ABI, {INNER} back-to-back launches per sample)")
