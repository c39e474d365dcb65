"""GPU box helper for profiling: a few calls of ssdx_nms_logits / ssdx_nms (and, with TAIL=1, of the one-call inference
tail ssdx_decode_nms_logits) on the bench's round-trip predictions (B = 1024).  NMS setting from the environment:
MPC, MTOT, IOU, THR (default 200, 200, 0.45, 0.01)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
B, A, C = 1024, 9600, 4
MPC, K = int(os.environ.get("MPC", 200)), int(os.environ.get("MTOT", 200))
IOU, THR = float(os.environ.get("IOU", 0.45)), float(os.environ.get("THR", 0.01))
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
coder = ssdseglib.datacoder.DataEncoderDecoder(C, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
labels, boxes, counts = orc.synth_gt(np.random.default_rng(1995), B, 16)
lab, off = coder.encode_batch(torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda())
g = torch.Generator(device="cuda"); g.manual_seed(7)
sc = (0.9 * lab + 0.1 * torch.softmax(torch.randn(lab.shape, generator=g, device="cuda"), dim=-1)).contiguous()
lg = torch.log(sc).contiguous()
bx = torch.empty((B, A, 4), device="cuda"); sp = torch.cuda.current_stream().cuda_stream
sx.check(sx.lib.ssdx_decode(coder._anchors.handle, off.data_ptr(), B, sx.std_array((0.1, 0.1, 0.2, 0.2)), 0, bx.data_ptr(), sp))
out = torch.empty((B, K, 6), device="cuda")
ws = torch.empty(int(sx.lib.ssdx_nms_logits_workspace_bytes(B, A, C, MPC, K)), dtype=torch.uint8, device="cuda")
stdv = sx.std_array((0.1, 0.1, 0.2, 0.2))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    if os.environ.get("TAIL"):
        sx.check(sx.lib.ssdx_decode_nms_logits(coder._anchors.handle, off.data_ptr(), stdv, lg.data_ptr(), None, B, C, MPC, K, IOU, THR, out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), sp))
        continue
    sx.check(sx.lib.ssdx_nms_logits(bx.data_ptr(), lg.data_ptr(), None, B, A, C, MPC, K, IOU, THR, out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 0, sp))
    sx.check(sx.lib.ssdx_nms(bx.data_ptr(), sc.data_ptr(), B, A, C, MPC, K, IOU, THR, out.data_ptr(), None, None, ws.data_ptr(), ws.numel(), 0, sp))
torch.cuda.synchronize()
