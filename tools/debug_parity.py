"""Developer aid (GPU box): print where the CUDA encode/NMS differ from the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np
import ssdseglib
from oracle import ssd_oracle as orc

STD = (0.1, 0.1, 0.2, 0.2)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
rng = np.random.default_rng(1993)
labels, boxes, counts = orc.synth_gt(rng, 8, 16)
coder = ssdseglib.datacoder.DataEncoderDecoder(4, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
Lr, Or, Mr = orc.encode_batch(anchors, labels, boxes, counts, 4, 0.5, STD)
bad = np.argwhere(M != Mr)
print("encode: match mismatches", len(bad), "of", M.size, "positives ref", (Mr >= 0).sum(), "gpu", (M >= 0).sum())
for b, a in bad[:10]:
    iou = orc.iou_matrix(anchors, boxes[b])
    print(f"  img {b} anchor {a}: gpu {M[b,a]} ref {Mr[b,a]} row-max {iou[a].max():.6f} argmax {iou[a].argmax()} "
          f"col-argmax-hits {[g for g in range(16) if iou[:, g].argmax() == a]}")
print("labels equal:", np.array_equal(L, Lr), " offsets max abs err:", np.abs(O - Or).max())
cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1)
off, scores = orc.synth_predictions(np.random.default_rng(1994), 2, 9600, 4)
yxyx = orc.decode_batched_yxyx(off, cent, STD)
det, valid, aidx = ssdseglib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False).call(yxyx, scores, return_aux=True)
dr, vr, ar = orc.combined_nms(yxyx, scores, 200, 200, 0.45, 0.01)
print("nms: valid", valid, vr, "idx equal", np.array_equal(aidx, ar), "first diff", np.argwhere(aidx != ar)[:5].tolist())
