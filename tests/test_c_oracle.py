"""The C/OpenMP port (oracle/ssd_oracle.c, the timed CPU baseline) must agree with the NumPy oracle."""
import numpy as np

from oracle import c_oracle
from oracle import ssd_oracle as orc

STD = (0.1, 0.1, 0.2, 0.2)


def test_c_encode_decode_nms_eval_match_numpy_oracle(std_anchors, golden):
    rng = np.random.default_rng(21)
    labels, boxes, counts = orc.synth_gt(rng, 6, 16, integer=True)
    boxes[:, 4] = boxes[:, 1]
    counts = np.array([16, 0, 5, 16, 1, 9], np.int32)
    L, O, M = c_oracle.encode_batch(std_anchors, labels, boxes, counts, 4, 0.5, STD, nthreads=4)
    Lr, Or, Mr = orc.encode_batch(std_anchors, labels, boxes, counts, 4, 0.5, STD, literal=True)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    np.testing.assert_allclose(O, Or, rtol=1e-5, atol=1e-6)

    cent = golden("anchors")["std_centroids"]
    g = golden("layers")
    for mode, fn in ((0, orc.decode_batched_yxyx), (1, orc.decode_to_centroids), (2, orc.decode_to_corners)):
        np.testing.assert_allclose(c_oracle.decode_batch(cent, g["offsets"], STD, mode, nthreads=2),
                                   fn(g["offsets"], cent, STD), rtol=1e-5, atol=1e-3)
    out, valid, aidx = c_oracle.combined_nms(g["decoded_yxyx"], g["scores"], 200, 200, 0.45, 0.01, nthreads=2)
    assert np.array_equal(out, g["det_200"])
    ro, rv, ra = orc.combined_nms(g["decoded_yxyx"], g["scores"], 7, 12, 0.1, 0.3)
    co, cv, ca = c_oracle.combined_nms(g["decoded_yxyx"], g["scores"], 7, 12, 0.1, 0.3)
    assert np.array_equal(ro, co) and np.array_equal(rv, cv) and np.array_equal(ra, ca)

    e = golden("eval")
    n = e["pred_labels"].shape[0]
    gc = np.full(n, e["gt_labels"].shape[1], np.int32)
    tp, cnt = c_oracle.eval_tp(e["pred_labels"], e["pred_boxes"], e["gt_boxes"], e["gt_labels"], gc, 4, 0, 0.5)
    ref = np.stack([orc.eval_tp_flags(e["pred_labels"][i], e["pred_boxes"][i], e["gt_labels"][i], e["gt_boxes"][i], 0.5)
                    for i in range(n)])
    assert np.array_equal(tp, ref)
    assert cnt[:, 0].sum() == ref.sum() and cnt[:, 2].sum() == gc.sum()
