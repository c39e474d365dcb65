"""Generate tests/golden/*.npz by executing the REFERENCE's own source files.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

  * ssdseglib/boxes.py      - executed unchanged (pure NumPy).
  * ssdseglib/evaluators.py - executed unchanged with `tensorflow` stubbed
                              (the detection functions never call it).
  * ssdseglib/datacoder.py, ssdseglib/layers.py - executed unchanged over
    tests/golden/tf_numpy_shim.py (NumPy emulation of the TF primitives).

Nothing here is read at test time except the .npz files it writes.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = os.environ.get("SSDX_REFERENCE", "/root/reference")

import tf_numpy_shim  # noqa: E402
from oracle import ssd_oracle as orc  # noqa: E402  (generators only)

warnings.simplefilter("ignore", DeprecationWarning)


def load_ref(name):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, "ssdseglib", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def write_csv(path, labels, boxes, newline="\r\n"):
    with open(path, "wb") as f:
        rows = [f"{int(l)},{repr(float(b[0]))},{repr(float(b[1]))},{repr(float(b[2]))},{repr(float(b[3]))}"
                for l, b in zip(labels, boxes)]
        f.write((newline.join(rows) + newline).encode() if rows else b"")  # csv.reader rejects a blank row


def main():
    tf_numpy_shim.install()
    ref_boxes = load_ref("boxes")
    ref_datacoder = load_ref("datacoder")
    ref_layers = load_ref("layers")
    ref_eval = load_ref("evaluators")

    # ---------------- anchors (boxes.py executed unchanged) ----------------
    std = ref_boxes.DefaultBoundingBoxes(
        feature_maps_shapes=((30, 40), (15, 20), (8, 10), (4, 5)),
        centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1),
        boxes_scales=(0.15, 0.95), additional_square_box=True)
    std.rescale_boxes_coordinates(image_shape=(480, 640))
    std_corners = std.get_boxes_coordinates_corners("ssd")
    std_centroids = std.get_boxes_coordinates_centroids("ssd")

    tut = ref_boxes.DefaultBoundingBoxes(  # notebook 01 cell 6 (default scales)
        feature_maps_shapes=((30, 40), (15, 20), (8, 10), (4, 5)),
        centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1), additional_square_box=True)
    tut.rescale_boxes_coordinates(image_shape=(480, 640))

    odd = ref_boxes.DefaultBoundingBoxes(  # per-map ratios, a 1-wide map, no extra square, unscaled
        feature_maps_shapes=((3, 1), (2, 2)), feature_maps_aspect_ratios=((1, 2), (1, 0.5, 3)),
        boxes_scales=(0.3, 0.8), centers_padding_from_borders_percentage=0.1, additional_square_box=False)
    odd.rescale_boxes_coordinates(image_shape=(33, 65))
    odd_once = odd.get_boxes_coordinates_corners("ssd").copy()
    odd.rescale_boxes_coordinates(image_shape=(33, 65))  # Q1: in-place, scales twice
    odd_twice = odd.get_boxes_coordinates_corners("ssd").copy()

    stress = ref_boxes.DefaultBoundingBoxes(feature_maps_shapes=((125, 160),),
                                            feature_maps_aspect_ratios=(1, 2, 3, 1 / 2, 1 / 3),
                                            additional_square_box=False)
    stress.rescale_boxes_coordinates(image_shape=(1000, 1280))
    stress_corners = stress.get_boxes_coordinates_corners("ssd")
    sel = np.arange(0, stress_corners.shape[0], 997)

    cx, cy, w, h = ref_boxes.coordinates_corners_to_centroids(std_corners[:, 0], std_corners[:, 1],
                                                             std_corners[:, 2], std_corners[:, 3])
    bx0, by0, bx1, by1 = ref_boxes.coordinates_centroids_to_corners(cx, cy, w, h)
    np.savez_compressed(
        os.path.join(HERE, "anchors.npz"),
        std_corners=std_corners, std_centroids=std_centroids, std_scales=std.boxes_scales,
        std_xmin=std.get_boxes_coordinates_xmin("ssd"), std_height=std.get_boxes_coordinates_height("ssd"),
        std_fm1_corners=std.get_boxes_coordinates_corners("feature-maps")[1],
        std_fm3_width=std.get_boxes_coordinates_width("feature-maps")[3],
        tut_corners=tut.get_boxes_coordinates_corners("ssd"),
        odd_once=odd_once, odd_twice=odd_twice,
        stress_rows=sel, stress_sel=stress_corners[sel], stress_shape=np.array(stress_corners.shape),
        stress_sum=np.array([np.sum(stress_corners, dtype=np.float64)]),
        conv_centroids=np.stack([cx, cy, w, h], 1), conv_corners=np.stack([bx0, by0, bx1, by1], 1))

    # ---------------- encode / decode (datacoder.py over the TF shim) ----------------
    tmp = tempfile.mkdtemp(prefix="ssdx_golden_")
    cases = {}
    rng = np.random.default_rng(1993)
    lab, box, _ = orc.synth_gt(rng, 8, 16)                       # BASELINE config 0
    cases["cfg0"] = (lab, box, 0.5, False)
    cases["cfg0_thr525_flip"] = (lab, box, 0.525, True)          # notebook threshold + h-flip
    lab_i, box_i, _ = orc.synth_gt(np.random.default_rng(7), 6, 12, integer=True)  # integer coords: many exact ties
    box_i[:, 3] = box_i[:, 1]; lab_i[:, 3] = 2                    # duplicated GT box with another label
    cases["ties_int"] = (lab_i, box_i, 0.5, False)
    lab_s, box_s, _ = orc.synth_gt(np.random.default_rng(11), 4, 5, wh_range=(4.0, 30.0))  # small boxes: step-1-only
    cases["small"] = (lab_s, box_s, 0.5, False)
    lab_o = lab.copy()[:2]; lab_o[:, 0] = 7                       # out-of-range label -> all-zero one-hot row
    cases["label_oob"] = (lab_o, box[:2], 0.3, False)

    enc_out = {}
    for name, (labels, boxes, thr, flip) in cases.items():
        coder = ref_datacoder.DataEncoderDecoder(
            num_classes=4, image_shape=(480, 640),
            xmin_boxes_default=std.get_boxes_coordinates_xmin("ssd"),
            ymin_boxes_default=std.get_boxes_coordinates_ymin("ssd"),
            xmax_boxes_default=std.get_boxes_coordinates_xmax("ssd"),
            ymax_boxes_default=std.get_boxes_coordinates_ymax("ssd"),
            iou_threshold=thr, standard_deviations_centroids_offsets=(0.1, 0.1, 0.2, 0.2),
            augmentation_horizontal_flip=flip)
        L, O, DC, DK = [], [], [], []
        for i in range(labels.shape[0]):
            p = os.path.join(tmp, f"{name}_{i}.csv")
            write_csv(p, labels[i], boxes[i])
            l, o = coder._encode_ground_truth_labels_boxes(p, flip)
            L.append(l); O.append(o)
            DC.append(coder.decode_to_centroids(o)); DK.append(coder.decode_to_corners(o))
        enc_out[name + "_gt_labels"] = labels
        enc_out[name + "_gt_boxes"] = boxes
        enc_out[name + "_thr"] = np.array([thr]); enc_out[name + "_flip"] = np.array([flip])
        enc_out[name + "_labels"] = np.stack(L).astype(np.float32)
        enc_out[name + "_offsets"] = np.stack(O).astype(np.float32)
        enc_out[name + "_dec_centroids"] = np.stack(DC).astype(np.float32)
        enc_out[name + "_dec_corners"] = np.stack(DK).astype(np.float32)
    # ctor given centroids instead of corners (datacoder.py:73-89)
    coder_c = ref_datacoder.DataEncoderDecoder(
        num_classes=4, image_shape=(480, 640),
        center_x_boxes_default=std.get_boxes_coordinates_center_x("ssd"),
        center_y_boxes_default=std.get_boxes_coordinates_center_y("ssd"),
        width_boxes_default=std.get_boxes_coordinates_width("ssd"),
        height_boxes_default=std.get_boxes_coordinates_height("ssd"))
    p = os.path.join(tmp, "cfg0_0.csv")
    l, o = coder_c._encode_ground_truth_labels_boxes(p, False)
    enc_out["centroid_ctor_labels"] = l.astype(np.float32); enc_out["centroid_ctor_offsets"] = o.astype(np.float32)
    enc_out["centroid_ctor_corners"] = np.stack([coder_c.xmin_boxes_default, coder_c.ymin_boxes_default,
                                                 coder_c.xmax_boxes_default, coder_c.ymax_boxes_default], 1)
    np.savez_compressed(os.path.join(HERE, "encode_decode.npz"), **enc_out)

    # ---------------- layers.py decode + NMS wrapper (over the TF shim) ----------------
    rng = np.random.default_rng(1994)
    off, scores = orc.synth_predictions(rng, 2, 9600, 4)
    dec = ref_layers.DecodeBoxesCentroidsOffsets(
        std.get_boxes_coordinates_center_x("ssd"), std.get_boxes_coordinates_center_y("ssd"),
        std.get_boxes_coordinates_width("ssd"), std.get_boxes_coordinates_height("ssd"), 0.1, 0.1, 0.2, 0.2)
    yxyx = dec.call(off)
    nms = ref_layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
    det = nms.call(yxyx, scores)
    nms_nb = ref_layers.NonMaximumSuppression(4, 10, 0.025, 0.2, True)  # notebook-03-like + background drop
    det_nb = nms_nb.call(yxyx, scores)
    np.savez_compressed(os.path.join(HERE, "layers.npz"), offsets=off, scores=scores, decoded_yxyx=yxyx,
                        det_200=det.astype(np.float32), det_bgdrop=det_nb.astype(np.float32))

    # ---------------- evaluators.py (executed unchanged) ----------------
    rng = np.random.default_rng(1997)
    N, K, G = 12, 24, 6
    gl, gb, _ = orc.synth_gt(rng, N, G, integer=True)
    paths = []
    for i in range(N):
        p = os.path.join(tmp, f"eval_{i}.csv")
        n = G if i != 3 else 0                     # image 3 has no ground truth at all
        write_csv(p, gl[i, :n], gb[i, :n], newline="\n")
        paths.append(p)
    pl = rng.integers(0, 4, size=(N, K)).astype(np.int32)
    pc = rng.uniform(0.05, 1.0, size=(N, K)).astype(np.float32)
    pb = np.zeros((N, K, 4), np.float32)
    for i in range(N):
        for k in range(K):
            g = rng.integers(0, G)
            jit = rng.normal(0, 12.0, size=4)
            pb[i, k] = gb[i, g] + jit
            if rng.uniform() < 0.5:
                pl[i, k] = gl[i, g]
    pl[5, :] = 0                                    # all-background image
    ious = [ref_eval._iou_boxes_pred_vs_true(pl[i], pb[i], gl[i], gb[i]) for i in range(N)]
    iou_nogt = ref_eval._iou_boxes_pred_vs_true(pl[0], pb[0], np.zeros(0, np.int32), np.zeros((0, 4), np.float32))
    ap = {}
    for thr in (0.5, 0.6, 0.7):
        r = ref_eval.average_precision_object_detection(pl, pc, pb, thr, paths, [0, 1, 2, 3], 0)
        ap[thr] = np.array([r[1], r[2], r[3]], dtype=np.float64)
    # survey KAT
    kat_iou = ref_eval._iou_boxes_pred_vs_true(
        np.array([1, 2, 0, 1], np.int32),
        np.array([[12, 11, 98, 99], [190, 60, 310, 190], [0, 0, 0, 0], [300, 300, 320, 320]], np.float32),
        np.array([1, 2], np.int32), np.array([[10, 10, 100, 100], [200, 50, 300, 200]], np.float32))
    np.savez_compressed(os.path.join(HERE, "eval.npz"), gt_labels=gl, gt_boxes=gb, no_gt_image=np.array([3]),
                        pred_labels=pl, pred_conf=pc, pred_boxes=pb, iou=np.stack(ious), iou_nogt=iou_nogt,
                        ap50=ap[0.5], ap60=ap[0.6], ap70=ap[0.7], kat_iou=kat_iou)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(HERE, f))} bytes")


if __name__ == "__main__":
    main()
