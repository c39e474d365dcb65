"""GPU box: randomized differential test of the encode drivers (one CTA per image / anchor-tiled) against the C oracle:
random anchor subsets and synthetic grids, GT counts (incl. 0 and ragged), integer coordinates (exact IoU ties),
duplicated GT boxes, GT boxes equal to anchors, thresholds, class counts, out-of-range labels.
usage: python tools/fuzz_encode.py [cases] [seed] [encode_path]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np
import ssdseglib
from oracle import c_oracle
from oracle import ssd_oracle as orc

STD = (0.1, 0.1, 0.2, 0.2)


def run(cases: int, seed: int, verbose: bool = True) -> int:
    c_oracle.build()
    rng = np.random.default_rng(seed)
    std_anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
    bad = 0
    n_grid = 0
    for it in range(cases):
        kind = rng.choice(["subset", "grid", "random", "dbb", "dbb"])
        image_hw = (480, 640)
        if kind == "dbb":  # a random DefaultBoundingBoxes table (boxes.py:6-151): what the grid drivers are for
            n_maps = int(rng.integers(1, 5))
            shapes, h, w = [], int(rng.integers(6, 40)), int(rng.integers(6, 48))
            for _ in range(n_maps):
                shapes.append((h, w))
                h, w = max(6, (h + 1) // 2), max(6, (w + 1) // 2)
            if rng.integers(0, 4) == 0:
                shapes[-1] = (1, 1)  # boxes smaller than a cell get negative extents (boxes.py:143): not a grid table
            ratios = tuple(float(r) for r in rng.choice([1, 2, 3, 0.5, 1 / 3, 1.5], size=int(rng.integers(1, 6)),
                                                         replace=False))
            image_hw = (int(rng.integers(64, 1200)), int(rng.integers(64, 1400)))
            anchors = orc.anchors_generate(
                feature_maps_shapes=tuple(shapes), feature_maps_aspect_ratios=ratios,
                boxes_scales=(float(rng.uniform(0.35, 0.5)), float(rng.uniform(0.6, 1.1))),
                centers_padding_from_borders_percentage=float(rng.uniform(0.0, 0.2)),
                additional_square_box=bool(rng.integers(0, 2)), image_shape=image_hw)
            if rng.integers(0, 4) == 0:  # the table as the centroid constructor derives it (datacoder.py:84-89)
                cx, cy, w_, h_ = orc.corners_to_centroids(*[anchors[:, i] for i in range(4)])
                anchors = np.stack(orc.centroids_to_corners(cx, cy, w_, h_), 1).astype(np.float32)
        elif kind == "subset":
            A = int(rng.choice([1, 31, 64, 65, 700, 4097, 9600]))
            anchors = std_anchors[rng.permutation(9600)[:A]]
        elif kind == "grid":  # integer grid of equal boxes: exact ties everywhere
            n = int(rng.integers(2, 24)); sz = float(rng.integers(8, 60)); step = float(rng.integers(4, 40))
            xs, ys = np.meshgrid(np.arange(n) * step, np.arange(n) * step)
            anchors = np.stack([xs.ravel(), ys.ravel(), xs.ravel() + sz - 1, ys.ravel() + sz - 1], 1).astype(np.float32)
        else:
            A = int(rng.integers(1, 3000))
            c = rng.uniform(0, 640, size=(A, 2)); wh = rng.uniform(4, 300, size=(A, 2))
            anchors = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        A = anchors.shape[0]
        n_grid += int(len(ssdseglib._ssdx.grid_probe(anchors)) > 0)
        B = int(rng.integers(1, 6)); G = int(rng.choice([1, 2, 7, 16, 33, 64])); C = int(rng.choice([1, 2, 4, 4, 7]))
        thr = float(rng.choice([0.0, 0.3, 0.5, 0.525, 0.9]))
        labels, boxes, counts = orc.synth_gt(rng, B, G, image_hw=image_hw, n_fg=max(C - 1, 1),
                                             integer=bool(rng.integers(0, 2)),
                                             wh_range=(float(rng.choice([4.0, 20.0])), float(rng.choice([60.0, 300.0]))))
        if kind == "dbb" and rng.integers(0, 5) == 0:  # boxes partly or wholly outside the image / the grids
            boxes = boxes + rng.uniform(-0.6, 0.6, size=(B, 1, 1)).astype(np.float32) * np.float32(image_hw[1])
        if kind == "dbb" and rng.integers(0, 8) == 0:
            boxes[:, 0] = np.array([5.0, 5.0, 3.0, 30.0], np.float32)  # a GT box with a negative width
        counts = rng.integers(0, G + 1, size=B).astype(np.int32)
        if G >= 3:
            boxes[:, 2] = boxes[:, 0]                                  # duplicated GT box
        if rng.integers(0, 2):
            k = rng.integers(0, A, size=B)
            boxes[np.arange(B), G - 1] = anchors[k]                    # a GT box equal to an anchor: IoU 1
        if rng.integers(0, 4) == 0:
            labels[:, 0] = C + 3                                       # label outside [0, C): all-zero label row
        coder = ssdseglib.datacoder.DataEncoderDecoder(C, image_hw, anchors[:, 0], anchors[:, 1], anchors[:, 2],
                                                       anchors[:, 3], iou_threshold=thr)
        L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
        Lr, Or, Mr = c_oracle.encode_batch(anchors, labels, boxes, counts, C, thr, STD)
        ok = (np.array_equal(M, Mr) and np.array_equal(L, Lr) and np.allclose(O, Or, rtol=1e-5, atol=1e-6)
              and np.array_equal(O != 0, Or != 0))
        if not ok:
            bad += 1
            print(f"MISMATCH case {it}: anchors={kind} A={A} B={B} G={G} C={C} thr={thr} counts={counts.tolist()} "
                  f"match diffs={int((M != Mr).sum())}", flush=True)
    if verbose:
        print(f"fuzz_encode: {cases} cases ({n_grid} on grid anchor tables), {bad} mismatches")
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 3:  # pin one encode driver: grid_fused | grid_split | tile_fused | tile_tiled
        ssdseglib._ssdx.set_option("encode_path", sys.argv[3])
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
