"""CPU-only tests: the host-side mirror (anchors, CSV ingest, argument validation), the C-ABI surface
(library loads and exports every symbol of include/ssdx.h) and the multi-rank host logic over gloo.
No compute entry point is called here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import ssd_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")


@pytest.fixture(scope="module")
def ssdseglib():
    sys.path.insert(0, ROOT)
    import __graft_entry__

    __graft_entry__.build()
    import ssdseglib as m

    return m


def test_library_exports_every_header_symbol(ssdseglib):
    header = open(os.path.join(ROOT, "include", "ssdx.h")).read()
    declared = set(re.findall(r"SSDX_API[^;]*?\b(ssdx_\w+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(os.path.join(PKG, "lib", "libssdx.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ssdx.h but not exported"
    assert declared == set(ssdseglib._ssdx.SIGNATURES), "ctypes signatures out of sync with ssdx.h"
    assert lib.ssdx_version() == 100
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "lib", "libssdx.so")],
                         capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert {e for e in exported if e.startswith("ssdx_")} == declared  # nothing undeclared leaks out


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(PKG, "lib", "libssdx.so")], capture_output=True,
                         text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(ssdseglib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ssdseglib.datacoder.DataEncoderDecoder(4, (480, 640), a[:, 0], a[:, 1], a[:, 2], a[:, 3])
    with pytest.raises(RuntimeError):
        ssdseglib.layers.NonMaximumSuppression(4, 10, 0.5, 0.5, False)(np.zeros((1, 8, 4), np.float32),
                                                                      np.zeros((1, 8, 2), np.float32))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


# ---------------------------------------------------------------- anchors mirror (boxes.py)
def test_boxes_mirror_matches_reference_fixtures(ssdseglib, golden):
    g = golden("anchors")
    B = ssdseglib.boxes.DefaultBoundingBoxes
    std = B(((30, 40), (15, 20), (8, 10), (4, 5)), centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1),
            boxes_scales=(0.15, 0.95), additional_square_box=True)
    assert std.feature_maps_boxes is None
    std.rescale_boxes_coordinates(image_shape=(480, 640))
    assert np.array_equal(std.get_boxes_coordinates_corners("ssd"), g["std_corners"])
    assert np.array_equal(std.get_boxes_coordinates_centroids("ssd"), g["std_centroids"])
    assert np.array_equal(std.boxes_scales, g["std_scales"])
    assert np.array_equal(std.get_boxes_coordinates_xmin("ssd"), g["std_xmin"])
    assert np.array_equal(std.get_boxes_coordinates_height("ssd"), g["std_height"])
    assert np.array_equal(std.get_boxes_coordinates_corners("feature-maps")[1], g["std_fm1_corners"])
    assert np.array_equal(std.get_boxes_coordinates_width("feature-maps")[3], g["std_fm3_width"])
    assert std.get_boxes_coordinates_ymax("ssd").shape == (9600,)
    assert len(std.get_boxes_coordinates_center_x("feature-maps")) == 4
    odd = B(((3, 1), (2, 2)), ((1, 2), (1, 0.5, 3)), (0.3, 0.8), 0.1, False)
    odd.rescale_boxes_coordinates((33, 65))
    assert np.array_equal(odd.get_boxes_coordinates_corners("ssd"), g["odd_once"])
    odd.rescale_boxes_coordinates((33, 65))  # reference quirk: in-place rescale compounds
    assert np.array_equal(odd.get_boxes_coordinates_corners("ssd"), g["odd_twice"])
    c = ssdseglib.boxes.coordinates_corners_to_centroids(*[g["std_corners"][:, i] for i in range(4)])
    assert np.array_equal(np.stack(c, 1), g["conv_centroids"])
    k = ssdseglib.boxes.coordinates_centroids_to_corners(*c)
    assert np.array_equal(np.stack(k, 1), g["conv_corners"])


def test_boxes_mirror_validation(ssdseglib):
    B = ssdseglib.boxes.DefaultBoundingBoxes
    with pytest.raises(ValueError):
        B(((4, 4),), centers_padding_from_borders_percentage=0.5)
    with pytest.raises(ValueError):
        B(((4, 4), (2, 2)), centers_padding_from_borders_percentage=(0.1,))
    with pytest.raises(ValueError):
        B(((4, 4), (2, 2)), feature_maps_aspect_ratios=((1, 2),))


# ---------------------------------------------------------------- GT ingest (host C parser)
def test_csv_parser_matches_oracle(ssdseglib):
    p = ssdseglib._ssdx.parse_gt_csv
    text = "1,10.5,20,30.25,40\r\n3,1,2,3,4\r\n2,100,50.125,200,220.75\r\n"
    l, b = p(text, True)
    lo, bo = orc.parse_gt_csv_text(text)
    assert np.array_equal(l, lo) and np.array_equal(b, bo) and b.dtype == np.float32
    l, b = p(text, True, flip_width=640.0)
    assert np.array_equal(b, orc.flip_boxes_horizontal(bo, 640))
    l, b = p("1,1,2,3,4\n2,5,6,7,8\n", False)
    assert l.tolist() == [1, 2] and b[1].tolist() == [5, 6, 7, 8]
    l, b = p("  \r\n", True)
    assert l.shape == (0,) and b.shape == (0, 4)
    l, b = p("", False)
    assert l.shape == (0,)
    # repr round trip of arbitrary float32 values
    vals = np.random.default_rng(0).uniform(-50, 700, (40, 4)).astype(np.float32)
    text = "".join(f"2,{float(v[0])!r},{float(v[1])!r},{float(v[2])!r},{float(v[3])!r}\r\n" for v in vals)
    l, b = p(text, True)
    assert np.array_equal(b, vals)
    with pytest.raises(ValueError):
        p("1,2,3\r\n", True)
    with pytest.raises(ValueError):
        p("1,2,3,4,x\n", False)
    with pytest.raises(ValueError):
        p("1,1,2,3,4\n2,5,6,7,8\n", True)  # datacoder.py:195 splits on CRLF only -> one malformed record


def test_batch_csv_reader_matches_per_file_parser(ssdseglib, tmp_path):
    """ssdx_read_gt_csv_batch (C++ worker threads) against the oracle's parser, file by file: ragged counts, an empty
    file, both newline conventions, per-file flips, capacity retry, and the error path."""
    rng = np.random.default_rng(5)
    paths, texts, flips = [], [], []
    for i in range(37):
        n = int(rng.integers(0, 9)) if i != 11 else 41          # one file larger than the default capacity
        labels = rng.integers(0, 4, n)
        boxes = rng.uniform(0, 600, (n, 4)).astype(np.float32)
        rows = [f"{int(l)},{float(b[0])!r},{float(b[1])!r},{float(b[2])!r},{float(b[3])!r}" for l, b in zip(labels, boxes)]
        text = "\r\n".join(rows) + ("\r\n" if n and i % 2 else "")
        p = tmp_path / f"gt_{i}.csv"
        p.write_bytes(text.encode())
        paths.append(str(p)); texts.append(text); flips.append(bool(i % 3 == 0))
    sx = ssdseglib._ssdx
    for threads in (1, 4, 0):
        L, Bx, Cn = sx.read_gt_csv_batch(paths, crlf_only=True, flips=flips, flip_width=640.0, gmax=8, threads=threads)
        assert L.dtype == np.int32 and Bx.dtype == np.float32 and Cn.dtype == np.int32
        assert L.shape == (37, 41) and Bx.shape == (37, 41, 4)
        for i, text in enumerate(texts):
            rl, rb = orc.parse_gt_csv_text(text) if text.strip() else (np.zeros(0, np.int32), np.zeros((0, 4), np.float32))
            if flips[i]:
                rb = orc.flip_boxes_horizontal(rb, 640) if len(rb) else rb
            assert Cn[i] == len(rl)
            assert np.array_equal(L[i, :Cn[i]], rl) and np.array_equal(Bx[i, :Cn[i]], rb)
            assert not L[i, Cn[i]:].any() and not Bx[i, Cn[i]:].any()          # zero padding
    # the evaluator front-end (csv.reader conventions: "\n" or "\r\n") and the datacoder front-end agree on these files
    L2, B2, C2 = ssdseglib.evaluators.read_ground_truth(paths)
    L3, B3, C3 = sx.read_gt_csv_batch(paths, crlf_only=True)
    assert np.array_equal(C2, C3) and np.array_equal(L2, L3) and np.array_equal(B2, B3)
    with pytest.raises(ValueError, match="cannot open"):
        sx.read_gt_csv_batch(paths + [str(tmp_path / "missing.csv")], crlf_only=True)
    bad = tmp_path / "bad.csv"
    bad.write_bytes(b"1,2,3\r\n")
    with pytest.raises(ValueError, match="bad.csv"):
        sx.read_gt_csv_batch([str(bad)], crlf_only=True)
    assert sx.read_gt_csv_batch([], crlf_only=True)[2].shape == (0,)


def test_datacoder_validation_before_device(ssdseglib):
    D = ssdseglib.datacoder.DataEncoderDecoder
    a = np.zeros(4, np.float32)
    with pytest.raises(ValueError, match="corners"):
        D(4, (480, 640), xmin_boxes_default=a)
    with pytest.raises(ValueError, match="centroids"):
        D(4, (480, 640), center_x_boxes_default=a)
    with pytest.raises(ValueError):  # reference: passing both raises (unreachable branch, datacoder.py:92-108)
        D(4, (480, 640), a, a, a, a, a, a, a, a)


def test_average_precision_host_formula(ssdseglib, golden):
    """The float64 AP tail (evaluators.py:164-185) from oracle TP flags reproduces the reference's numbers."""
    g = golden("eval")
    N = g["pred_labels"].shape[0]
    no_gt = int(g["no_gt_image"][0])
    for thr, key in ((0.5, "ap50"), (0.7, "ap70")):
        tp = np.stack([orc.eval_tp_flags(g["pred_labels"][i], g["pred_boxes"][i],
                                         g["gt_labels"][i] if i != no_gt else np.zeros(0, np.int32),
                                         g["gt_boxes"][i] if i != no_gt else np.zeros((0, 4), np.float32), thr)
                       for i in range(N)])
        n_gt = {c: int(sum((g["gt_labels"][i] == c).sum() for i in range(N) if i != no_gt)) for c in (1, 2, 3)}
        sc = g["pred_labels"] != 0
        ap = ssdseglib.evaluators.average_precision_from_flags(tp[sc], g["pred_conf"][sc], g["pred_labels"][sc],
                                                               n_gt, 0)
        np.testing.assert_allclose([ap[1], ap[2], ap[3]], g[key], rtol=0, atol=1e-12)


# ---------------------------------------------------------------- multi-rank host logic (gloo, world size 2)
def test_shard_range(ssdseglib):
    from ssdseglib._dist import shard_range

    for n, w in ((1024, 8), (10, 4), (3, 8), (4096, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_all_gather_counts_gloo_world2(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {PKG!r})
        import torch, torch.distributed as dist
        from ssdseglib._dist import all_gather_counts, shard_range
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        lo, hi = shard_range(10, r, w)
        counts = torch.tensor([[r + 1, 10 * (r + 1), hi - lo]] * 4, dtype=torch.int64)
        g = all_gather_counts(counts)
        assert g.shape == (w, 4, 3)
        tot = g.sum(0)
        assert tot[0].tolist() == [3, 30, 10], tot
        dist.barrier()
        if r == 0:
            print("GATHER_OK")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------------------- grid detection and windows (host)
def _windows_cover(ssdseglib, anchors, boxes, t):
    """Every pair with IoU >= t (and > 0) lies inside the window of its (GT box, grid); returns the cell count."""
    grids = ssdseglib._ssdx.grid_probe(anchors)
    assert len(grids) > 0
    win = ssdseglib._ssdx.grid_windows(anchors, boxes, t)
    iou = orc.iou_matrix(anchors, boxes)
    covered = np.zeros(iou.shape, bool)
    cells = 0
    for g in range(boxes.shape[0]):
        for q, (first, H, W, nb) in enumerate(grids):
            r0, nr, c0, nc = win[g, q]
            if nr == 0:
                continue
            assert 0 <= r0 and r0 + nr <= H and 0 <= c0 and c0 + nc <= W
            rr, cc = np.meshgrid(np.arange(r0, r0 + nr), np.arange(c0, c0 + nc), indexing="ij")
            covered[(first + (rr * W + cc) * nb).ravel(), g] = True
            cells += nr * nc
    need = (iou >= t[None, :]) & (iou > 0)
    assert not (need & ~covered).any(), np.argwhere(need & ~covered)[:5]
    return cells, int(need.sum())


def test_grid_detection_on_reference_tables(ssdseglib, std_anchors):
    probe = ssdseglib._ssdx.grid_probe
    g = probe(std_anchors)  # notebook-03 table: 4 maps x 6 shapes (boxes.py:198-205 order)
    assert len(g) == 24 and g[0].tolist() == [0, 30, 40, 6] and g[6].tolist() == [7200, 15, 20, 6]
    assert g[18].tolist() == [9480, 4, 5, 6] and g[23].tolist() == [9485, 4, 5, 6]
    g = probe(orc.anchors_generate(**orc.STRESS_ANCHOR_CFG))
    assert [r.tolist() for r in g] == [[k, 125, 160, 5] for k in range(5)]
    # the table as the centroid constructor derives it (datacoder.py:84-89) keeps the structure
    cx, cy, w, h = orc.corners_to_centroids(*[std_anchors[:, i] for i in range(4)])
    assert len(probe(np.stack(orc.centroids_to_corners(cx, cy, w, h), 1).astype(np.float32))) == 24
    # anything else is not a grid: subsets, permutations, a perturbed coordinate, random boxes
    rng = np.random.default_rng(0)
    assert len(probe(std_anchors[rng.permutation(9600)[:700]])) == 0
    assert len(probe(std_anchors[::-1])) == 0
    bad = std_anchors.copy()  # one perturbed coordinate in row 18 of map 0: rows 0-17 | row 18 | rows 19-29 | maps 1-3
    bad[4321, 0] = np.nextafter(bad[4321, 0], np.float32(1e9))
    g = probe(bad)
    assert len(g) == 36 and g[0].tolist() == [0, 18, 40, 6] and g[6].tolist() == [4320, 1, 40, 6]
    assert g[12].tolist() == [4560, 11, 40, 6]
    _windows_cover(ssdseglib, bad, orc.synth_gt(rng, 1, 16)[1][0], np.full(16, 0.4, np.float32))
    assert len(probe(rng.uniform(0, 500, size=(1000, 4)).astype(np.float32))) == 0
    bad = std_anchors.copy()
    bad[17, 2] = np.nan
    assert len(probe(bad)) == 0


def test_grid_windows_are_conservative_and_tight(ssdseglib, std_anchors):
    rng = np.random.default_rng(1993)
    _, boxes, _ = orc.synth_gt(rng, 12, 16)
    cells = pairs = 0
    for b in range(12):
        for t in (np.full(16, 0.5, np.float32), rng.uniform(0, 0.9, 16).astype(np.float32), np.zeros(16, np.float32)):
            c, p = _windows_cover(ssdseglib, std_anchors, boxes[b], t)
            if t[0] == 0.5 and np.all(t == 0.5):
                cells += c
                pairs += p
    assert cells < 1.6 * pairs, (cells, pairs)  # ~250 cells for ~200 pairs above 0.5 per image
    # integer coordinates (boxes that coincide with cell borders), boxes outside the image, tiny and huge boxes
    _, boxes, _ = orc.synth_gt(rng, 4, 16, integer=True)
    boxes[1] += np.float32(500.0)
    boxes[2, :, 2:] = boxes[2, :, :2] + np.float32(0.25)
    boxes[3, :8] = np.array([-300.0, -200.0, 1500.0, 900.0], np.float32)
    for b in range(4):
        for t in (0.0, 0.3, 0.5, 0.95):
            _windows_cover(ssdseglib, std_anchors, boxes[b], np.full(16, t, np.float32))


def test_grid_windows_random_tables(ssdseglib):
    rng = np.random.default_rng(7)
    done = 0
    for it in range(40):
        shapes, h, w = [], int(rng.integers(6, 30)), int(rng.integers(6, 36))
        for _ in range(int(rng.integers(1, 4))):
            shapes.append((h, w))
            h, w = max(6, (h + 1) // 2), max(6, (w + 1) // 2)
        ratios = tuple(float(r) for r in rng.choice([1, 2, 3, 0.5, 1 / 3, 1.5], size=int(rng.integers(1, 5)),
                                                     replace=False))
        hw = (int(rng.integers(64, 1200)), int(rng.integers(64, 1400)))
        anchors = orc.anchors_generate(
            feature_maps_shapes=tuple(shapes), feature_maps_aspect_ratios=ratios,
            boxes_scales=(float(rng.uniform(0.35, 0.5)), float(rng.uniform(0.6, 1.1))),
            centers_padding_from_borders_percentage=float(rng.uniform(0.0, 0.2)),
            additional_square_box=bool(rng.integers(0, 2)), image_shape=hw)
        if len(ssdseglib._ssdx.grid_probe(anchors)) == 0:
            continue
        done += 1
        _, boxes, _ = orc.synth_gt(rng, 1, 12, image_hw=hw, integer=bool(rng.integers(0, 2)),
                                   wh_range=(4.0, float(rng.choice([60.0, 600.0]))))
        boxes = boxes[0] + np.float32(rng.uniform(-0.3, 0.3) * hw[1] * rng.integers(0, 2))
        _windows_cover(ssdseglib, anchors, boxes, rng.choice([0.0, 0.1, 0.5, 0.7], size=12).astype(np.float32))
    assert done >= 30


# ------------------------------------------------------------------------------- Keras-graph form of the layers
def _stub_tensorflow():
    """A minimal stand-in for the parts of TensorFlow keras_layers.py touches (TensorFlow itself is not installable
    in the build / test images): Layer with Keras' call protocol, register_keras_serializable, numpy_function."""
    import types

    registry = {}

    class Layer:
        def __init__(self, name=None, trainable=True, **kwargs):
            assert not kwargs, kwargs
            self.name, self.trainable = name, trainable

        def __call__(self, *args, **kwargs):
            return self.call(*args, **kwargs)

    class FakeTensor:
        def __init__(self, value):
            self.value = np.asarray(value)
            self.shape = self.value.shape

        def set_shape(self, shape):
            self.static_shape = tuple(shape)

    def register(package="Custom", name=None):
        def deco(cls):
            registry[name or cls.__name__] = cls
            return cls
        return deco

    def numpy_function(func, inp, Tout, name=None):
        return FakeTensor(func(*[t.value for t in inp]))

    tf = types.SimpleNamespace(
        keras=types.SimpleNamespace(layers=types.SimpleNamespace(Layer=Layer),
                                    saving=types.SimpleNamespace(register_keras_serializable=register)),
        is_tensor=lambda x: isinstance(x, FakeTensor), numpy_function=numpy_function, float32="float32")
    return tf, registry, FakeTensor


def test_keras_layer_forms_mirror_the_reference_contract(ssdseglib):
    """models.py:378-417 constructs the layers with keyword arguments + name=, sets .trainable, calls them on graph
    tensors and relies on get_config / the registered names for save + load_model (layers.py:4,83-93,95,170-177,179)."""
    from ssdseglib import keras_layers

    tf, registry, FakeTensor = _stub_tensorflow()
    calls = []

    class PlainNms:
        def __init__(self, *a):
            self.args = a

        def call(self, boxes, probs):
            calls.append((boxes.dtype, probs.dtype, boxes.shape))
            return np.zeros((boxes.shape[0], self.args[1], 6), np.float32)

        def call_on_logits(self, boxes, logits):
            return "logits-path"

    class PlainDecode:
        def __init__(self, *a):
            self.args = a

        def call(self, off):
            return off * 2

    Dec, Nms, Seg = keras_layers.make_keras_layers(tf, PlainDecode, PlainNms, lambda: None)
    assert set(registry) == {"DecodeBoxesCentroidsOffsets", "NonMaximumSuppression", "SegmentationSuppression"}
    assert issubclass(Nms, tf.keras.layers.Layer) and issubclass(Dec, tf.keras.layers.Layer)
    nms = Nms(max_number_of_boxes_per_class=4, max_number_of_boxes_per_sample=10, boxes_iou_threshold=0.025,
              labels_probability_threshold=0.725, suppress_background_boxes=False, name="output-object-detection")
    nms.trainable = False
    assert nms.get_config() == {"max_number_of_boxes_per_class": 4, "max_number_of_boxes_per_sample": 10,
                                "boxes_iou_threshold": 0.025, "labels_probability_threshold": 0.725,
                                "suppress_background_boxes": False}
    assert Nms(**nms.get_config()).get_config() == nms.get_config()  # from_config round trip
    out = nms(boxes_corners_coordinates=FakeTensor(np.zeros((3, 50, 4), np.float64)),
              labels_probabilities=FakeTensor(np.zeros((3, 50, 4), np.float64)))
    assert isinstance(out, FakeTensor) and out.static_shape == (3, 10, 6)
    assert calls == [(np.dtype("float32"), np.dtype("float32"), (3, 50, 4))]  # fp32 host arrays reach the C-ABI path
    assert Nms(4, 10, 0.1, 0.5, True)(FakeTensor(np.zeros((2, 5, 4))), FakeTensor(np.zeros((2, 5, 4)))).static_shape \
        == (None, 6)
    # non-TensorFlow inputs bypass the graph machinery; the extra entry points stay reachable
    assert isinstance(nms(np.zeros((2, 5, 4), np.float32), np.zeros((2, 5, 4), np.float32)), np.ndarray)
    assert nms.call_on_logits(None, None) == "logits-path"
    cx = np.arange(5, dtype=np.float32)
    dec = Dec(center_x_boxes_default=cx, center_y_boxes_default=cx, width_boxes_default=cx, height_boxes_default=cx,
              standard_deviation_center_x_offsets=0.1, standard_deviation_center_y_offsets=0.1,
              standard_deviation_width_offsets=0.2, standard_deviation_height_offsets=0.2, name="decode-output-boxes")
    cfg = dec.get_config()
    assert list(cfg) == ["center_x_boxes_default", "center_y_boxes_default", "width_boxes_default",
                         "height_boxes_default", "standard_deviation_center_x_offsets",
                         "standard_deviation_center_y_offsets", "standard_deviation_width_offsets",
                         "standard_deviation_height_offsets"] and cfg["center_x_boxes_default"] is cx
    y = dec(FakeTensor(np.ones((2, 5, 4))))
    assert y.static_shape == (2, 5, 4) and np.all(y.value == 2.0)
    assert Seg(name="segmentation-suppression").get_config() == {}
    # the package itself: plain callables here (no TensorFlow), aliases always present
    assert ssdseglib.layers.NonMaximumSuppression is ssdseglib.layers.PlainNonMaximumSuppression


def test_nccl_binding_is_optional_and_fails_cleanly(ssdseglib):
    """libssdx binds NCCL at first use with dlopen (no link dependency): asking for the version either works (a
    libnccl.so.2 is loadable in this process) or returns SSDX_ERR_NCCL with a message - never a crash; a NULL
    communicator is an invalid argument."""
    sx = ssdseglib._ssdx
    needed = subprocess.run(["readelf", "-d", os.path.join(PKG, "lib", "libssdx.so")], capture_output=True,
                            text=True).stdout
    assert "libnccl" not in needed  # no DT_NEEDED entry for NCCL
    ver = ctypes.c_int(-1)
    st = sx.lib.ssdx_nccl_version(ctypes.byref(ver))
    assert st in (0, sx.SSDX_ERR_NCCL)
    if st == 0:
        assert ver.value >= 20000
    else:
        assert b"NCCL" in sx.lib.ssdx_last_error()
    assert sx.lib.ssdx_allgather_counts(None, None, None, 4, None) == sx.SSDX_ERR_INVALID
    with pytest.raises(ValueError):
        sx.set_option("encode_path", "no-such-path")
    sx.set_option("encode_path", "auto")
    # dispatch pins of the NMS entry points: every documented value is accepted, anything else is refused
    for key, values in (("nms_path", ("perclass", "pcta", "generic", "blocked", "auto")),
                        ("nms_cta", ("256", "512", "1024", "auto"))):
        for v in values:
            sx.set_option(key, v)
        with pytest.raises(ValueError):
            sx.set_option(key, "300")
    with pytest.raises(ValueError):
        sx.set_option("no-such-key", "auto")


# ---- design invariants of the merged NMS kernel, restated in NumPy (csrc/nms.cu: make_hbox_cls, class_unit, rank_sort_512)
def _f16_down(x64):
    """float64 -> float16 rounded towards -inf (what __float2half_rd(__fadd_rd(.)) yields for these exact sums)."""
    with np.errstate(over="ignore"):  # (beyond the fp16 range: +-inf, stepped back to the largest finite value)
        h = x64.astype(np.float16)
    bad = h.astype(np.float64) > x64
    h[bad] = np.nextafter(h[bad], np.float16(-np.inf))
    return h


def _f16_up(x64):
    with np.errstate(over="ignore"):
        h = x64.astype(np.float16)
    bad = h.astype(np.float64) < x64
    h[bad] = np.nextafter(h[bad], np.float16(np.inf))
    return h


def _class_unit(ext):
    e = int((np.float32(ext).view(np.uint32) >> 23) & 0xFF) - 127 + 2
    return np.float64(2.0 ** min(max(e, -8), 14))


@pytest.mark.parametrize("scale", [1.0, 640.0, 3000.0])
def test_class_shifted_fp16_prefilter_is_conservative_and_separates_classes(scale):
    """Prefilter of ssdx_nms_fast_all_kernel: boxes of class c are shifted by unit * (c & 15, c >> 4) and rounded
    outwards to fp16.  (1) Same class: exact fp32 overlap implies prefilter overlap (never a false negative).
    (2) Different classes: no pair passes while the shifted coordinates stay inside the fp16 range."""
    rng = np.random.default_rng(int(scale))
    n, C = 400, 40
    c0 = rng.uniform(-0.2, 1.2, size=(n, 2)) * scale
    wh = rng.uniform(0.0, 0.5, size=(n, 2)) * scale
    box = np.concatenate([c0 - wh / 2, c0 + wh / 2], 1).astype(np.float32)  # normalised (y0, x0, y1, x1)
    box[::7, 2] = box[::7, 0]  # degenerate
    cls = rng.integers(0, C, size=n)
    unit = _class_unit(np.abs(box).max())
    oy, ox = unit * (cls & 15), unit * (cls >> 4)
    b64 = box.astype(np.float64)
    mn = np.stack([_f16_down(b64[:, 0] + oy), _f16_down(b64[:, 1] + ox)], 1).astype(np.float64)
    mx = np.stack([_f16_up(b64[:, 2] + oy), _f16_up(b64[:, 3] + ox)], 1).astype(np.float64)
    pre = ((mx[:, None, 0] > mn[None, :, 0]) & (mx[None, :, 0] > mn[:, None, 0]) &
           (mx[:, None, 1] > mn[None, :, 1]) & (mx[None, :, 1] > mn[:, None, 1]))
    exact = ((box[:, None, 2] > box[None, :, 0]) & (box[None, :, 2] > box[:, None, 0]) &
             (box[:, None, 3] > box[None, :, 1]) & (box[None, :, 3] > box[:, None, 1]))
    same = cls[:, None] == cls[None, :]
    assert not np.any(exact & same & ~pre)
    in_range = 15 * unit + np.abs(box).max() < 65504  # (beyond it the shifts saturate: only more exact tests)
    if in_range:
        assert not np.any(pre & ~same)
    else:
        assert scale == 3000.0


def test_rank_sort_bins_are_monotone_and_ranks_exact():
    """rank_sort_512 (csrc/nms.cu): bin = min(255, int(float32(key - kmin) * inv)) must be non-decreasing in the key, so
    that (keys in higher bins) + (larger keys in the own bin) is the exact descending rank."""
    rng = np.random.default_rng(5)
    for trial in range(50):
        n = int(rng.integers(2, 513))
        kind = trial % 4
        if kind == 0:
            hi = rng.integers(0x80000000, 0xBF800000, size=n, dtype=np.uint64)          # any positive score
        elif kind == 1:
            hi = np.uint64(0xBF000000) + rng.integers(0, 4096, size=n, dtype=np.uint64)  # crowded
        elif kind == 2:
            hi = np.full(n, 0xBF7FFFFF, np.uint64)                                       # all equal
        else:
            hi = rng.choice(np.array([1, 0x80000000, 0xFF7FFFFF, 0xFF800000], np.uint64), size=n)  # extremes
        lo = rng.permutation(1 << 20)[:n].astype(np.uint64)                               # distinct composites
        keys = (hi << np.uint64(32)) | lo
        kmin, kmax = hi.min(), hi.max()
        inv = np.float32(256.0) / (np.float32(np.float32(kmax - kmin)) + np.float32(1.0))
        bins = np.minimum(255, (np.float32((hi - kmin).astype(np.float32)) * inv).astype(np.int64))
        order = np.argsort(hi, kind="stable")
        assert np.all(np.diff(bins[order]) >= 0)
        rank = np.array([(bins > bins[i]).sum() + ((bins == bins[i]) & (keys > keys[i])).sum() for i in range(n)])
        assert np.array_equal(np.sort(rank), np.arange(n))
        assert np.array_equal(keys[np.argsort(rank)], np.sort(keys)[::-1])


def test_chunk_fix_point_equals_sequential_greedy():
    """Resolve step of ssdx_nms_fast_all_kernel (csrc/nms.cu, step (c)): 64 candidates at a time, candidates suppressed
    by boxes kept in earlier 64-blocks are marked first, then rounds of: undecided + a kept suppressor -> suppressed;
    undecided + every suppressor decided as suppressed -> kept.  Must equal the sequential greedy pass."""
    rng = np.random.default_rng(17)
    for trial in range(40):
        m = int(rng.integers(1, 257))
        dens = float(rng.choice([0.0, 0.01, 0.05, 0.3, 0.9]))
        M = np.tril(rng.random((m, m)) < dens, -1)            # M[i, j]: earlier candidate j suppresses i when kept
        pre = rng.random(m) < float(rng.choice([0.0, 0.2]))   # suppressed by boxes of earlier chunks
        kept_seq = np.zeros(m, bool)
        for i in range(m):
            kept_seq[i] = not pre[i] and not np.any(M[i, :i] & kept_seq[:i])
        kept = np.zeros(m, bool)
        for s0 in range(0, m, 64):
            blk = slice(s0, min(s0 + 64, m))
            sup = pre[blk] | np.any(M[blk, :s0] & kept[None, :s0], axis=1)
            k = np.zeros(sup.size, bool)
            Mb = M[blk, blk]
            rounds = 0
            while not np.all(k | sup):
                und = ~(k | sup)
                hit = np.any(Mb & k[None, :], axis=1)
                free = ~np.any(Mb & ~sup[None, :], axis=1)
                new_k, new_s = und & ~hit & free, und & hit
                assert new_k.any() or new_s.any()               # the lowest undecided candidate is always decided
                k |= new_k
                sup |= new_s
                rounds += 1
            kept[blk] = k
        assert np.array_equal(kept, kept_seq)


def test_sample_sort_network_sorts():
    """Exchange pattern of ssdx_ap_sample_kernel (csrc/ap.cu): 4096 keys, key e = r * 1024 + tid, stage (k, j) pairs e
    with e ^ j and keeps the minimum where ((e & j) == 0) == ((e & k) == 0): ascending order at the end."""
    rng = np.random.default_rng(23)
    v = rng.integers(0, 1 << 62, size=4096, dtype=np.int64)
    v[0:4000:5] = v[1:4001:5]  # duplicates
    e = np.arange(4096)
    k = 2
    while k <= 4096:
        j = k >> 1
        while j > 0:
            p = v[e ^ j]
            take_min = ((e & j) == 0) == ((e & k) == 0)
            v = np.where(take_min, np.minimum(v, p), np.maximum(v, p))
            j >>= 1
        k <<= 1
    assert np.all(np.diff(v) >= 0)
