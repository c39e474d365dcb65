"""GPU parity tests: the CUDA path (through the C ABI, via the ssdseglib mirror) against the CPU oracle and
the committed golden fixtures.  Bar: matched indices / labels / NMS keep-sets / TP flags / counts bit-exact;
offsets and decoded coordinates within 1e-5 relative in fp32 (absolute floors stated at each check)."""
import os

import numpy as np
import pytest

from oracle import ssd_oracle as orc

pytestmark = pytest.mark.gpu

STD = (0.1, 0.1, 0.2, 0.2)
RTOL = 1e-5  # BASELINE.json north_star: "within 1e-5 relative in fp32"


@pytest.fixture(scope="module")
def lib():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ssdseglib

    return ssdseglib


@pytest.fixture(autouse=True, params=["grid_fused", "grid_split", "tile_fused", "tile_tiled"])
def encode_path(request, lib):
    """Every encode test runs against the four encode drivers: closed-form windows on grid anchor tables (one CTA
    per image / split into GT-side and row-span kernels) and the tile match for arbitrary tables (one CTA per image /
    anchor-tiled + emit).  A forced grid driver falls back to the tile match when the table is not a grid."""
    if request.param != "grid_fused" and "encode" not in request.node.name:
        pytest.skip("encode driver does not matter for this test")
    lib._ssdx.set_option("encode_path", request.param)
    yield request.param
    lib._ssdx.set_option("encode_path", "auto")


@pytest.fixture(autouse=True, params=["fast", "fast_256t", "fast_512t", "blocked", "perclass", "pcta", "generic"])
def nms_path(request, lib):
    """NMS tests run on the default dispatch (merged one-pass-per-image pipeline when max_per_class >= max_total,
    else the per-class pipeline: one warp per segment for caps up to 32, one CTA per segment above; generic only for
    handed-over work), at the pinned CTA widths of the merged pipeline ("fast_256t": the kernel of the bench's batch
    size; "fast_512t"), with the merged pipeline's 64-candidate blocked greedy pass instead of the chunked bit matrix
    ("blocked"), with the per-class pipeline forced ("perclass"), with the per-class CTA kernel forced for every
    cap ("pcta"), and with the exact generic kernel alone."""
    if request.param != "fast":
        if "nms" not in request.node.name:
            pytest.skip("NMS driver does not matter for this test")
        if request.param.startswith("fast_"):  # the default pipeline at a pinned CTA width (auto: 1024 threads per
            lib._ssdx.set_option("nms_cta", request.param[5:-1])  # image at these batch sizes)
        else:
            lib._ssdx.set_option("nms_path", request.param)
    yield request.param
    lib._ssdx.set_option("nms_path", "auto")
    lib._ssdx.set_option("nms_cta", "auto")


@pytest.fixture(scope="module")
def coder(lib, std_anchors):
    return lib.datacoder.DataEncoderDecoder(
        num_classes=4, image_shape=(480, 640), xmin_boxes_default=std_anchors[:, 0],
        ymin_boxes_default=std_anchors[:, 1], xmax_boxes_default=std_anchors[:, 2],
        ymax_boxes_default=std_anchors[:, 3], iou_threshold=0.5, standard_deviations_centroids_offsets=STD)


def assert_offsets_close(got, ref):
    # log() differs from NumPy/Eigen by a few ulp; tx,ty are exact IEEE divisions
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-6)


def assert_coords_close(got, ref, anchor_wh):
    # (exp(t)-1)*w cancels for tiny t: absolute floor 1e-5 * anchor size, as SURVEY.md section 7
    floor = 1e-5 * np.maximum(anchor_wh.max(axis=-1), 1.0)[..., None]
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert np.all(err <= RTOL * np.abs(ref) + floor), f"max err {err.max()}"


def encode_ref(anchors, labels, boxes, counts, C, thr):
    return orc.encode_batch(anchors, labels, boxes, counts, C, thr, STD)


# ------------------------------------------------------------------------------------------- encode
CASES = ["cfg0", "cfg0_thr525_flip", "ties_int", "small", "label_oob"]


@pytest.mark.parametrize("case", CASES)
def test_encode_golden(lib, golden, std_anchors, case):
    g = golden("encode_decode")
    labels, boxes = g[case + "_gt_labels"], g[case + "_gt_boxes"].copy()
    thr, flip = float(g[case + "_thr"][0]), bool(g[case + "_flip"][0])
    if flip:
        boxes = np.stack([orc.flip_boxes_horizontal(b, 640) for b in boxes])
    c = lib.datacoder.DataEncoderDecoder(4, (480, 640), std_anchors[:, 0], std_anchors[:, 1], std_anchors[:, 2],
                                         std_anchors[:, 3], iou_threshold=thr)
    L, O, M = c.encode_batch(boxes, labels, return_match=True)
    assert L.dtype == np.float32 and O.dtype == np.float32 and M.dtype == np.int32
    assert np.array_equal(L, g[case + "_labels"])                      # labels bit-exact
    assert np.array_equal(O != 0, g[case + "_offsets"] != 0)
    assert_offsets_close(O, g[case + "_offsets"])
    _, _, Mref = encode_ref(std_anchors, labels, boxes, np.full(len(labels), labels.shape[1]), 4, thr)
    assert np.array_equal(M, Mref)                                      # matched indices bit-exact


def test_encode_config0_and_ragged(lib, coder, std_anchors):
    rng = np.random.default_rng(1993)
    labels, boxes, counts = orc.synth_gt(rng, 8, 16)
    counts = np.array([16, 0, 1, 7, 16, 3, 0, 12], np.int32)          # ragged + empty images
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)
    assert np.all(L[1, :, 0] == 1) and np.all(O[1] == 0) and np.all(M[1] == -1)


def test_encode_tie_heavy_and_duplicates(lib, coder, std_anchors):
    rng = np.random.default_rng(77)
    labels, boxes, counts = orc.synth_gt(rng, 24, 12, integer=True)
    boxes[:, 5] = boxes[:, 2]
    boxes[:, 9] = boxes[:, 2]
    labels[:, 5] = 1
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)


@pytest.mark.parametrize("C,thr", [(1, 0.5), (3, 0.3), (6, 0.525), (21, 0.5)])
def test_encode_other_class_counts(lib, std_anchors, C, thr):
    rng = np.random.default_rng(C)
    labels, boxes, counts = orc.synth_gt(rng, 5, 9, n_fg=max(C - 1, 1))
    c = lib.datacoder.DataEncoderDecoder(C, (480, 640), std_anchors[:, 0], std_anchors[:, 1], std_anchors[:, 2],
                                         std_anchors[:, 3], iou_threshold=thr)
    L, O, M = c.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, C, thr)
    assert L.shape == (5, 9600, C)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)


@pytest.mark.parametrize("A,C", [(9563, 4), (9563, 3), (1001, 5), (63, 4)])
def test_encode_odd_anchor_counts(lib, std_anchors, A, C):
    """Anchor counts that are not multiples of the tile / vector sizes (padding slots of the match order, partial
    background blocks, label rows that cannot be bulk-stored because the image stride is not 16-byte aligned)."""
    rng = np.random.default_rng(A + C)
    anchors = std_anchors[rng.permutation(9600)[:A]]          # arbitrary order: the match order is built inside
    labels, boxes, counts = orc.synth_gt(rng, 150, 12, n_fg=max(C - 1, 1))
    counts[::7] = 0
    c = lib.datacoder.DataEncoderDecoder(C, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
    L, O, M = c.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(anchors, labels, boxes, counts, C, 0.5)
    assert L.shape == (150, A, C) and np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)
    # decode on the same odd shape: batched layer form and the masked single-image form
    cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1).astype(np.float32)
    layer = lib.layers.DecodeBoxesCentroidsOffsets(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], *STD)
    assert_coords_close(layer(O), orc.decode_batched_yxyx(Or, cent, STD), cent[None, :, 2:])
    assert_coords_close(c.decode_to_corners(O[3]), orc.decode_to_corners(Or[3], cent, STD), cent[:, 2:])


def test_encode_randomized_differential(lib, encode_path):
    """150 random cases (anchor subsets, integer grids with exact ties, random anchors; ragged / empty GT, duplicated GT
    boxes, GT equal to an anchor, thresholds, class counts, labels out of range) against the C oracle on the encode
    driver the fixture forces (tools/fuzz_encode.py)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "fuzz_encode", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_encode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    seed = {"grid_fused": 21, "grid_split": 22, "tile_fused": 23, "tile_tiled": 24}[encode_path]
    assert mod.run(150, seed, verbose=False) == 0


def test_encode_stress_shape_small_batch(lib):
    """BASELINE config 3 shape (A = 100k, G = 256) at batch 2 against the oracle."""
    anchors = orc.anchors_generate(**orc.STRESS_ANCHOR_CFG)
    rng = np.random.default_rng(1996)
    labels, boxes, counts = orc.synth_gt(rng, 2, 256, image_hw=(1000, 1280))
    c = lib.datacoder.DataEncoderDecoder(4, (1000, 1280), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
    L, O, M = c.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(anchors, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)


def test_encode_more_gt_than_one_tile(lib, coder, std_anchors):
    rng = np.random.default_rng(3)
    labels, boxes, counts = orc.synth_gt(rng, 2, 300, wh_range=(10.0, 120.0))  # > 256 GT: two shared-memory tiles
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)


def test_encode_per_image_file_api(lib, coder, std_anchors, tmp_path, golden):
    g = golden("encode_decode")
    p = tmp_path / "gt.csv"
    rows = [f"{int(l)},{float(b[0])!r},{float(b[1])!r},{float(b[2])!r},{float(b[3])!r}"
            for l, b in zip(g["cfg0_gt_labels"][0], g["cfg0_gt_boxes"][0])]
    p.write_bytes(("\r\n".join(rows) + "\r\n").encode())
    L, O = coder._encode_ground_truth_labels_boxes(str(p), False)
    assert np.array_equal(L, g["cfg0_labels"][0])
    assert_offsets_close(O, g["cfg0_offsets"][0])
    # the body of the tf.numpy_function inside read_and_encode: bytes path + NumPy flag, as TensorFlow hands them over
    Ln, On = coder._encode_np(np.array(str(p).encode()), np.bool_(False))
    assert Ln.dtype == np.float32 and np.array_equal(Ln, L) and np.array_equal(On, O)
    Lf, Of = coder._encode_ground_truth_labels_boxes(str(p), True)
    Lr, Or, _ = orc.encode_image(std_anchors, g["cfg0_gt_labels"][0],
                                 orc.flip_boxes_horizontal(g["cfg0_gt_boxes"][0], 640), 4, 0.5, STD)
    assert np.array_equal(Lf, Lr)
    assert_offsets_close(Of, Or)


# ------------------------------------------------------------------------------------------- decode
def test_decode_golden_and_modes(lib, coder, golden):
    g = golden("encode_decode")
    anc = golden("anchors")["std_centroids"]
    wh = anc[:, 2:]
    for case in CASES:
        off = g[case + "_offsets"]
        cen = coder.decode_to_centroids(off)
        cor = coder.decode_to_corners(off)
        assert cen.shape == off.shape and cor.shape == off.shape
        bg = np.abs(off).sum(-1) == 0
        assert np.all(cen[bg] == 0) and np.all(cor[bg] == 0)           # background rows exactly zero
        assert_coords_close(cen, g[case + "_dec_centroids"], wh[None])
        assert_coords_close(cor, g[case + "_dec_corners"], wh[None])
    x0, y0, x1, y1 = coder.decode_to_corners(g["cfg0_offsets"][0], output_decoded_corners_separately=True)
    assert x0.shape == (9600,)


def test_decode_layer_yxyx(lib, golden):
    g = golden("layers")
    a = golden("anchors")["std_centroids"]
    layer = lib.layers.DecodeBoxesCentroidsOffsets(a[:, 0], a[:, 1], a[:, 2], a[:, 3], 0.1, 0.1, 0.2, 0.2)
    out = layer(g["offsets"])
    assert out.shape == g["decoded_yxyx"].shape and out.dtype == np.float32
    assert_coords_close(out, g["decoded_yxyx"], a[None, :, 2:] * 8.0)  # decoded boxes ~ 2x anchor size here
    # xy offsets are exact arithmetic (no transcendental): centre = (min + max) / 2 must be bit-stable
    rng = np.random.default_rng(0)
    off = rng.normal(0, 1, (3, 9600, 4)).astype(np.float32)
    np.testing.assert_allclose(layer(off), orc.decode_batched_yxyx(off, a, STD), rtol=RTOL, atol=1e-3)


def test_encode_decode_round_trip_full_batch(lib, coder, std_anchors):
    """Size-independent property at the bench batch size (1024): decoding the encoded targets returns the
    matched ground-truth box; background rows decode to zero."""
    rng = np.random.default_rng(2024)
    labels, boxes, counts = orc.synth_gt(rng, 1024, 16)
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    dec = coder.decode_to_corners(O)
    assert dec.shape == (1024, 9600, 4)
    pos = M >= 0
    assert 100 * 1024 < pos.sum() < 400 * 1024
    gt = np.take_along_axis(boxes, np.maximum(M, 0)[..., None].repeat(4, -1), axis=1)
    np.testing.assert_allclose(dec[pos], gt[pos], atol=5e-3, rtol=0)
    assert np.all(dec[~pos] == 0)
    assert np.array_equal(L.sum(-1), np.ones((1024, 9600), np.float32))
    lab_gt = np.take_along_axis(labels, np.maximum(M, 0), axis=1)
    assert np.array_equal(L.argmax(-1)[pos], lab_gt[pos]) and np.all(L[~pos][:, 0] == 1)
    # every GT box got its best anchor (step 1) unless a later GT took the same anchor
    assert np.all([(np.isin(np.arange(16), M[b][M[b] >= 0])).sum() >= 14 for b in range(0, 1024, 97)])


# ------------------------------------------------------------------------------------------- NMS
def nms_check(lib, yxyx, scores, mpc, mtot, iou, thr):
    layer = lib.layers.NonMaximumSuppression(mpc, mtot, iou, thr, False)
    out, valid, aidx = layer.call(yxyx, scores, return_aux=True)
    ref, vref, aref = orc.combined_nms(yxyx, scores, mpc, mtot, iou, thr)
    assert np.array_equal(valid, vref)
    assert np.array_equal(aidx, aref)          # keep-set and order bit-exact
    assert np.array_equal(out, ref)            # rows are copies of inputs: bit-exact
    return out


def test_nms_golden_layer_output(lib, golden):
    g = golden("layers")
    out = nms_check(lib, g["decoded_yxyx"], g["scores"], 200, 200, 0.45, 0.01)
    assert np.array_equal(out, g["det_200"])
    layer = lib.layers.NonMaximumSuppression(4, 10, 0.025, 0.2, True)
    assert np.array_equal(layer(g["decoded_yxyx"], g["scores"]), g["det_bgdrop"])


def test_nms_config1_batch(lib, golden):
    """BASELINE config 1 generator (decode + NMS, thr .01, IoU .45, top-200) at batch 8."""
    a = golden("anchors")["std_centroids"]
    off, scores = orc.synth_predictions(np.random.default_rng(1994), 8, 9600, 4)
    yxyx = orc.decode_batched_yxyx(off, a, STD)
    out = nms_check(lib, yxyx, scores, 200, 200, 0.45, 0.01)
    assert np.all(np.diff(out[..., 1], axis=1) <= 0)  # scores non-increasing


@pytest.mark.parametrize("mpc,mtot,iou,thr", [(4, 10, 0.025, 0.725), (50, 20, 0.3, 0.3), (300, 1000, 0.6, 0.05),
                                              (1, 1, 0.0, 0.0), (200, 200, 1.0, 0.01),
                                              (500, 300, 0.5, 0.05),     # merged pipeline, 256 < max_total <= 512
                                              (600, 600, 0.5, 0.05)])    # more than 512 boxes per class: generic kernel
def test_nms_parameter_grid(lib, golden, mpc, mtot, iou, thr):
    a = golden("anchors")["std_centroids"]
    off, scores = orc.synth_predictions(np.random.default_rng(mpc), 3, 9600, 4)
    nms_check(lib, orc.decode_batched_yxyx(off, a, STD), scores, mpc, mtot, iou, thr)


def test_nms_deep_suppression_clusters(lib):
    """Clustered boxes: thousands of candidates are visited before 200 survive (several selection rounds)."""
    rng = np.random.default_rng(9)
    B, A, C = 2, 6000, 3
    centers = rng.uniform(100, 500, size=(B, 12, 2))
    which = rng.integers(0, 12, size=(B, A))
    c = np.take_along_axis(centers, which[..., None].repeat(2, -1), axis=1) + rng.normal(0, 6, size=(B, A, 2))
    wh = rng.uniform(60, 90, size=(B, A, 2))
    yxyx = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)
    scores = orc.dedup_scores(rng.uniform(0.02, 1, size=(B, A, C)).astype(np.float32))
    out = nms_check(lib, yxyx, scores, 200, 300, 0.45, 0.01)
    assert (out[..., 1] > 0).sum() < 2 * 3 * 60  # far fewer than 200 per class survive


def test_nms_massive_score_ties(lib):
    """Thousands of scores exactly 1.0f (saturated soft-max): ties are visited lower anchor first."""
    rng = np.random.default_rng(10)
    B, A, C = 2, 5000, 2
    ctr = rng.uniform(50, 600, size=(B, A, 2))
    wh = rng.uniform(10, 40, size=(B, A, 2))
    yxyx = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    scores = np.ones((B, A, C), np.float32)
    scores[:, ::7, 1] = 0.5
    nms_check(lib, yxyx, scores, 200, 400, 0.3, 0.01)


@pytest.mark.parametrize("layout", ["spread", "stacked", "mixed"])
def test_nms_massive_score_ties_merged_pipeline(lib, layout):
    """max_per_class >= max_total (merged one-pass pipeline, no fallback kernel behind it) with far more equal
    scores than the candidate list holds: the first list overflows and the composite-band refills have to split the
    ties by class and anchor.  'stacked' = every box identical (one survivor per class after visiting all ties)."""
    rng = np.random.default_rng(11)
    B, A, C = 2, 6000, 3
    ctr = rng.uniform(50, 600, size=(B, A, 2))
    wh = rng.uniform(10, 40, size=(B, A, 2))
    if layout == "stacked":
        ctr[:], wh[:] = 300.0, 50.0
    yxyx = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    scores = np.ones((B, A, C), np.float32)
    if layout == "mixed":  # a few distinct high scores, a plateau of 0.75 that overflows, a distinct tail
        scores = orc.dedup_scores(rng.uniform(0.02, 0.7, size=(B, A, C)).astype(np.float32))
        scores[:, 100:4000, 1:] = 0.75
        scores[:, :20, 0] = np.linspace(0.99, 0.8, 20, dtype=np.float32)
    out = nms_check(lib, yxyx, scores, 400, 400, 0.3, 0.01)
    nms_check(lib, yxyx, scores, 200, 200, 0.45, 0.01)
    if layout == "stacked":
        assert ((out[..., 1] > 0).sum(axis=1) == C).all()  # one survivor per class


@pytest.mark.parametrize("mpc,mtot,iou,thr", [(4, 10, 0.025, 0.725), (1, 3, 0.0, 0.5), (32, 100, 0.2, 0.05),
                                              (8, 40, 0.6, 0.0), (20, 50, 0.3, 0.3)])
def test_nms_small_per_class_caps(lib, golden, mpc, mtot, iou, thr):
    """Per-class caps up to 32 (one warp per (image, class); the reference's own setting is 4 / 10 / 0.025 / ~0.7) on
    the bench's round-trip predictions (every anchor is a candidate of some class above 0.7: the listed tier runs dry
    and band refills follow), on clustered boxes and on thousands of equal scores."""
    a = golden("anchors")["std_centroids"]
    rng = np.random.default_rng(mpc * 7 + mtot)
    off, scores = orc.synth_predictions(rng, 3, 9600, 4)
    yxyx = orc.decode_batched_yxyx(off, a, STD)
    nms_check(lib, yxyx, scores, mpc, mtot, iou, thr)
    # near one-hot scores: ~9500 candidates of class 0 above 0.7
    onehot = np.zeros((2, 9600, 4), np.float32)
    onehot[..., 0] = 1
    onehot[:, ::53] = np.array([0, 0, 1, 0], np.float32)
    s2 = orc.dedup_scores((0.9 * onehot + 0.1 * orc.softmax_f32(rng.normal(0, 1, onehot.shape).astype(np.float32))))
    nms_check(lib, yxyx[:2], s2, mpc, mtot, iou, thr)
    # thousands of exactly equal scores, identical and disjoint boxes
    s3 = np.full((2, 9600, 4), 0.8, np.float32)
    s3[:, ::3, 1] = 0.9
    nms_check(lib, yxyx[:2], s3, mpc, mtot, iou, thr)


def test_nms_degenerate_boxes_and_small_sets(lib):
    rng = np.random.default_rng(12)
    # flipped corners + zero-area boxes (never suppress, never suppressed)
    B, A, C = 3, 300, 5
    yxyx = rng.uniform(0, 100, size=(B, A, 4)).astype(np.float32)
    yxyx[:, ::5, 2] = yxyx[:, ::5, 0]
    scores = orc.dedup_scores(rng.uniform(0, 1, size=(B, A, C)).astype(np.float32))
    nms_check(lib, yxyx, scores, 200, 200, 0.45, 0.01)   # max_per_class < A
    nms_check(lib, yxyx, scores, 500, 50, 0.45, 0.5)     # max_per_class > A
    nms_check(lib, yxyx, scores, 10, 10, 0.45, 2.0)      # nothing passes the score threshold
    out = nms_check(lib, yxyx[:, :1], scores[:, :1], 3, 4, 0.5, 0.0)  # a single box
    assert out.shape == (3, 4, 6)


def test_nms_many_classes_and_cross_class_ties(lib):
    """C = 21 (scalar score path) with max_per_class >= max_total (merged one-pass pipeline), plus scores that tie
    ACROSS classes: the merge takes the lower class first, then the lower anchor."""
    rng = np.random.default_rng(21)
    B, A, C = 3, 2500, 21
    ctr = rng.uniform(40, 600, size=(B, A, 2))
    wh = rng.uniform(20, 120, size=(B, A, 2))
    yxyx = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    scores = orc.dedup_scores(rng.uniform(0, 1, size=(B, A, C)).astype(np.float32))
    nms_check(lib, yxyx, scores, 100, 100, 0.5, 0.3)
    nms_check(lib, yxyx, scores, 300, 150, 0.45, 0.05)
    tied = scores.copy()
    tied[:, :, 5] = tied[:, :, 2]          # every score of class 5 ties with class 2 on the same anchor
    tied[:, 100:200, 7] = tied[:, 300:400, 2]  # ties across classes on different anchors
    nms_check(lib, yxyx, tied, 100, 100, 0.5, 0.3)


def test_nms_merged_runs_dry_and_refills(lib):
    """Heavily clustered boxes with max_total = max_per_class: the first list of the merged pipeline runs dry and
    the key-band refills have to fetch more candidates of all classes."""
    rng = np.random.default_rng(31)
    B, A, C = 2, 8000, 4
    centers = rng.uniform(80, 560, size=(B, 40, 2))
    which = rng.integers(0, 40, size=(B, A))
    c = np.take_along_axis(centers, which[..., None].repeat(2, -1), axis=1) + rng.normal(0, 4, size=(B, A, 2))
    wh = rng.uniform(50, 70, size=(B, A, 2))
    yxyx = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)
    scores = orc.dedup_scores(rng.uniform(0.02, 1, size=(B, A, C)).astype(np.float32))
    out = nms_check(lib, yxyx, scores, 200, 200, 0.45, 0.01)
    assert (out[..., 1] > 0).sum(axis=1).max() <= 4 * 40 + 40  # ~ one survivor per (cluster, class)


@pytest.mark.parametrize("batch", [40, 200, 460])
def test_nms_merged_pipeline_cta_widths(lib, batch):
    """The merged pipeline picks 1024-, 512- or 256-thread CTAs from the batch size (up to one, up to three, more than
    three images per SM).  Half of the images hold clustered boxes (several chunks per image: kept-box tests, chunk
    sizing from the keep rate), a few hold boxes with swapped / equal corners (output rows rebuilt from the
    normalised box); pixel and unit-square coordinates (class shift of the prefilter).  C oracle on every image."""
    from oracle import c_oracle
    rng = np.random.default_rng(batch)
    A, C = 1500, 5
    c = rng.uniform(60, 580, size=(batch, A, 2))
    centers = rng.uniform(80, 560, size=(batch, 12, 2))
    which = rng.integers(0, 12, size=(batch, A))
    clustered = np.take_along_axis(centers, which[..., None].repeat(2, -1), axis=1) + rng.normal(0, 5, size=(batch, A, 2))
    c[::2] = clustered[::2]
    wh = rng.uniform(30, 120, size=(batch, A, 2))
    yxyx = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)
    yxyx[1::7, ::3] = yxyx[1::7, ::3][..., [2, 1, 0, 3]]      # y corners swapped
    yxyx[3::7, ::5] = yxyx[3::7, ::5][..., [0, 3, 2, 1]]      # x corners swapped
    yxyx[5::7, ::4, 2] = yxyx[5::7, ::4, 0]                   # no height
    yxyx[1::4] /= 640.0                                       # unit-square coordinates
    scores = orc.dedup_scores(rng.uniform(0.0, 1, size=(batch, A, C)).astype(np.float32))
    layer = lib.layers.NonMaximumSuppression(120, 120, 0.45, 0.05, False)
    out, valid, aidx = layer.call(yxyx, scores, return_aux=True)
    ref, vref, aref = c_oracle.combined_nms(yxyx, scores, 120, 120, 0.45, 0.05, nthreads=os.cpu_count() or 1)
    assert np.array_equal(valid, vref) and np.array_equal(aidx, aref) and np.array_equal(out, ref)


@pytest.mark.parametrize("C,mpc,mtot", [(4, 200, 200), (4, 400, 300), (4, 20, 50), (7, 100, 100)])
def test_nms_on_logits_equals_softmax_then_nms(lib, golden, C, mpc, mtot):
    """soft-max fused into the NMS score filter.  The soft-max arithmetic (hardware exponential) is
    checked against the oracle's at the path's fp32 tolerance (1e-5 relative); the fused call must equal NMS on the materialised probabilities bit for bit
    (4 classes + merged pipeline: probabilities only in registers; other shapes: materialised in the workspace), and
    that NMS is itself checked against the oracle on the same probabilities."""
    a = golden("anchors")["std_centroids"]
    rng = np.random.default_rng(40 + C + mpc)
    off, _ = orc.synth_predictions(rng, 3, 9600, 4)
    yxyx = orc.decode_batched_yxyx(off, a, STD)
    logits = rng.normal(0, 2.5, size=(3, 9600, C)).astype(np.float32)
    logits[0, :500] = 0.0                                  # equal logits -> equal probabilities (ties by class, anchor)
    logits[1, ::9] += 12.0                                 # saturated rows
    layer = lib.layers.NonMaximumSuppression(mpc, mtot, 0.45, 0.05, False)
    probs = layer.softmax_probabilities(logits)
    np.testing.assert_allclose(probs, orc.softmax_f32(logits), rtol=RTOL, atol=1e-30)
    assert np.abs(probs.sum(-1) - 1).max() < 1e-5
    scale = np.array([1.0] + [0.0 if c == 2 else 1.0 for c in range(1, C)], np.float32)
    for cs in (None, scale):
        got = layer.call_on_logits(yxyx, logits, return_aux=True, class_scale=cs)
        ref = layer.call(yxyx, probs, return_aux=True, class_scale=cs)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)
    # and the chain to the oracle: NMS of those probabilities
    nms_check(lib, yxyx, probs, mpc, mtot, 0.45, 0.05)


def test_nms_randomized_differential(lib, nms_path):
    """250 random (shape, class count, caps, thresholds, score distribution, box layout) cases against the C oracle
    on whichever NMS path the fixture forces (tools/fuzz_nms.py; 6,600 cases were run when the kernels were written)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "fuzz_nms", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_nms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    seed = {"fast": 11, "perclass": 12, "pcta": 13, "generic": 14, "blocked": 15, "fast_256t": 16, "fast_512t": 17}[nms_path]
    assert mod.run(250 if nms_path.startswith("fast") or nms_path == "blocked" else 120, seed, verbose=False) == 0


@pytest.mark.parametrize("C,mpc,mtot", [(4, 200, 200), (4, 4, 10), (6, 50, 50)])
def test_inference_tail_offsets_and_logits_in_one_call(lib, golden, C, mpc, mtot):
    """decode + soft-max + NMS as one call (boxes decoded per candidate, probabilities inside the score filter) equals
    the three separate layers bit for bit, on the fused path (4 classes, merged pipeline) and on the others."""
    a = golden("anchors")["std_centroids"]
    rng = np.random.default_rng(60 + C + mpc)
    off, _ = orc.synth_predictions(rng, 3, 9600, 4)
    logits = rng.normal(0, 2.5, size=(3, 9600, C)).astype(np.float32)
    dec = lib.layers.DecodeBoxesCentroidsOffsets(a[:, 0], a[:, 1], a[:, 2], a[:, 3], *STD)
    layer = lib.layers.NonMaximumSuppression(mpc, mtot, 0.45, 0.05, False)
    scale = np.array([1.0, 0.0] + [1.0] * (C - 2), np.float32)
    for cs in (None, scale):
        got = layer.call_on_offsets_and_logits(dec, off, logits, return_aux=True, class_scale=cs)
        ref = layer.call(dec(off), layer.softmax_probabilities(logits), return_aux=True, class_scale=cs)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)


def test_nms_two_phase_on_a_side_stream(lib, golden):
    """select (scores only, on a side stream) + finish == call, on the merged and on the per-class pipeline."""
    import torch

    g = golden("layers")
    boxes = torch.from_numpy(g["decoded_yxyx"]).cuda()
    scores = torch.from_numpy(g["scores"]).cuda()
    side = torch.cuda.Stream()
    for mpc, mtot, iou, thr in ((200, 200, 0.45, 0.01), (4, 10, 0.025, 0.2)):
        layer = lib.layers.NonMaximumSuppression(mpc, mtot, iou, thr, False)
        ref, vref, aref = layer.call(boxes, scores, return_aux=True)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            sel = layer.select(scores)
        out, valid, aidx = layer.finish(sel, boxes, return_aux=True)
        assert torch.equal(out, ref) and torch.equal(valid, vref) and torch.equal(aidx, aref)
    assert np.array_equal(ref.cpu().numpy().shape, (2, 10, 6))


def test_nms_on_offsets_equals_decode_then_nms(lib, golden):
    """ssdx_decode_nms (boxes decoded per candidate inside the NMS kernels) == DecodeBoxesCentroidsOffsets ->
    NonMaximumSuppression, bit for bit, on every NMS path (merged / per-class / generic via the fixture) and with a
    class scale; also checked against the oracle."""
    g = golden("layers")
    a = golden("anchors")["std_centroids"]
    dec = lib.layers.DecodeBoxesCentroidsOffsets(a[:, 0], a[:, 1], a[:, 2], a[:, 3], 0.1, 0.1, 0.2, 0.2)
    off, scores = g["offsets"], g["scores"]
    yxyx = dec(off)
    for mpc, mtot, iou, thr, scale in ((200, 200, 0.45, 0.01, None), (4, 10, 0.025, 0.2, None),
                                       (200, 200, 0.45, 0.01, [1, 0, 1, 1])):
        layer = lib.layers.NonMaximumSuppression(mpc, mtot, iou, thr, False)
        sc = None if scale is None else np.asarray(scale, np.float32)
        ref, vref, aref = layer.call(yxyx, scores, return_aux=True, class_scale=sc)
        out, valid, aidx = layer.call_on_offsets(dec, off, scores, return_aux=True, class_scale=sc)
        assert np.array_equal(out, ref) and np.array_equal(valid, vref) and np.array_equal(aidx, aref)
    o_ref, _, a_ref = orc.combined_nms(yxyx, scores, 200, 200, 0.45, 0.01)
    out, _, aidx = lib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False).call_on_offsets(dec, off, scores, True)
    assert np.array_equal(aidx, a_ref) and np.array_equal(out, o_ref)


def test_nms_torch_in_torch_out(lib, golden):
    import torch

    g = golden("layers")
    boxes = torch.from_numpy(g["decoded_yxyx"]).cuda()
    scores = torch.from_numpy(g["scores"]).cuda()
    out = lib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)(boxes, scores)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert np.array_equal(out.cpu().numpy(), g["det_200"])


def test_nms_rejects_bad_iou_threshold(lib, golden):
    g = golden("layers")
    with pytest.raises(ValueError):
        lib.layers.NonMaximumSuppression(10, 10, 1.5, 0.1, False)(g["decoded_yxyx"], g["scores"])


# ------------------------------------------------------------------------------------------- full bench size
def test_round_trip_at_bench_size_against_c_oracle(lib, coder, std_anchors, golden):
    """BASELINE's full configuration (batch 1024, A = 9600, G = 16, C = 4, thr .01, IoU .45, top-200): every image of
    encode -> decode -> NMS against the C/OpenMP port of the oracle (itself checked against the NumPy oracle in
    tests/test_c_oracle.py).  Matched indices, labels, keep-sets and detection rows bit-exact; offsets/boxes 1e-5."""
    import os

    from oracle import c_oracle

    NT = os.cpu_count() or 1
    B = 1024
    rng = np.random.default_rng(1995)
    labels, boxes, counts = orc.synth_gt(rng, B, 16)
    counts[::97] = rng.integers(0, 16, size=counts[::97].shape)           # a few ragged / empty images
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = c_oracle.encode_batch(std_anchors, labels, boxes, counts, 4, 0.5, STD, nthreads=NT)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)
    cent = golden("anchors")["std_centroids"]
    layer = lib.layers.DecodeBoxesCentroidsOffsets(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], *STD)
    yxyx = layer(O)
    assert_coords_close(yxyx, c_oracle.decode_batch(cent, O, STD, 0, nthreads=NT), cent[None, :, 2:])
    noise = rng.normal(0, 1, size=L.shape).astype(np.float32)
    scores = (np.float32(0.9) * L + np.float32(0.1) * orc.softmax_f32(noise)).astype(np.float32)
    nms = lib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
    det, valid, aidx = nms.call(yxyx, scores, return_aux=True)
    det_r, valid_r, aidx_r = c_oracle.combined_nms(yxyx, scores, 200, 200, 0.45, 0.01, nthreads=NT)
    assert np.array_equal(aidx, aidx_r) and np.array_equal(valid, valid_r) and np.array_equal(det, det_r)
    assert np.all(np.diff(det[..., 1], axis=1) <= 0)                      # rows sorted by score


# ------------------------------------------------------------------------------------------- segmentation suppression
def test_segmentation_suppression_golden(lib, golden):
    g = golden("seg")
    layer = lib.layers.SegmentationSuppression()
    for case in "abcd":
        pres = lib.layers.SegmentationSuppression.class_presence(g[case + "_mask"])
        assert pres.dtype == np.float32 and np.array_equal(pres, orc.seg_class_presence(g[case + "_mask"]))
        out = layer(g[case + "_mask"], g[case + "_probs"])
        assert out.shape == g[case + "_probs"].shape and np.array_equal(out, g[case + "_suppressed"])


@pytest.mark.parametrize("cm,shape", [(4, (3, 120, 160)), (4, (1, 7, 5)), (3, (2, 33, 17)), (6, (2, 40, 24))])
def test_segmentation_presence_shapes_and_rare_classes(lib, cm, shape):
    """A class that wins a single pixel of the whole batch is present; classes >= 4 fall outside one_hot(depth=4)."""
    rng = np.random.default_rng(cm * 100 + shape[1])
    mask = rng.uniform(0.0, 0.5, size=shape + (cm,)).astype(np.float32)
    mask[..., 0] += 1.0                       # background wins everywhere ...
    mask[-1, shape[1] // 2, shape[2] // 3, min(2, cm - 1)] = 5.0   # ... except one pixel
    if cm > 4:
        mask[0, 0, 0, 5] = 9.0                # index 5: dropped by one_hot(depth=4)
    pres = lib.layers.SegmentationSuppression.class_presence(mask)
    assert np.array_equal(pres, orc.seg_class_presence(mask))
    assert pres[0] == 1 and pres[min(2, cm - 1)] == 1 and pres.sum() == 2


def test_nms_class_scale_equals_nms_of_suppressed_scores(lib, golden):
    """The fused path (scores multiplied as they are read) against the NMS of the materialised product, on the merged
    pipeline (200/200), the per-class pipeline (4/10) and with a non-0/1 scale."""
    g = golden("layers")
    yxyx, scores = g["decoded_yxyx"], g["scores"]
    for scale in ([1, 0, 1, 1], [0, 1, 0, 1], [1, 1, 1, 1], [0.5, 0.25, 2.0, 1.0]):
        sc = np.asarray(scale, np.float32)
        sup = (scores * sc).astype(np.float32)
        for mpc, mtot, iou, thr in ((200, 200, 0.45, 0.01), (4, 10, 0.025, 0.2)):
            layer = lib.layers.NonMaximumSuppression(mpc, mtot, iou, thr, False)
            fused, v1, a1 = layer.call(yxyx, scores, return_aux=True, class_scale=sc)
            ref, v2, a2 = orc.combined_nms(yxyx, sup, mpc, mtot, iou, thr)
            assert np.array_equal(a1, a2) and np.array_equal(v1, v2) and np.array_equal(fused, ref)


# ------------------------------------------------------------------------------------------- training-time box metric
def test_box_iou_metric_golden_and_full_batch(lib, golden, std_anchors, coder):
    """metrics.jaccard_iou_bounding_boxes: the fused kernel against the fixture made by the reference's own file, and
    at batch 1024 against the oracle (fp32 sum order differs: 1e-5 relative, the north star's fp32 tolerance)."""
    g = golden("metric")
    cent = golden("anchors")["std_centroids"]
    metric = lib.metrics.jaccard_iou_bounding_boxes(cent[:, 0], cent[:, 1], cent[:, 2], cent[:, 3], STD)
    got = metric(g["y_true"], g["y_pred"])
    assert got.shape == (6,) and got.dtype == np.float32
    np.testing.assert_allclose(got, g["value"], rtol=RTOL)
    rng = np.random.default_rng(8)
    labels, boxes, counts = orc.synth_gt(rng, 1024, 16)
    counts[5] = 0                                                        # no positives -> 0 / 0 = NaN, as the reference
    _, y_true = coder.encode_batch(boxes, labels, counts)
    y_pred = (y_true + rng.normal(0, 0.5, y_true.shape)).astype(np.float32)
    got = metric(y_true, y_pred)
    ref = orc.metric_iou_boxes(y_true, y_pred, cent, STD)
    assert np.isnan(got[5]) and np.isnan(ref[5])
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=RTOL)
    with pytest.raises(ValueError):
        metric(y_true[:, :100], y_pred[:, :100])


# ------------------------------------------------------------------------------------------- training losses
def assert_kept_equal(keep, keep_ref, v_ref):
    """The kept background set is index work: equal to the oracle's except, possibly, anchors whose loss lies within a
    few ulp of the selection threshold (logf on the GPU vs NumPy's log differ by <= 2 ulp)."""
    diff = keep != keep_ref
    if not diff.any():
        return
    thr = v_ref[keep_ref == 1].min()
    assert np.all(np.abs(v_ref[diff] - thr) <= 8 * np.spacing(np.float32(thr))), int(diff.sum())
    assert int(keep.sum()) == int(keep_ref.sum())


def test_localization_loss_golden_and_full_batch(lib, golden, coder):
    g = golden("losses")
    got = lib.losses.localization_loss(g["loc_y_true"], g["loc_y_pred"])
    assert got.shape == (6,) and got.dtype == np.float32
    np.testing.assert_allclose(got, g["loc_value"], rtol=RTOL)
    assert got[2] == 0 and got[4] == 0
    rng = np.random.default_rng(18)
    labels, boxes, counts = orc.synth_gt(rng, 300, 16)
    counts[7] = 0
    _, y_true = coder.encode_batch(boxes, labels, counts)
    y_pred = (y_true + rng.normal(0, 0.9, y_true.shape)).astype(np.float32)
    loss, grad = lib.losses.localization_loss(y_true, y_pred, with_gradient=True)
    ref, gref = orc.localization_loss(y_true, y_pred, with_gradient=True)
    np.testing.assert_allclose(loss, ref, rtol=RTOL)
    np.testing.assert_allclose(grad, gref, rtol=RTOL, atol=1e-12)
    assert np.array_equal(grad != 0, gref != 0) and loss[7] == 0
    with pytest.raises(ValueError):
        lib.losses.localization_loss(y_true[:, :, :3], y_pred[:, :, :3])


def test_confidence_loss_golden(lib, golden):
    g = golden("losses")
    got, keep = lib.losses.confidence_loss(g["conf_y_true"], g["conf_y_pred"], return_kept_background=True)
    assert got.shape == (6,) and got.dtype == np.float32
    np.testing.assert_allclose(got, g["conf_value"], rtol=RTOL)
    _, keep_ref, v_ref = orc.confidence_loss(g["conf_y_true"], g["conf_y_pred"], return_keep=True)
    assert_kept_equal(keep, keep_ref, v_ref)
    got = lib.losses.confidence_loss(g["conf_y_true"][2:3], g["conf_y_pred"][2:3])      # no positive: k = 0
    assert np.array_equal(got, g["conf_value_no_positive"])
    got = lib.losses.confidence_loss(g["conf_few_y_true"], g["conf_few_y_pred"])        # k = all negatives
    np.testing.assert_allclose(got, g["conf_few_value"], rtol=RTOL)


@pytest.mark.parametrize("B,A,C", [(64, 9600, 4), (5, 1001, 3), (3, 777, 21)])
def test_confidence_loss_random_batches_with_gradient(lib, B, A, C):
    rng = np.random.default_rng(B + C)
    cls = rng.integers(1, C, (B, A)) * (rng.uniform(size=(B, A)) < 0.02)
    y_true = np.zeros((B, A, C), np.float32)
    np.put_along_axis(y_true, cls[..., None], 1.0, axis=-1)
    y_pred = orc.softmax_f32((rng.normal(0, 2, (B, A, C)) + 1.5 * y_true).astype(np.float32))
    y_pred[0, :50] = 0
    y_pred[0, :50, 0] = 1                                    # saturated: both clip bounds, zero gradient there
    loss, grad, keep = lib.losses.confidence_loss(y_true, y_pred, with_gradient=True, return_kept_background=True)
    ref, gref, keep_ref, v_ref = orc.confidence_loss(y_true, y_pred, with_gradient=True, return_keep=True)
    np.testing.assert_allclose(loss, ref, rtol=RTOL)
    assert_kept_equal(keep, keep_ref, v_ref)
    same = (keep == keep_ref)[..., None] & np.ones(C, bool)
    np.testing.assert_allclose(grad[same], gref[same], rtol=RTOL, atol=1e-12)
    assert int(keep.sum()) == min(3 * int((cls > 0).sum()), int((cls == 0).sum()))


def test_confidence_loss_threshold_ties_are_split_in_index_order(lib):
    """Thousands of background boxes with the SAME loss straddle the threshold (saturated predictions): the tied ones
    are kept lower flat index first (tf.math.top_k), across image boundaries; here the kept set is bit-exact."""
    B, A, C = 6, 2000, 4
    y_true = np.zeros((B, A, C), np.float32)
    y_true[..., 0] = 1
    y_true[:, 10:40] = np.array([0, 0, 1, 0], np.float32)          # 30 positives per image -> k = 540
    y_pred = np.full((B, A, C), 0.25, np.float32)                   # every background loss = log 4
    y_pred[:, 100:110] = np.array([0.125, 0.5, 0.25, 0.125], np.float32)   # 10 per image clearly above
    loss, keep = lib.losses.confidence_loss(y_true, y_pred, return_kept_background=True)
    ref, keep_ref, _ = orc.confidence_loss(y_true, y_pred, return_keep=True)
    assert np.array_equal(keep, keep_ref)
    assert keep.sum(axis=1).tolist() == keep_ref.sum(axis=1).tolist() and keep[0].sum() == 490 and keep[2].sum() == 10
    np.testing.assert_allclose(loss, ref, rtol=RTOL)
    # no background box at all: the background part is zero (losses.py:123-124)
    y_true[..., 0], y_true[..., 1] = 0, 1
    got = lib.losses.confidence_loss(y_true, y_pred)
    np.testing.assert_allclose(got, orc.confidence_loss(y_true, y_pred), rtol=RTOL)


def test_losses_randomized_differential(lib):
    """120 random cases (shapes, class counts, positive fractions 0..1, soft-max / saturated / constant / quantised
    predictions: massive ties at the hard-negative threshold) of both losses, their gradients and the kept-background
    mask against the NumPy oracle (tools/fuzz_losses.py)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "fuzz_losses", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(120, 31, verbose=False) == 0


def test_losses_on_device_tensors(lib, golden):
    import torch

    g = golden("losses")
    yt, yp = torch.from_numpy(g["conf_y_true"]).cuda(), torch.from_numpy(g["conf_y_pred"]).cuda()
    out = lib.losses.confidence_loss(yt, yp)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    np.testing.assert_allclose(out.cpu().numpy(), g["conf_value"], rtol=RTOL)
    lt, lp = torch.from_numpy(g["loc_y_true"]).cuda(), torch.from_numpy(g["loc_y_pred"]).cuda()
    out, grad = lib.losses.localization_loss(lt, lp, with_gradient=True)
    assert out.is_cuda and grad.shape == lt.shape
    np.testing.assert_allclose(out.cpu().numpy(), g["loc_value"], rtol=RTOL)


# ------------------------------------------------------------------------------------------- eval
def test_eval_golden(lib, golden, tmp_path):
    g = golden("eval")
    N = g["pred_labels"].shape[0]
    no_gt = int(g["no_gt_image"][0])
    paths = []
    for i in range(N):
        p = tmp_path / f"gt_{i}.csv"
        n = 0 if i == no_gt else g["gt_labels"].shape[1]
        p.write_text("".join(f"{int(l)},{float(b[0])!r},{float(b[1])!r},{float(b[2])!r},{float(b[3])!r}\n"
                             for l, b in zip(g["gt_labels"][i, :n], g["gt_boxes"][i, :n])))
        paths.append(str(p))
    for i in range(N):
        iou = lib.evaluators._iou_boxes_pred_vs_true(g["pred_labels"][i], g["pred_boxes"][i], g["gt_labels"][i],
                                                     g["gt_boxes"][i])
        assert np.array_equal(iou, g["iou"][i])                          # IoU matrix bit-exact
    z = lib.evaluators._iou_boxes_pred_vs_true(g["pred_labels"][0], g["pred_boxes"][0], [], np.zeros((0, 4)))
    assert np.array_equal(z, g["iou_nogt"])
    for thr, key in ((0.5, "ap50"), (0.6, "ap60"), (0.7, "ap70")):
        ap = lib.evaluators.average_precision_object_detection(g["pred_labels"], g["pred_conf"], g["pred_boxes"], thr,
                                                               paths, [0, 1, 2, 3], 0)
        np.testing.assert_allclose([ap[1], ap[2], ap[3]], g[key], rtol=0, atol=1e-12)


def test_eval_tp_flags_and_counts(lib):
    rng = np.random.default_rng(1997)
    N, K, G, C = 64, 200, 16, 4
    gl, gb, gc = orc.synth_gt(rng, N, G, integer=True)
    gc = rng.integers(0, G + 1, size=N).astype(np.int32)
    pl = rng.integers(0, C, size=(N, K)).astype(np.int32)
    pick = rng.integers(0, G, size=(N, K))
    pb = (np.take_along_axis(gb, pick[..., None].repeat(4, -1), axis=1) + rng.normal(0, 10, (N, K, 4))).astype(np.float32)
    for thr in (0.5, 0.6, 0.7):
        tp, counts = lib.evaluators.true_positives_and_counts(pl, pb, gl, gb, gc, thr, C, 0)
        ref = np.stack([orc.eval_tp_flags(pl[i], pb[i], gl[i, :gc[i]], gb[i, :gc[i]], thr, 0) for i in range(N)])
        assert np.array_equal(tp, ref)
        for c in range(C):
            sc = (pl == c) & (pl != 0)
            assert counts[c, 0] == ref[sc].sum() and counts[c, 1] == sc.sum() - ref[sc].sum()
            assert counts[c, 2] == sum((gl[i, :gc[i]] == c).sum() for i in range(N))


# ------------------------------------------------------------------------------------------- determinism
def test_repeated_runs_are_bit_identical(lib, coder, std_anchors, golden):
    """Races would show up as run-to-run differences: the fused encode patches positives over TMA bulk stores, the
    NMS appends candidates with atomics in arbitrary order, the tiled encode publishes arg-max results across CTAs.
    40 repetitions of each at the bench batch size must give the same bits."""
    import torch

    rng = np.random.default_rng(7)
    labels, boxes, counts = orc.synth_gt(rng, 1024, 16)
    gb, gl, gc = (torch.from_numpy(x).cuda() for x in (boxes, labels, counts))
    L0, O0 = coder.encode_batch(gb, gl, gc)
    scores = (0.9 * L0 + 0.1 * torch.softmax(torch.randn(L0.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(3)), -1))
    a = golden("anchors")["std_centroids"]
    dec = lib.layers.DecodeBoxesCentroidsOffsets(a[:, 0], a[:, 1], a[:, 2], a[:, 3], *STD)
    nms = lib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
    yx0 = dec(O0)
    D0 = nms(yx0, scores)
    for _ in range(40):
        L, O = coder.encode_batch(gb, gl, gc)
        assert torch.equal(L, L0) and torch.equal(O, O0)
        assert torch.equal(nms(dec(O), scores), D0)
        assert torch.equal(nms.call_on_offsets(dec, O, scores), D0)


# ------------------------------------------------------------------------------------------- DLPack
class _DLPackOnly:
    """A device array that is NOT a torch tensor: only the DLPack protocol (what a TF / CuPy / JAX array offers)."""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__() if stream is None else self._t.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_dlpack_device_arrays_are_used_in_place(lib, coder, golden, std_anchors):
    """north star: 'DLPack for zero-copy device tensors'.  Any object with __dlpack__ is consumed on the device it
    lives on (no host round trip) and device tensors come back."""
    import torch

    g = golden("layers")
    a = golden("anchors")["std_centroids"]
    dec = lib.layers.DecodeBoxesCentroidsOffsets(a[:, 0], a[:, 1], a[:, 2], a[:, 3], 0.1, 0.1, 0.2, 0.2)
    off_t = torch.from_numpy(g["offsets"]).cuda()
    sc_t = torch.from_numpy(g["scores"]).cuda()
    yxyx = dec(_DLPackOnly(off_t))
    assert isinstance(yxyx, torch.Tensor) and yxyx.is_cuda
    assert np.array_equal(yxyx.cpu().numpy(), dec(g["offsets"]))
    box_t = torch.from_numpy(g["decoded_yxyx"]).cuda()  # the fixture's boxes: det_200 holds copies of their bits
    det = lib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)(_DLPackOnly(box_t), _DLPackOnly(sc_t))
    assert det.is_cuda and np.array_equal(det.cpu().numpy(), g["det_200"])
    rng = np.random.default_rng(5)
    labels, boxes, counts = orc.synth_gt(rng, 4, 16)
    L, O = coder.encode_batch(_DLPackOnly(torch.from_numpy(boxes).cuda()), _DLPackOnly(torch.from_numpy(labels).cuda()),
                              _DLPackOnly(torch.from_numpy(counts).cuda()))
    Lr, Or, _ = encode_ref(std_anchors, labels, boxes, counts, 4, 0.5)
    assert L.is_cuda and np.array_equal(L.cpu().numpy(), Lr)
    assert_offsets_close(O.cpu().numpy(), Or)


# ------------------------------------------------------------------------------------------- average precision
def test_average_precision_on_device(lib):
    """ssdx_average_precision (radix sort + scan + trapezoid on the GPU) against the reference's NumPy formula
    (evaluators.py:159-187 as restated in average_precision_from_flags) on tie-free confidences, including a class
    without predictions, a class without ground truth and segments longer than one scan chunk."""
    import torch

    rng = np.random.default_rng(17)
    n, C, bg = 50_000, 6, 0
    label = rng.integers(0, C, size=n).astype(np.int32)
    label[label == 4] = 1                                   # class 4: no predictions
    conf = rng.permutation(n).astype(np.float32) / np.float32(n)   # distinct
    tp = (rng.uniform(size=n) < 0.3).astype(np.uint8)
    tp[label == bg] = 0
    n_gt = {1: 4000, 2: 700, 3: 0, 4: 50, 5: 123}                  # class 3: no ground truth
    counts = np.zeros((C, 3), np.int64)
    for c, v in n_gt.items():
        counts[c, 2] = v
    ref = lib.evaluators.average_precision_from_flags(tp[label != bg], conf[label != bg], label[label != bg], n_gt, bg)
    sx = lib._ssdx
    got = sx.average_precision(torch.from_numpy(tp).cuda(), torch.from_numpy(conf).cuda(), torch.from_numpy(label).cuda(),
                               torch.from_numpy(counts).cuda(), C, bg).cpu().numpy()
    assert got.dtype == np.float64 and got[bg] == 0.0 and got[3] == 0.0 and got[4] == 0.0
    for c in (1, 2, 5):
        assert abs(got[c] - ref[c]) <= 1e-12 * max(1.0, abs(ref[c])), (c, got[c], ref[c])
    empty = sx.average_precision(torch.zeros(0, dtype=torch.uint8, device="cuda"), torch.zeros(0, device="cuda"),
                                 torch.zeros(0, dtype=torch.int32, device="cuda"), torch.from_numpy(counts).cuda(), C, bg)
    assert not empty.cpu().numpy().any()


# ------------------------------------------------------------------------------------------- C ABI details
def test_error_paths_and_workspace(lib, coder):
    import ctypes as C

    import torch

    s = lib._ssdx
    assert s.version() == 100
    with pytest.raises(ValueError):
        s.Anchors(np.zeros((0, 4), np.float32), None)
    gt = torch.zeros((2, 4, 4), device="cuda")
    lab = torch.zeros((2, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros((2,), dtype=torch.int32, device="cuda")
    out_l = torch.empty((2, 9600, 4), device="cuda")
    out_o = torch.empty((2, 9600, 4), device="cuda")
    st = s.lib.ssdx_encode(coder._anchors.handle, gt.data_ptr(), lab.data_ptr(), cnt.data_ptr(), 2, 4, 4, 0.5,
                           s.std_array(STD), out_l.data_ptr(), out_o.data_ptr(), None, None, 0, None)
    assert st == s.SSDX_ERR_WORKSPACE and b"workspace" in s.lib.ssdx_last_error()
    assert s.lib.ssdx_encode_workspace_bytes(9600, 2, 4) >= 2 * 9600 * 4 + 2 * 4 * 8
    n = C.c_int(0)
    assert s.lib.ssdx_device_count(C.byref(n)) == 0 and n.value >= 1


# ------------------------------------------------------------------------------------------- BASELINE configs, full batch
def test_encode_stress_config_full_batch(lib, encode_path):
    """BASELINE config 3 (dense stress: 100,000 anchors x 256 GT boxes, batch 256) - EVERY image against the C oracle:
    matched indices and labels bit-exact, offsets within 1e-5 relative.  (The 25.6 M-pair IoU matrix per image is
    what the oracle computes; the device evaluates ~72 k window cells per image on the grid drivers.)"""
    if encode_path == "tile_fused":
        pytest.skip("100k anchors do not fit one CTA per image: same launch as tile_tiled")
    import os

    from oracle import c_oracle

    anchors = orc.anchors_generate(**orc.STRESS_ANCHOR_CFG)
    rng = np.random.default_rng(1996)
    B, G = 256, 256
    labels, boxes, counts = orc.synth_gt(rng, B, G, image_hw=(1000, 1280))
    c = lib.datacoder.DataEncoderDecoder(4, (1000, 1280), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
    import torch

    gb, gl, gc = (torch.from_numpy(x).cuda() for x in (boxes, labels, counts))
    for lo in range(0, B, 64):  # 64 images at a time: 64 x 100k x (4 + 4 + 1) floats per chunk on either side
        L, O, M = c.encode_batch(gb[lo:lo + 64], gl[lo:lo + 64], gc[lo:lo + 64], return_match=True)
        Lr, Or, Mr = c_oracle.encode_batch(anchors, labels[lo:lo + 64], boxes[lo:lo + 64], counts[lo:lo + 64], 4, 0.5,
                                           STD, nthreads=os.cpu_count() or 1)
        assert np.array_equal(M.cpu().numpy(), Mr), f"matched indices differ in images {lo}..{lo + 63}"
        assert np.array_equal(L.cpu().numpy(), Lr)
        assert_offsets_close(O.cpu().numpy(), Or)


def test_eval_config_full_batch(lib):
    """BASELINE config 5 inputs (4096 images x 200 detections, 16 GT boxes): TP flags and per-class counts of ALL images
    against the C oracle, AP at 0.5 / 0.6 / 0.7 (03-...ipynb cell 29) against the NumPy restatement of
    evaluators.py:159-187."""
    import torch

    from oracle import c_oracle

    B, K, G, C = 4096, 200, 16, 4
    pl, conf, pb, gl, gb, gc = orc.synth_eval(np.random.default_rng(1997), B, K, G, C)
    d = [torch.from_numpy(x).cuda() for x in (pl, pb, gb, gl, gc, conf)]
    for thr in (0.5, 0.6, 0.7):
        tp, counts = lib._ssdx.eval_tp(d[0], d[1], d[2], d[3], d[4], C, 0, thr)
        ref_tp, ref_counts = c_oracle.eval_tp(pl, pb, gb, gl, gc, C, 0, thr)
        assert np.array_equal(tp.cpu().numpy(), ref_tp) and np.array_equal(counts.cpu().numpy(), ref_counts)
        ap = lib._ssdx.average_precision(tp.reshape(-1), d[5].reshape(-1), d[0].reshape(-1), counts, C, 0).cpu().numpy()
        ref_ap = orc.average_precision_from_tp(ref_tp, conf, pl, {c: int(ref_counts[c, 2]) for c in range(1, C)})
        for c in range(1, C):
            assert abs(ap[c] - ref_ap[c]) < 1e-9, (thr, c, ap[c], ref_ap[c])


def test_encode_decode_centroid_constructor_and_feature_map_inputs(lib, std_anchors):
    """DataEncoderDecoder built from centroids (datacoder.py:73-89: the corners are then DERIVED, a different fp32
    table) and from the getters' 'feature-maps' style outputs flattened as notebook 03 cell 6 does."""
    dbb = lib.boxes.DefaultBoundingBoxes(
        feature_maps_shapes=((30, 40), (15, 20), (8, 10), (4, 5)),
        centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1), boxes_scales=(0.15, 0.95),
        additional_square_box=True)
    dbb.rescale_boxes_coordinates((480, 640))
    cx, cy, w, h = dbb.get_boxes_coordinates_centroids("ssd").T
    coder = lib.datacoder.DataEncoderDecoder(4, (480, 640), center_x_boxes_default=cx, center_y_boxes_default=cy,
                                             width_boxes_default=w, height_boxes_default=h, iou_threshold=0.5)
    derived = np.stack(orc.centroids_to_corners(cx, cy, w, h), 1).astype(np.float32)  # datacoder.py:84-89
    assert not np.array_equal(derived, std_anchors)  # a few coordinates move by an ulp: not the corner-built table
    assert np.array_equal(np.stack([coder.xmin_boxes_default, coder.ymin_boxes_default, coder.xmax_boxes_default,
                                    coder.ymax_boxes_default], 1), derived)
    rng = np.random.default_rng(8)
    labels, boxes, counts = orc.synth_gt(rng, 160, 16, integer=True)
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(derived, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)
    cent = np.stack([cx, cy, w, h], 1).astype(np.float32)  # decode uses the centroids as passed (datacoder.py:78-81)
    assert_coords_close(coder.decode_to_corners(O[5]), orc.decode_to_corners(Or[5], cent, STD), cent[:, 2:])
    assert_coords_close(coder.decode_to_centroids(O[5]), orc.decode_to_centroids(Or[5], cent, STD), cent[:, 2:])
    # 'feature-maps' style getters: per-map (H, W, nb, 4) arrays, flattened map-major -> the same 'ssd' table
    fm = dbb.get_boxes_coordinates_corners("feature-maps")
    flat = np.concatenate([m.reshape(-1, 4) for m in fm], axis=0)
    assert np.array_equal(flat, std_anchors)
    c2 = lib.datacoder.DataEncoderDecoder(4, (480, 640), flat[:, 0], flat[:, 1], flat[:, 2], flat[:, 3])
    L2, O2, M2 = c2.encode_batch(boxes[:8], labels[:8], counts[:8], return_match=True)
    Lr2, Or2, Mr2 = encode_ref(std_anchors, labels[:8], boxes[:8], counts[:8], 4, 0.5)
    assert np.array_equal(M2, Mr2) and np.array_equal(L2, Lr2)
    assert_offsets_close(O2, Or2)


def test_encode_many_positives_overflow_path(lib, coder, std_anchors, encode_path):
    """More positives per image than the fused grid driver's slot table holds (threshold 0.05: thousands of anchors
    above it) - the exact per-anchor loop takes over; also G = 256, the most GT boxes one CTA stages."""
    rng = np.random.default_rng(31)
    labels, boxes, counts = orc.synth_gt(rng, 150, 16)
    c = lib.datacoder.DataEncoderDecoder(4, (480, 640), std_anchors[:, 0], std_anchors[:, 1], std_anchors[:, 2],
                                         std_anchors[:, 3], iou_threshold=0.05)
    L, O, M = c.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, 4, 0.05)
    assert (Mr >= 0).sum(axis=1).max() > 1024
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)
    # 64 GT boxes at threshold 0.05: more pairs above the threshold than the split driver's hit lists hold
    labels, boxes, counts = orc.synth_gt(rng, 20, 64)
    L, O, M = c.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, 4, 0.05)
    iou = orc.iou_matrix(std_anchors, boxes[0])
    assert (iou > 0.05).sum() > 1.5 * 9600
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)
    labels, boxes, counts = orc.synth_gt(rng, 150, 256, wh_range=(10.0, 200.0))
    counts[::3] = rng.integers(0, 257, size=len(counts[::3]))
    L, O, M = coder.encode_batch(boxes, labels, counts, return_match=True)
    Lr, Or, Mr = encode_ref(std_anchors, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(M, Mr) and np.array_equal(L, Lr)
    assert_offsets_close(O, Or)


# ------------------------------------------------------------------------------------------- the collective (NCCL)
def test_all_gather_counts_over_nccl(lib, tmp_path):
    """ssdx_allgather_counts on a real NCCL communicator: one rank per visible GPU (at most 2), launched like the
    bench; with a single GPU the world has one rank (NCCL still runs the collective on the device)."""
    import os
    import subprocess
    import sys
    import textwrap

    import torch

    world = min(torch.cuda.device_count(), 2)
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(lib.__file__)))
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {pkg!r})
        import torch, torch.distributed as dist
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{{local}}"))
        import ssdseglib
        from ssdseglib import _dist, _ssdx
        r, w = dist.get_rank(), dist.get_world_size()
        ver = _ssdx.C.c_int(0)
        _ssdx.check(_ssdx.lib.ssdx_nccl_version(_ssdx.C.byref(ver)))
        assert ver.value >= 20000, ver.value
        lo, hi = _dist.shard_range(10, r, w)
        counts = torch.tensor([[r + 1, 10 * (r + 1), hi - lo]] * 4, dtype=torch.int64, device=f"cuda:{{local}}")
        for _ in range(3):  # the first call creates the communicator (collective), the next ones reuse it
            g = _dist.all_gather_counts(counts)
        assert g.is_cuda and g.shape == (w, 4, 3)
        tot = g.sum(0).cpu()
        assert tot[0].tolist() == [w * (w + 1) // 2, 10 * w * (w + 1) // 2, 10], tot
        assert g[r].cpu().tolist() == counts.cpu().tolist()
        # through the public evaluation entry point: counts summed over the ranks
        import numpy as np
        rng = np.random.default_rng(5 + r)
        pl = rng.integers(0, 4, size=(8, 20)).astype(np.int32)
        gb = rng.uniform(0, 300, size=(8, 6, 4)).astype(np.float32); gb[..., 2:] += gb[..., :2]
        pb = np.repeat(gb[:, :1], 20, axis=1) + rng.normal(0, 6, size=(8, 20, 4)).astype(np.float32)
        gl = rng.integers(1, 4, size=(8, 6)).astype(np.int32); gc = np.full(8, 6, np.int32)
        tp, c_all = ssdseglib.evaluators.true_positives_and_counts(pl, pb, gl, gb, gc, 0.5, 4, 0,
                                                                   process_group=dist.group.WORLD)
        tp1, c_one = ssdseglib.evaluators.true_positives_and_counts(pl, pb, gl, gb, gc, 0.5, 4, 0)
        mine = torch.from_numpy(np.asarray(c_one)).cuda()
        dist.all_reduce(mine)
        assert np.array_equal(np.asarray(c_all), mine.cpu().numpy()) and np.array_equal(tp, tp1)
        assert torch.cuda.current_device() == local
        # the grouped gather of the AP inputs against torch.distributed's all-gathers
        comm = _dist.CountsCommunicator(None, local)
        n = 1000
        tpf = torch.from_numpy((rng.uniform(size=n) < 0.3).astype(np.uint8)).cuda()
        cf = torch.from_numpy(rng.uniform(size=n).astype(np.float32)).cuda()
        lb = torch.from_numpy(rng.integers(0, 4, size=n).astype(np.int32)).cuda()
        gc_, gt_, gf_, gl_ = comm.all_gather_eval(counts, tpf, cf, lb)
        for mine, got in ((tpf, gt_), (cf, gf_), (lb, gl_), (counts.view(-1), gc_.view(-1))):
            ref = torch.empty(w * mine.numel(), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(ref, mine.contiguous())
            assert torch.equal(ref, got)
        comm.close()
        _dist.close_communicators()
        dist.barrier()
        if r == 0:
            print("NCCL_GATHER_OK world", w)
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and f"NCCL_GATHER_OK world {world}" in r.stdout, r.stdout + r.stderr


def test_entry_points_leave_the_current_device_alone(lib, std_anchors, golden):
    """A process that drives two GPUs from one thread: calls on device 1 must not change the thread's current device
    (every entry point saves and restores it)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    coder1 = lib.datacoder.DataEncoderDecoder(4, (480, 640), std_anchors[:, 0], std_anchors[:, 1], std_anchors[:, 2],
                                              std_anchors[:, 3], device=1)
    assert torch.cuda.current_device() == 0
    rng = np.random.default_rng(2)
    labels, boxes, counts = orc.synth_gt(rng, 4, 16)
    gb, gl, gc = (torch.from_numpy(x).to("cuda:1") for x in (boxes, labels, counts))
    L, O = coder1.encode_batch(gb, gl, gc)
    assert torch.cuda.current_device() == 0 and L.device.index == 1
    yx = coder1._decode(O, 0)
    assert torch.cuda.current_device() == 0
    nms = lib.layers.NonMaximumSuppression(200, 200, 0.45, 0.01, False)
    det = nms(yx, L)
    assert torch.cuda.current_device() == 0 and det.device.index == 1
    Lr, Or, _ = encode_ref(std_anchors, labels, boxes, counts, 4, 0.5)
    assert np.array_equal(L.cpu().numpy(), Lr)
