"""Pin the CPU oracle against fixtures produced by the reference's own source
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import ssd_oracle as orc

STD = (0.1, 0.1, 0.2, 0.2)


def test_anchor_generation_matches_reference(golden):
    g = golden("anchors")
    a = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
    assert a.dtype == np.float32 and a.shape == (9600, 4)
    assert np.array_equal(a, g["std_corners"])
    # survey-time known answers
    np.testing.assert_allclose(a[0], [-12.698076, -16.930172, 44.648075, 40.88017], rtol=1e-6)
    np.testing.assert_allclose(a[-1], [385.28595, 241.38498, 764.914, 620.81506], rtol=1e-6)
    tut = orc.anchors_generate(((30, 40), (15, 20), (8, 10), (4, 5)),
                               centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1),
                               additional_square_box=True, image_shape=(480, 640))
    assert np.array_equal(tut, g["tut_corners"])
    odd = orc.anchors_generate(((3, 1), (2, 2)), ((1, 2), (1, 0.5, 3)), (0.3, 0.8), 0.1, False, (33, 65))
    assert np.array_equal(odd, g["odd_once"])
    stress = orc.anchors_generate(**orc.STRESS_ANCHOR_CFG)
    assert tuple(g["stress_shape"]) == stress.shape == (100000, 4)
    assert np.array_equal(stress[g["stress_rows"]], g["stress_sel"])
    assert np.sum(stress, dtype=np.float64) == g["stress_sum"][0]


def test_centroid_conversions(golden):
    g = golden("anchors")
    a = g["std_corners"]
    c = np.stack(orc.corners_to_centroids(a[:, 0], a[:, 1], a[:, 2], a[:, 3]), 1)
    assert np.array_equal(c, g["conv_centroids"]) and np.array_equal(c, g["std_centroids"])
    k = np.stack(orc.centroids_to_corners(c[:, 0], c[:, 1], c[:, 2], c[:, 3]), 1)
    assert np.array_equal(k, g["conv_corners"])


CASES = ["cfg0", "cfg0_thr525_flip", "ties_int", "small", "label_oob"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("literal", [True, False])
def test_encode_matches_reference_source(golden, std_anchors, case, literal):
    g = golden("encode_decode")
    labels, boxes = g[case + "_gt_labels"], g[case + "_gt_boxes"]
    thr, flip = float(g[case + "_thr"][0]), bool(g[case + "_flip"][0])
    fn = orc.encode_image_literal if literal else orc.encode_image
    for i in range(labels.shape[0]):
        b = orc.flip_boxes_horizontal(boxes[i], 640) if flip else boxes[i]
        L, O, m = fn(std_anchors, labels[i], b, 4, thr, STD)
        assert np.array_equal(L, g[case + "_labels"][i])
        assert np.array_equal(O, g[case + "_offsets"][i])  # same NumPy log => bit-equal here
        assert (m >= 0).sum() == (np.abs(O).sum(1) > 0).sum() or case == "label_oob"


@pytest.mark.parametrize("case", CASES)
def test_decode_matches_reference_source(golden, std_anchors, case):
    g = golden("encode_decode")
    anc = golden("anchors")["std_centroids"]
    for i in range(g[case + "_offsets"].shape[0]):
        o = g[case + "_offsets"][i]
        assert np.array_equal(orc.decode_to_centroids(o, anc, STD), g[case + "_dec_centroids"][i])
        assert np.array_equal(orc.decode_to_corners(o, anc, STD), g[case + "_dec_corners"][i])


def test_encode_decode_round_trip_recovers_gt(golden, std_anchors):
    g = golden("encode_decode")
    anc = golden("anchors")["std_centroids"]
    L, O, m = orc.encode_image(std_anchors, g["cfg0_gt_labels"][0], g["cfg0_gt_boxes"][0], 4, 0.5, STD)
    dec = orc.decode_to_corners(O, anc, STD)
    pos = np.nonzero(m >= 0)[0]
    assert pos.size > 16
    np.testing.assert_allclose(dec[pos], g["cfg0_gt_boxes"][0][m[pos]], atol=2e-3)
    assert np.all(dec[m < 0] == 0)


def test_centroid_ctor_variant(golden):
    """datacoder.py:73-89: anchors given as centroids; corners are derived and
    the encode re-derives centroids from those corners."""
    g = golden("encode_decode")
    a = golden("anchors")
    c = a["std_centroids"]
    corners = np.stack(orc.centroids_to_corners(c[:, 0], c[:, 1], c[:, 2], c[:, 3]), 1)
    assert np.array_equal(corners, g["centroid_ctor_corners"])
    L, O, _ = orc.encode_image(corners, g["cfg0_gt_labels"][0], g["cfg0_gt_boxes"][0], 4, 0.5, STD)
    assert np.array_equal(L, g["centroid_ctor_labels"]) and np.array_equal(O, g["centroid_ctor_offsets"])


def test_layers_decode_and_nms_wrapper(golden):
    g = golden("layers")
    anc = golden("anchors")["std_centroids"]
    yxyx = orc.decode_batched_yxyx(g["offsets"], anc, STD)
    assert np.array_equal(yxyx, g["decoded_yxyx"])
    out, valid, aidx = orc.combined_nms(yxyx, g["scores"], 200, 200, 0.45, 0.01)
    assert np.array_equal(out, g["det_200"])
    out2, valid2, _ = orc.combined_nms(yxyx, g["scores"], 4, 10, 0.025, 0.2)
    flat = out2.reshape(-1, 6)
    assert np.array_equal(flat[flat[:, 0] > 0], g["det_bgdrop"])  # layers.py:165-166


def test_nms_against_torchvision(golden):
    tv = pytest.importorskip("torchvision")
    import torch

    g = golden("layers")
    yxyx = g["decoded_yxyx"][0]
    s = g["scores"][0, :, 1]
    keep = orc.nms_per_class(yxyx, s, 10 ** 6, 0.45, 0.01)
    cand = np.nonzero(s > np.float32(0.01))[0]
    y0 = np.minimum(yxyx[:, 0], yxyx[:, 2]); y1 = np.maximum(yxyx[:, 0], yxyx[:, 2])
    x0 = np.minimum(yxyx[:, 1], yxyx[:, 3]); x1 = np.maximum(yxyx[:, 1], yxyx[:, 3])
    xyxy = torch.from_numpy(np.stack([x0, y0, x1, y1], 1)[cand])
    tkeep = tv.ops.nms(xyxy, torch.from_numpy(s[cand]), 0.45).numpy()
    assert cand[tkeep].tolist() == keep


def test_eval_matches_reference(golden):
    g = golden("eval")
    N = g["pred_labels"].shape[0]
    no_gt = int(g["no_gt_image"][0])
    gl = [g["gt_labels"][i] if i != no_gt else np.zeros(0, np.int32) for i in range(N)]
    gb = [g["gt_boxes"][i] if i != no_gt else np.zeros((0, 4), np.float32) for i in range(N)]
    for i in range(N):
        iou = orc.iou_pred_vs_true(g["pred_labels"][i], g["pred_boxes"][i], g["gt_labels"][i], g["gt_boxes"][i])
        assert np.array_equal(iou, g["iou"][i])
    assert np.array_equal(orc.iou_pred_vs_true(g["pred_labels"][0], g["pred_boxes"][0], [], np.zeros((0, 4))),
                          g["iou_nogt"])
    for thr, key in ((0.5, "ap50"), (0.6, "ap60"), (0.7, "ap70")):
        ap = orc.average_precision(g["pred_labels"], g["pred_conf"], g["pred_boxes"], thr, gl, gb, [0, 1, 2, 3], 0)
        np.testing.assert_allclose([ap[1], ap[2], ap[3]], g[key], rtol=0, atol=1e-12)
        assert g[key].max() > 0
    np.testing.assert_allclose(g["kat_iou"], [[0.935032, 0], [0, 0.74036145], [0, 0], [0, 0]], rtol=1e-6)


def test_closed_form_equals_literal_on_tie_heavy_inputs(std_anchors):
    """P3: closed form == concat -> unique -> sequential scatter, including
    duplicated GT boxes and integer coordinates (many exact fp32 ties)."""
    rng = np.random.default_rng(5)
    n_dup = 0
    for trial in range(12):
        lab, box, _ = orc.synth_gt(rng, 1, 10, integer=True)
        box[0, 4] = box[0, 2]
        box[0, 7] = box[0, 2]
        iou = orc.iou_matrix(std_anchors, box[0])
        lit = orc.match_literal(iou, 0.5)
        cf = orc.match_closed_form(iou, 0.5)
        ref = np.full(std_anchors.shape[0], -1, np.int32)
        for a, gidx in lit:
            ref[a] = gidx
        n_dup += len(lit) - len(set(lit[:, 0].tolist()))
        assert np.array_equal(ref, cf)
    assert n_dup > 0


def test_csv_front_end():
    l, b = orc.parse_gt_csv_text("1,10.5,20,30,40\r\n3,1,2,3,4\r\n")
    assert l.tolist() == [1, 3] and b.dtype == np.float32 and b[0].tolist() == [10.5, 20, 30, 40]
    l, b = orc.parse_gt_csv_text("", sep=None)
    assert l.shape == (0,) and b.shape == (0, 4)
    f = orc.flip_boxes_horizontal(np.array([[10, 5, 50, 60]], np.float32), 640)
    assert f.tolist() == [[590, 5, 630, 60]]


def test_segmentation_suppression_matches_reference_source(golden):
    """the oracle against layers.SegmentationSuppression executed from the reference's file."""
    g = golden("seg")
    for case, expect in (("a", [1, 1, 1, 1]), ("b", [1, 1, 0, 1]), ("c", [1, 0, 0, 0]), ("d", [1, 0, 0, 1])):
        pres = orc.seg_class_presence(g[case + "_mask"])
        assert pres.tolist() == expect
        assert np.array_equal(orc.segmentation_suppression(g[case + "_mask"], g[case + "_probs"]),
                              g[case + "_suppressed"])


def test_box_iou_metric_matches_reference_source(golden, std_anchors):
    """the oracle against metrics.jaccard_iou_bounding_boxes executed from the reference's file."""
    g = golden("metric")
    cent = np.stack(orc.corners_to_centroids(*[std_anchors[:, i] for i in range(4)]), 1).astype(np.float32)
    got = orc.metric_iou_boxes(g["y_true"], g["y_pred"], cent, STD)
    np.testing.assert_allclose(got, g["value"], rtol=1e-6)
    assert abs(float(got[2]) - 1.0) < 1e-6      # perfect predictions


def test_losses_match_reference_source(golden):
    """the oracle against losses.localization_loss / confidence_loss executed from the reference's
    file (fixture: tests/golden/make_golden_losses.py), incl. an all-background image, saturated probabilities (both
    clip bounds), a batch without positives (k = 0) and one with more positives than a third of the negatives."""
    g = golden("losses")
    np.testing.assert_allclose(orc.localization_loss(g["loc_y_true"], g["loc_y_pred"]), g["loc_value"], rtol=1e-6)
    assert g["loc_value"][2] == 0 and g["loc_value"][4] == 0      # all background / perfect prediction
    np.testing.assert_allclose(orc.confidence_loss(g["conf_y_true"], g["conf_y_pred"]), g["conf_value"], rtol=1e-6)
    np.testing.assert_allclose(orc.confidence_loss(g["conf_y_true"][2:3], g["conf_y_pred"][2:3]),
                               g["conf_value_no_positive"], rtol=0, atol=0)
    np.testing.assert_allclose(orc.confidence_loss(g["conf_few_y_true"], g["conf_few_y_pred"]), g["conf_few_value"],
                               rtol=1e-6)
    # hard-negative mining: exactly min(3 * positives, negatives) background boxes are kept, the largest losses first
    _, keep, v = orc.confidence_loss(g["conf_y_true"], g["conf_y_pred"], return_keep=True)
    n_pos = int((g["conf_y_true"][..., 0] == 0).sum())
    assert int(keep.sum()) == 3 * n_pos and v[keep == 1].min() >= v[(keep == 0) & (g["conf_y_true"][..., 0] == 1)].max()


def test_loss_gradients_match_finite_differences():
    """The analytic gradients of the oracle (what the CUDA kernels are checked against) against central differences of
    a float64 restatement of the same statements on a small case."""
    rng = np.random.default_rng(4)
    B, A, C = 2, 40, 4
    y_true = np.zeros((B, A, 4), np.float32)
    y_true[:, ::5] = rng.normal(0, 1, (B, 8, 4))
    y_pred = (y_true + rng.normal(0, 0.8, y_true.shape)).astype(np.float32)
    _, g = orc.localization_loss(y_true, y_pred, with_gradient=True)

    def loc64(p):
        nb = (np.abs(y_true).sum(-1) > 0).astype(np.float64)
        d = y_true.astype(np.float64) - p
        l = np.where(np.abs(d) < 1, 0.5 * d * d, np.abs(d) - 0.5).sum(-1) * nb
        return l.sum(-1) / np.maximum(nb.sum(-1), 1)

    lab = np.zeros((B, A, C), np.float32)
    cls = rng.integers(0, C, (B, A)) * (rng.uniform(size=(B, A)) < 0.15)
    np.put_along_axis(lab, cls[..., None], 1.0, axis=-1)
    prob = orc.softmax_f32(rng.normal(0, 1, (B, A, C)).astype(np.float32))
    _, gc, keep, _ = orc.confidence_loss(lab, prob, with_gradient=True, return_keep=True)

    def conf64(p):
        la = -(lab * np.log(np.clip(p, 1e-7, 1 - 1e-7))).sum(-1)
        w = (1 - lab[..., 0]) + lab[..., 0] * keep
        return (la * w).sum(-1) / np.maximum((1 - lab[..., 0]).sum(-1), 1)

    for f, x0, grad in ((loc64, y_pred.astype(np.float64), g), (conf64, prob.astype(np.float64), gc)):
        for _ in range(25):
            b, a, c = rng.integers(0, B), rng.integers(0, A), rng.integers(0, x0.shape[-1])
            xp, xm = x0.copy(), x0.copy()
            xp[b, a, c] += 1e-6
            xm[b, a, c] -= 1e-6
            fd = (f(xp)[b] - f(xm)[b]) / 2e-6
            assert abs(fd - grad[b, a, c]) <= 1e-4 * max(1.0, abs(fd)), (b, a, c, fd, grad[b, a, c])


def test_closed_form_equals_literal_property():
    """Property test (hypothesis): on small integer grids, where exact IoU ties and duplicated winners are the rule,
    the closed form of SURVEY P3 equals the literal concat -> unique -> sequential scatter for every threshold."""
    from hypothesis import given, settings, strategies as st

    box = st.tuples(st.integers(0, 12), st.integers(0, 12), st.integers(1, 8), st.integers(1, 8)).map(
        lambda t: (t[0], t[1], t[0] + t[2] - 1, t[1] + t[3] - 1))

    @settings(max_examples=300, deadline=None)
    @given(st.lists(box, min_size=1, max_size=24), st.lists(box, min_size=1, max_size=6),
           st.sampled_from([0.0, 0.25, 0.5, 0.525, 0.9]))
    def check(anchors, gts, thr):
        a = np.asarray(anchors, np.float32)
        g = np.asarray(gts, np.float32)
        iou = orc.iou_matrix(a, g)
        ref = np.full(len(a), -1, np.int32)
        for ai, gi in orc.match_literal(iou, thr):
            ref[ai] = gi
        assert np.array_equal(ref, orc.match_closed_form(iou, thr))

    check()
