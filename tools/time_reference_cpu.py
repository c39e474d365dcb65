"""BUILD CONTAINER ONLY (needs /root/reference): time the parts of the reference's own CPU path that run without
TensorFlow, unchanged - ssdseglib/boxes.py (pure NumPy) and the detection side of ssdseglib/evaluators.py (an empty
`tensorflow` stub module satisfies its import; the two functions never touch it) - next to the C port the bench times.
BASELINE.md section 4 / BASELINE.json config 0.  Output: one JSON object (commit it under profiles/).
usage: python tools/time_reference_cpu.py > profiles/r2_reference_cpu_build_container.json"""
import importlib.util, json, os, sys, tempfile, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/ssdseglib"
from oracle import c_oracle, ssd_oracle as orc  # noqa: E402


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))


import warnings
warnings.simplefilter("ignore")
out = {"host": {"cpus": os.cpu_count(), "where": "build container (no GPU); /root/reference executed unchanged"}}
boxes = load("boxes")


def make_anchors():
    d = boxes.DefaultBoundingBoxes(feature_maps_shapes=((30, 40), (15, 20), (8, 10), (4, 5)),
                                   centers_padding_from_borders_percentage=(0.025, 0.05, 0.075, 0.1),
                                   boxes_scales=(0.15, 0.95), additional_square_box=True)
    d.rescale_boxes_coordinates((480, 640))
    return d.get_boxes_coordinates_corners("ssd")


b, m = best_of(make_anchors, 20)
anchors = make_anchors()
assert np.array_equal(anchors, orc.anchors_generate(**orc.STD_ANCHOR_CFG))
out["boxes.DefaultBoundingBoxes + rescale + corners('ssd'), 9600 anchors"] = {"best_ms": 1e3 * b, "median_ms": 1e3 * m}

sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
ev = load("evaluators")
rng = np.random.default_rng(1997)
N, K, G, C = 256, 200, 16, 4
pl, conf, pb, gl, gb, gc = orc.synth_eval(rng, N, K, G, C)
b, m = best_of(lambda: ev._iou_boxes_pred_vs_true(pl[0], pb[0], gl[0], gb[0]), 200)
out["evaluators._iou_boxes_pred_vs_true, 200 predictions x 16 GT"] = {"best_us": 1e6 * b, "median_us": 1e6 * m}
with tempfile.TemporaryDirectory() as tmp:
    paths = []
    for i in range(N):
        p = os.path.join(tmp, f"{i}.csv")
        with open(p, "w") as f:
            for l, bx in zip(gl[i], gb[i]):
                f.write(f"{int(l)},{float(bx[0])!r},{float(bx[1])!r},{float(bx[2])!r},{float(bx[3])!r}\n")
        paths.append(p)
    t0 = time.perf_counter()
    ap_ref = ev.average_precision_object_detection(pl, conf, pb, 0.5, paths, [0, 1, 2, 3], 0)
    t_ref = time.perf_counter() - t0
tp, cnt = c_oracle.eval_tp(pl, pb, gb, gl, gc, C, 0, 0.5, nthreads=1)
ap_port = orc.average_precision_from_tp(tp, conf, pl, {c: int(cnt[c, 2]) for c in range(1, C)})
assert all(abs(ap_ref[c] - ap_port[c]) < 1e-9 for c in (1, 2, 3)), (ap_ref, ap_port)
out[f"evaluators.average_precision_object_detection, {N} images x 200 detections x 16 GT (CSV files)"] = {
    "seconds": t_ref, "images_per_s": N / t_ref, "equals_port_to": 1e-9}
b1, _ = best_of(lambda: c_oracle.eval_tp(pl, pb, gb, gl, gc, C, 0, 0.5, nthreads=1), 10)
out["C port of the TP/FP pass on the same inputs, 1 thread"] = {"images_per_s": N / b1}

labels, gtb, counts = orc.synth_gt(np.random.default_rng(1993), 8, 16)
for nt in (1, os.cpu_count() or 1):
    b, m = best_of(lambda: c_oracle.encode_batch(anchors, labels, gtb, counts, 4, 0.5, (0.1, 0.1, 0.2, 0.2), nthreads=nt), 50)
    out[f"config 0: C port of encode (datacoder.py:205-300), 8 images x 16 GT, {nt} thread(s)"] = {
        "images_per_s": 8 / b, "flags": " ".join(c_oracle.CFLAGS)}
b, m = best_of(lambda: [orc.encode_image(anchors, labels[i], gtb[i], 4) for i in range(8)], 5)
out["config 0: NumPy fp32 restatement of encode (oracle/ssd_oracle.py), 8 images x 16 GT, 1 thread"] = {"images_per_s": 8 / b}
out["note"] = ("datacoder.py / layers.py need TensorFlow 2.13 (205 + 72 tf.* call sites) and cannot run here; the NumPy "
               "restatement above executes the same statements op by op and is the closest stand-in for the TF-eager cost")
print(json.dumps(out, indent=1))
