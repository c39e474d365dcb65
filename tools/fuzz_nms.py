"""GPU box: randomized differential test of the NMS dispatch (merged / warp-per-segment / CTA-per-segment / generic)
against the C oracle: random shapes, class counts, caps, thresholds, score distributions (distinct, quantised to force
ties, saturated), box layouts (spread, clustered, degenerate).  usage: python tools/fuzz_nms.py [cases] [seed] [nms_cta]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np
import ssdseglib
from oracle import c_oracle


def run(cases: int, seed: int, verbose: bool = True) -> int:
    c_oracle.build()
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(cases):
        B = int(rng.integers(1, 5)); A = int(rng.choice([1, 7, 33, 300, 1500, 5000, 12000])); C = int(rng.choice([1, 2, 4, 4, 4, 9]))
        mpc = int(rng.choice([1, 2, 4, 8, 20, 32, 33, 64, 200, 600])); mtot = int(rng.choice([1, 3, 10, 50, 200, 300, 512, 700]))
        iou = float(rng.choice([0.0, 0.025, 0.3, 0.45, 0.7, 1.0])); thr = float(rng.choice([-1.0, 0.0, 0.01, 0.3, 0.725, 0.95]))
        layout = rng.choice(["spread", "cluster", "same", "degenerate"])
        if layout == "cluster":
            ctr = rng.uniform(50, 600, size=(B, 1 + A // 40, 2))[np.arange(B)[:, None], rng.integers(0, 1 + A // 40, size=(B, A))] + rng.normal(0, 5, size=(B, A, 2))
        elif layout == "same":
            ctr = np.full((B, A, 2), 300.0)
        else:
            ctr = rng.uniform(0, 640, size=(B, A, 2))
        wh = rng.uniform(5, 150, size=(B, A, 2)) if layout != "same" else np.full((B, A, 2), 60.0)
        yxyx = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
        if layout == "degenerate":
            yxyx[:, ::3, 2] = yxyx[:, ::3, 0]          # zero height
            yxyx[:, 1::5] = yxyx[:, 1::5][..., [2, 3, 0, 1]]  # flipped corners
        kind = rng.choice(["distinct", "quantised", "saturated", "onehot"])
        s = rng.uniform(0, 1, size=(B, A, C)).astype(np.float32)
        if kind == "quantised":
            s = np.round(s * 8) / 8
        elif kind == "saturated":
            s = np.where(s > 0.3, np.float32(1.0), s)
        elif kind == "onehot":
            k = rng.integers(0, C, size=(B, A)); s = (0.1 * s); np.put_along_axis(s, k[..., None], 0.9 + 0.1 * rng.uniform(size=(B, A, 1)), axis=-1)
        s = s.astype(np.float32)
        layer = ssdseglib.layers.NonMaximumSuppression(mpc, mtot, iou, thr, False)
        out, valid, aidx = layer.call(yxyx, s, return_aux=True)
        ref, vref, aref = c_oracle.combined_nms(yxyx, s, mpc, mtot, iou, thr)
        ok = np.array_equal(valid, vref) and np.array_equal(aidx, aref) and np.array_equal(out, ref)
        if not ok:
            bad += 1
            print(f"MISMATCH case {it}: B={B} A={A} C={C} mpc={mpc} mtot={mtot} iou={iou} thr={thr} layout={layout} scores={kind} "
                f"valid {valid.tolist()} vs {vref.tolist()}", flush=True)
    if verbose:
        print(f"fuzz_nms: {cases} cases, {bad} mismatches")
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 3:  # pinned CTA width of the merged pipeline: 256 | 512 | 1024
        from ssdseglib import _ssdx
        _ssdx.set_option("nms_cta", sys.argv[3])
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
