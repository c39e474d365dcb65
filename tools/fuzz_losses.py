"""GPU box: randomized differential test of the SSD losses against the NumPy oracle: random shapes and class counts,
positive fractions from 0 to more than a third of the boxes, soft-max / saturated / constant predictions (massive loss
ties at the hard-negative threshold), gradients and kept-background masks.
usage: python tools/fuzz_losses.py [cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np
import ssdseglib
from oracle import ssd_oracle as orc


def run(cases: int, seed: int, verbose: bool = True) -> int:
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(cases):
        B = int(rng.integers(1, 7)); A = int(rng.choice([1, 5, 63, 512, 513, 2000, 9600])); C = int(rng.choice([2, 3, 4, 4, 4, 8]))
        frac = float(rng.choice([0.0, 0.01, 0.05, 0.3, 0.6, 1.0]))
        cls = rng.integers(1, C, (B, A)) * (rng.uniform(size=(B, A)) < frac)
        y_true = np.zeros((B, A, C), np.float32)
        np.put_along_axis(y_true, cls[..., None], 1.0, axis=-1)
        kind = rng.choice(["softmax", "saturated", "constant", "quantised"])
        p = orc.softmax_f32((rng.normal(0, 2, (B, A, C)) + 1.5 * y_true).astype(np.float32))
        if kind == "saturated":
            sat = rng.uniform(size=(B, A)) < 0.5
            p[sat] = 0.0
            p[sat, 0] = 1.0
        elif kind == "constant":
            p[:] = np.float32(1.0 / C)
        elif kind == "quantised":
            p = (np.round(p * 16) / 16).astype(np.float32)
        p = p.astype(np.float32)
        loss, grad, keep = ssdseglib.losses.confidence_loss(y_true, p, with_gradient=True, return_kept_background=True)
        ref, gref, kref, v = orc.confidence_loss(y_true, p, with_gradient=True, return_keep=True)
        ok = np.allclose(loss, ref, rtol=1e-5, atol=1e-30)
        diff = keep != kref
        if diff.any():  # only losses within a few ulp of the threshold may be classified differently (logf vs np.log)
            t = v[kref == 1].min() if kref.any() else np.float32(0)
            ok = ok and bool(np.all(np.abs(v[diff] - t) <= 8 * np.spacing(np.float32(t)))) and int(keep.sum()) == int(kref.sum())
        same = ~diff
        ok = ok and np.allclose(grad[same], gref[same], rtol=1e-5, atol=1e-12)
        # localization on the matching offsets tensor
        t_off = np.zeros((B, A, 4), np.float32)
        t_off[cls > 0] = rng.normal(0, 1, (int((cls > 0).sum()), 4))
        p_off = (t_off + rng.normal(0, rng.choice([0.1, 1.0, 3.0]), t_off.shape)).astype(np.float32)
        l2, g2 = ssdseglib.losses.localization_loss(t_off, p_off, with_gradient=True)
        r2, gr2 = orc.localization_loss(t_off, p_off, with_gradient=True)
        ok = ok and np.allclose(l2, r2, rtol=1e-5, atol=1e-30) and np.allclose(g2, gr2, rtol=1e-5, atol=1e-12)
        if not ok:
            bad += 1
            print(f"MISMATCH case {it}: B={B} A={A} C={C} frac={frac} pred={kind} keep diffs={int(diff.sum())} "
                  f"loss {loss.tolist()} vs {ref.tolist()}", flush=True)
    if verbose:
        print(f"fuzz_losses: {cases} cases, {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
