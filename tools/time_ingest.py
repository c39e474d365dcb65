"""Host-side: files/s of the batched GT CSV ingest (ssdx_read_gt_csv_batch, C++ threads) next to the reference's own
per-file Python loop (evaluators.py:109-121: open + csv.reader + int()/float()).  No GPU needed."""
import csv, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np
from ssdseglib import _ssdx as sx
N, G = 8192, 16
rng = np.random.default_rng(0)
d = tempfile.mkdtemp()
paths = []
for i in range(N):
    b = rng.uniform(0, 600, (G, 4)).astype(np.float32)
    l = rng.integers(1, 4, G)
    p = os.path.join(d, f"{i}.csv")
    with open(p, "w", newline="") as f:
        f.write("".join(f"{int(a)},{float(x[0])!r},{float(x[1])!r},{float(x[2])!r},{float(x[3])!r}\r\n" for a, x in zip(l, b)))
    paths.append(p)
t0 = time.perf_counter()
for p in paths:  # the reference's loop
    labels, boxes = [], []
    with open(p, "r") as f:
        for label, xmin, ymin, xmax, ymax in csv.reader(f):
            labels.append(int(label)); boxes.append([float(xmin), float(ymin), float(xmax), float(ymax)])
t_ref = time.perf_counter() - t0
print(f"reference-style Python loop: {N / t_ref:,.0f} files/s (1 thread)")
for th in (1, 4, 0):
    sx.read_gt_csv_batch(paths[:64], crlf_only=False, threads=th)
    t0 = time.perf_counter()
    L, B, C = sx.read_gt_csv_batch(paths, crlf_only=False, threads=th)
    dt = time.perf_counter() - t0
    print(f"ssdx_read_gt_csv_batch threads={th or os.cpu_count()}: {N / dt:,.0f} files/s  ({N * G / dt / 1e6:.2f} M boxes/s)")
