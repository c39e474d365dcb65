#!/usr/bin/env python
"""bench.py - images/s of the SSD box path (encode -> decode -> NMS round trip) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic input:
    encode(GT boxes) -> labels, offsets ; decode(offsets) -> boxes ; NMS(boxes, scores) -> detections
with B = 1024 images per GPU, A = 9600 anchors, G = 16 GT boxes, C = 4 classes, top-200.  Images are
independent: ranks shard the batch with no data-path collective ("scaling": "weak", 1024 images per GPU).

`value`  : device-timed (CUDA events, max over ranks), inputs resident in HBM, C ABI called directly.
`e2e`    : same step through the public ssdseglib API; every step copies the step's inputs (GT + class
           scores) from pinned host memory and reads the detections back to the host.
`roofline`: dominant stage, algorithmic bytes / measured stage time vs the measured HBM copy peak.
`cpu_baseline` / `--impl reference`: the C/OpenMP port of the reference path (oracle/ssd_oracle.c) on the
           host cores of this box (TensorFlow 2.13, the reference's backend, cannot be installed here).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

STD = (0.1, 0.1, 0.2, 0.2)
A_STD, G_STD, C_STD, K_STD = 9600, 16, 4, 200
BYTES_PER_IMAGE = {  # BASELINE.md section 3 (algorithmic / compulsory HBM traffic)
    "encode": 20 * G_STD + 4 * A_STD * (C_STD + 4),
    "decode": 32 * A_STD,
    "nms": 16 * A_STD + 4 * A_STD * C_STD + 24 * K_STD,
}
WORKLOADS = {
    "roundtrip_b1024": dict(batch=1024, stages=("encode", "decode", "nms")),
    "decode_nms_b64": dict(batch=64, stages=("decode", "nms")),  # BASELINE config 1
    # BASELINE config 5: detections vs GT, TP/FP flags + per-class counts, 4096 images in total sharded over the ranks,
    # NCCL all-gather of the (C,3) count tables, AP tail on the global arrays (evaluators.py:65-187)
    "eval_b4096": dict(batch=4096, stages=("eval",)),
}
EVAL_K, EVAL_THR = 200, 0.5
BYTES_PER_IMAGE["eval"] = 24 * EVAL_K + 20 * G_STD + EVAL_K  # predictions + GT in, TP flags out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------
# synthetic inputs: GT generator of config 1; scores = 0.9*onehot + 0.1*softmax(noise)
# ----------------------------------------------------------------------------------------------------------
def make_inputs(batch: int, seed: int):
    from oracle import ssd_oracle as orc  # generators only (input synthesis, not the measured path)

    rng = np.random.default_rng(seed)
    anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
    labels, boxes, counts = orc.synth_gt(rng, batch, G_STD)
    return anchors, labels, boxes, counts


def make_scores_torch(onehot, seed: int):
    """scores = 0.9 * one-hot targets + 0.1 * softmax(N(0,1) logits), generated on the device once (setup)."""
    import torch

    g = torch.Generator(device=onehot.device)
    g.manual_seed(seed)
    noise = torch.randn(onehot.shape, generator=g, device=onehot.device, dtype=torch.float32)
    return (0.9 * onehot + 0.1 * torch.softmax(noise, dim=-1)).contiguous()


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) running during the timed regions
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: C/OpenMP port of the reference path
# ----------------------------------------------------------------------------------------------------------
def cpu_roundtrip(c_oracle, anchors, cent, labels, boxes, counts, scores, stages, threads):
    """One pass of the same stages over `len(boxes)` images on the host."""
    off = None
    if "encode" in stages:
        _, off, _ = c_oracle.encode_batch(anchors, labels, boxes, counts, C_STD, 0.5, STD, nthreads=threads,
                                          want_match=False)
    else:
        off = scores[1]
    yxyx = c_oracle.decode_batch(cent, off, STD, mode=0, nthreads=threads)
    return c_oracle.combined_nms(yxyx, scores[0], K_STD, K_STD, 0.45, 0.01, nthreads=threads)


def cpu_inputs(batch, seed, stages):
    from oracle import ssd_oracle as orc

    anchors, labels, boxes, counts = make_inputs(batch, seed)
    cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1).astype(np.float32)
    rng = np.random.default_rng(seed + 7)
    if "encode" in stages:
        from oracle import c_oracle

        onehot, _, _ = c_oracle.encode_batch(anchors, labels, boxes, counts, C_STD, 0.5, STD,
                                             nthreads=os.cpu_count() or 1, want_match=False)
        scores = (np.float32(0.9) * onehot + np.float32(0.1) * orc.softmax_f32(
            rng.normal(0, 1, onehot.shape).astype(np.float32)))
        pred_off = None
    else:
        pred_off, scores = orc.synth_predictions(rng, batch, A_STD, C_STD)
    return anchors, cent, labels, boxes, counts, (scores.astype(np.float32), pred_off)


def run_cpu_sample(stages, sample_images, seed, threads, target_seconds=10.0, max_passes=64):
    """Times the C/OpenMP port on `sample_images` images, repeating the pass until about `target_seconds` of host
    work have been done.  Returns (images/s over all passes, seconds, passes)."""
    from oracle import c_oracle

    c_oracle.build()
    anchors, cent, labels, boxes, counts, scores = cpu_inputs(sample_images, seed, stages)
    cpu_roundtrip(c_oracle, anchors, cent, labels[:threads], boxes[:threads], counts[:threads],
                  (scores[0][:threads], None if scores[1] is None else scores[1][:threads]), stages, threads)  # warm
    total, passes = 0.0, 0
    while passes < max_passes and (passes == 0 or total < target_seconds):
        t0 = time.perf_counter()
        cpu_roundtrip(c_oracle, anchors, cent, labels, boxes, counts, scores, stages, threads)
        total += time.perf_counter() - t0
        passes += 1
    return sample_images * passes / total, total, passes


def cpu_config0(c_oracle, threads):
    """BASELINE config 0: the reference CPU path's own runnable case - encode of 8 synthetic images x 16 GT boxes
    (3 foreground classes) against the SSDLite-MobileNetV2 anchor set - on one core and on all cores of this host."""
    from oracle import ssd_oracle as orc

    anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
    labels, boxes, counts = orc.synth_gt(np.random.default_rng(1993), 8, G_STD)
    out = {"workload": "encode, 8 images x 16 GT boxes, 9600 anchors (C port of datacoder.py:205-300)"}
    for name, nt in (("one_core", 1), ("all_cores", threads)):
        c_oracle.encode_batch(anchors, labels, boxes, counts, C_STD, 0.5, STD, nthreads=nt)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 0.5:
            c_oracle.encode_batch(anchors, labels, boxes, counts, C_STD, 0.5, STD, nthreads=nt)
            reps += 1
        out[name] = {"images_per_s": 8 * reps / (time.perf_counter() - t0), "threads": nt}
    return out


def reference_arm(args, rank):
    """`--impl reference`: the host-CPU port of the reference path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    from oracle import c_oracle

    c_oracle.build()
    if args.workload == "eval_b4096":
        return reference_arm_eval(args, c_oracle, threads)
    sample = wl["batch"]  # the GPU arm's batch: same config keys, same images per step
    anchors, cent, labels, boxes, counts, scores = cpu_inputs(sample, 1995, wl["stages"])
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cpu_roundtrip(c_oracle, anchors, cent, labels, boxes, counts, scores, wl["stages"], threads)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "images/sec encode+decode+NMS", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "gpus_used": 0,  # n_gpus = the N of the launch this line pairs with; the arm is CPU-only
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": roundtrip_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} images per step of the {args.workload} workload, C/OpenMP port "
                                   f"(oracle/ssd_oracle.c, gcc {' '.join(c_oracle.CFLAGS)} -fopenmp), {threads} threads"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def roundtrip_config(args, world):
    """`config` of the JSON line - the same dict from both arms (the driver compares them): the workload, and how the
    GPU arm launches it.  Arm-specific facts (oracle check, host threads, sample) live outside `config`."""
    wl = WORKLOADS[args.workload]
    B, stages = wl["batch"], wl["stages"]
    n_pipe = max(1, min(args.pipeline, B))
    if args.pipeline_split:
        n_pipe = len(args.pipeline_split.split(","))
    n_streams = max(1, min(args.pipeline_streams, n_pipe))
    nms_split = max(1, min(args.nms_split, B))
    launch = ("eager" if args.no_graph else "one CUDA graph per step") + (
        "" if args.no_overlap else (
            (f"; NMS stage in {nms_split} sub-batches: ssdx_nms_select of one on the step stream while ssdx_nms_finish "
             "of the previous one runs on a second stream" if nms_split > 1 else
             "; NMS candidate selection (ssdx_nms_select) on a second stream, concurrent with encode/decode, joined "
             "before ssdx_nms_finish") if n_pipe == 1 else
            f"; {n_pipe} sub-batches of {B // n_pipe} images pipelined on {n_streams} streams"))
    return {"workload": args.workload, "stages": list(stages), "images_per_gpu": B, "global_batch": B * world,
            "anchors": A_STD, "gt_per_image": G_STD, "classes": C_STD, "max_total": K_STD, "score_thr": 0.01,
            "iou_thr": 0.45, "parallelism": f"batch-sharded x{world}, no collective",
            "l2": "no flush: each step streams %.0f MB (> 126 MB L2)" % (
                sum(BYTES_PER_IMAGE[s] for s in stages) * B / 1e6),
            "gpu_launch": launch}


def eval_config(world):
    B = WORKLOADS["eval_b4096"]["batch"]
    return {"workload": "eval_b4096", "stages": ["eval"], "global_batch": B, "images_per_gpu": B // world,
            "predictions_per_image": EVAL_K, "gt_per_image": G_STD, "classes": C_STD, "iou_thr": EVAL_THR,
            "parallelism": f"images sharded x{world}" + (
                "; ssdx_allgather_eval: the (C,3) int64 count tables and the TP / confidence / label arrays for the AP "
                "tail in one NCCL launch (ncclGroupStart/End)" if world > 1 else ", no collective"),
            "l2": "flushed with a 256 MB fill before every timed step (the 17 MB working set would stay in L2)",
            "gpu_launch": "eager, per-step events"}


def reference_arm_eval(args, c_oracle, threads):
    """BASELINE config 5 on the host: TP flags + counts (C/OpenMP port of evaluators.py:106-157) and the AP tail
    (NumPy, evaluators.py:159-187) over all 4096 images per step."""
    from oracle import ssd_oracle as orc

    B = WORKLOADS["eval_b4096"]["batch"]
    pl, conf, pb, gl, gb, gc = orc.synth_eval(np.random.default_rng(1997), B, EVAL_K, G_STD, C_STD)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tp, cnt = c_oracle.eval_tp(pl, pb, gb, gl, gc, C_STD, 0, EVAL_THR, nthreads=threads)
        orc.average_precision_from_tp(tp, conf, pl, {c: int(cnt[c, 2]) for c in range(1, C_STD)})
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = B / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "images/sec eval (TP/FP + counts + AP)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": eval_config(max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"all {B} images per step: C/OpenMP port of the TP/FP pass ({threads} threads) + "
                                   "NumPy AP tail (1 thread)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_arm_eval(args, rank, world, local_rank):
    """BASELINE config 5: 4096 images of detections vs GT sharded over the ranks (strong scaling), per-class TP/FP/GT
    counts all-gathered over NCCL (ssdx_allgather_counts), AP on the device from the global flags."""
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    import __graft_entry__ as entry

    if rank == 0:
        entry.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = local_rank
    import ssdseglib
    from ssdseglib import _dist, _ssdx as sx
    from oracle import ssd_oracle as orc  # generators (input synthesis) and, after the timed regions, the checker

    B, K, G, C = WORKLOADS["eval_b4096"]["batch"], EVAL_K, G_STD, C_STD
    pl, conf, pb, gl, gb, gc = orc.synth_eval(np.random.default_rng(1997), B, K, G, C)  # same global batch on every rank
    lo, hi = _dist.shard_range(B, rank, world)
    n = hi - lo
    cuda = f"cuda:{dev}"
    host = [torch.from_numpy(np.ascontiguousarray(x[lo:hi])).pin_memory() for x in (pl, conf, pb, gl, gb, gc)]
    d_pl, d_conf, d_pb, d_gl, d_gb, d_gc = [t.to(cuda) for t in host]
    tp = torch.empty((n, K), dtype=torch.uint8, device=cuda)
    counts = torch.zeros((C, 3), dtype=torch.int64, device=cuda)
    gathered = torch.zeros((world, C, 3), dtype=torch.int64, device=cuda)
    total_counts = torch.zeros((C, 3), dtype=torch.int64, device=cuda)
    # global flag arrays for the AP tail (rank-major = image-major: the shards are contiguous)
    sizes = [_dist.shard_range(B, r, world) for r in range(world)]
    even = all(h - l == n for l, h in sizes)
    g_tp = torch.empty((B, K), dtype=torch.uint8, device=cuda) if world > 1 else tp
    g_conf = torch.empty((B, K), dtype=torch.float32, device=cuda) if world > 1 else d_conf
    g_pl = torch.empty((B, K), dtype=torch.int32, device=cuda) if world > 1 else d_pl
    ap = torch.empty((C,), dtype=torch.float64, device=cuda)
    ws_ap = torch.empty(int(sx.lib.ssdx_average_precision_workspace_bytes(B * K, C)), dtype=torch.uint8, device=cuda)
    comm = _dist.CountsCommunicator(None, dev) if world > 1 else None
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    assert even or world == 1, "eval_b4096 shards evenly over 1, 2, 4 and 8 ranks"

    def step():
        counts.zero_()
        sx.check(sx.lib.ssdx_eval_tp(d_pl.data_ptr(), d_pb.data_ptr(), d_gb.data_ptr(), d_gl.data_ptr(), d_gc.data_ptr(),
                                     n, K, G, C, 0, EVAL_THR, tp.data_ptr(), counts.data_ptr(), dev, sp))
        if world > 1:
            # the one collective of the path: (C,3) int64 per rank, NCCL all-gather on this stream
            # and the flags the AP tail integrates (7.4 MB in all), in the same NCCL launch
            sx.check(sx.lib.ssdx_allgather_eval(comm.handle, counts.data_ptr(), gathered.data_ptr(), C, tp.data_ptr(),
                                                d_conf.data_ptr(), d_pl.data_ptr(), n * K, g_tp.data_ptr(),
                                                g_conf.data_ptr(), g_pl.data_ptr(), sp))
            sx.check(sx.lib.ssdx_sum_counts(gathered.data_ptr(), world, C, total_counts.data_ptr(), dev, sp))
            cnt = total_counts
        else:
            cnt = counts
        sx.check(sx.lib.ssdx_average_precision(g_tp.data_ptr(), g_conf.data_ptr(), g_pl.data_ptr(), B * K, C,
                                               cnt.data_ptr(), 0, ap.data_ptr(), ws_ap.data_ptr(), ws_ap.numel(), dev, sp))
        return cnt

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step()
    # the working set (17 MB per rank at N = 1) fits the 126 MB L2: flush it between timed steps with a 256 MB fill
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=cuda)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_tp = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        evs[k][0].record(stream)
        counts.zero_()
        sx.check(sx.lib.ssdx_eval_tp(d_pl.data_ptr(), d_pb.data_ptr(), d_gb.data_ptr(), d_gl.data_ptr(), d_gc.data_ptr(),
                                     n, K, G, C, 0, EVAL_THR, tp.data_ptr(), counts.data_ptr(), dev, sp))
        ev_tp[k].record(stream)
        if world > 1:
            sx.check(sx.lib.ssdx_allgather_eval(comm.handle, counts.data_ptr(), gathered.data_ptr(), C, tp.data_ptr(),
                                                d_conf.data_ptr(), d_pl.data_ptr(), n * K, g_tp.data_ptr(),
                                                g_conf.data_ptr(), g_pl.data_ptr(), sp))
            sx.check(sx.lib.ssdx_sum_counts(gathered.data_ptr(), world, C, total_counts.data_ptr(), dev, sp))
        sx.check(sx.lib.ssdx_average_precision(g_tp.data_ptr(), g_conf.data_ptr(), g_pl.data_ptr(), B * K, C,
                                               (total_counts if world > 1 else counts).data_ptr(), 0, ap.data_ptr(),
                                               ws_ap.data_ptr(), ws_ap.numel(), dev, sp))
        evs[k][1].record(stream)
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    tp_ms = float(np.median([evs[k][0].elapsed_time(ev_tp[k]) for k in range(args.steps)]))
    total_ms = float(np.sum(step_ms))

    # ---------------- end to end through the public API with host arrays ----------------
    pg = dist.group.WORLD if dist is not None else None
    h_np = [t.numpy() for t in host]

    def step_e2e():
        tp_h, cnt_h = ssdseglib.evaluators.true_positives_and_counts(h_np[0], h_np[2], h_np[3], h_np[4], h_np[5],
                                                                      EVAL_THR, C, 0, process_group=pg)
        return tp_h, cnt_h

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tp_h, cnt_h = step_e2e()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    sampler.stop()
    h2d = sum(h_np[i].nbytes for i in (0, 2, 3, 4, 5))
    d2h = tp_h.nbytes + cnt_h.nbytes

    # ---------------- check: ALL images against the oracle (outside the timed regions) ----------------
    final_counts = step().cpu().numpy()
    torch.cuda.synchronize(dev)
    check = "skipped"
    if rank == 0:
        from oracle import c_oracle

        ref_tp, ref_cnt = c_oracle.eval_tp(pl, pb, gb, gl, gc, C, 0, EVAL_THR)
        ref_ap = orc.average_precision_from_tp(ref_tp, conf, pl, {c: int(ref_cnt[c, 2]) for c in range(1, C)})
        ok = np.array_equal(g_tp.cpu().numpy(), ref_tp) and np.array_equal(final_counts, ref_cnt)
        ok = ok and np.array_equal(np.asarray(cnt_h), ref_cnt) and np.array_equal(np.asarray(tp_h), ref_tp[lo:hi])
        ap_h = ap.cpu().numpy()
        ok = ok and all(abs(ap_h[c] - ref_ap[c]) < 1e-9 for c in range(1, C))
        check = "pass" if ok else "FAIL"
    if dist is not None:
        t = torch.tensor([total_ms, e2e_ms, tp_ms], dtype=torch.float64, device=cuda)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, tp_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        peak, peak_src = measured_peak()
        ms_per_step = total_ms / args.steps
        achieved = BYTES_PER_IMAGE["eval"] * n / (tp_ms / 1e3) / 1e9
        line = {
            "metric": "images/sec eval (TP/FP + counts + AP)", "value": B * args.steps / (total_ms / 1e3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": eval_config(world), "oracle_check": check + " (all 4096 images: flags, counts, AP)",
            "roofline": {"bound": "hbm", "kernel": "ssdx_eval_tp_kernel", "stage": "eval", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_IMAGE["eval"] * n,
                         "note": "instruction-bound: 13 M label-masked IoU pairs over 5,320 B per image",
                         "stages": {"eval_tp": {"ms": tp_ms}, "collective+ap": {"ms": ms_per_step - tp_ms}}},
            "e2e": {"value": B * e2e_steps / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "ssdseglib.evaluators.true_positives_and_counts(NumPy in, NumPy out"
                           + (", process_group=WORLD)" if world > 1 else ")")},
            # eval_tp + the seven kernels of ssdx_average_precision (sample, bucket, columns, offsets, scatter, terms, sum)
            # (+ sum_counts after the all-gather)
            "gpu_launches": (8 + (1 if world > 1 else 0)) * args.steps,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import c_oracle

            threads = os.cpu_count() or 1
            t0 = time.perf_counter()
            passes = 0
            while time.perf_counter() - t0 < 10.0:
                tpc, cc = c_oracle.eval_tp(pl, pb, gb, gl, gc, C, 0, EVAL_THR, nthreads=threads)
                orc.average_precision_from_tp(tpc, conf, pl, {c: int(cc[c, 2]) for c in range(1, C)})
                passes += 1
            secs = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": passes * B / secs, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"{passes} passes over all {B} images ({secs:.1f} s of host work): C/OpenMP "
                                              f"TP/FP pass on {threads} threads + NumPy AP tail"}
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    _dist.close_communicators()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def bind_to_gpu_cores(torch, dev, allowed):
    """sched_setaffinity to the CPU cores NVML reports as local to the GPU (PCIe root / NUMA node).  Returns a note
    for the JSON line; any failure leaves the affinity untouched."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(dev)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & set(allowed)
        if cpus and cpus != set(allowed):
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} of {len(allowed)} cores (GPU-local)"
        return f"all {len(allowed)} cores (GPU reports no narrower affinity)"
    except Exception as exc:  # no NVML, container cpuset, ...
        return f"unchanged ({type(exc).__name__})"


def gpu_arm(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    import __graft_entry__ as entry

    if rank == 0:
        entry.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = local_rank
    # host side of the end-to-end leg: run this rank (and first-touch its pinned buffers) on the CPU cores next to its
    # GPU, so that N ranks copying at PCIe rate do not all pull through one socket's memory controller / the
    # inter-socket link.  The original affinity is restored for the CPU baseline, which uses every core.
    cpu_all = os.sched_getaffinity(0)
    numa_note = bind_to_gpu_cores(torch, dev, cpu_all) if not args.no_numa else "off"
    import ssdseglib
    from ssdseglib import _ssdx as sx

    wl = WORKLOADS[args.workload]
    B, stages = wl["batch"], wl["stages"]
    anchors, labels, boxes, counts = make_inputs(B, 1995 + rank)
    coder = ssdseglib.datacoder.DataEncoderDecoder(C_STD, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2],
                                                   anchors[:, 3], iou_threshold=0.5,
                                                   standard_deviations_centroids_offsets=STD, device=dev)
    dec_layer = ssdseglib.layers.DecodeBoxesCentroidsOffsets(coder.center_x_boxes_default, coder.center_y_boxes_default,
                                                             coder.width_boxes_default, coder.height_boxes_default,
                                                             *STD, device=dev)
    nms_layer = ssdseglib.layers.NonMaximumSuppression(K_STD, K_STD, 0.45, 0.01, False)
    handle = coder._anchors.handle

    # pinned host inputs (the e2e leg copies from these every step)
    h_boxes = torch.from_numpy(boxes).pin_memory()
    h_labels = torch.from_numpy(labels).pin_memory()
    h_counts = torch.from_numpy(counts).pin_memory()
    d_boxes, d_labels, d_counts = h_boxes.cuda(dev), h_labels.cuda(dev), h_counts.cuda(dev)
    # setup (untimed): class scores derived from the encoded targets, resident in HBM
    if "encode" in stages:
        onehot, _ = coder.encode_batch(d_boxes, d_labels, d_counts)
        d_scores = make_scores_torch(onehot, 4242 + rank)
        del onehot
        d_pred_off = None
    else:
        from oracle import ssd_oracle as orc

        off, sc = orc.synth_predictions(np.random.default_rng(1994 + rank), B, A_STD, C_STD)
        d_scores, d_pred_off = torch.from_numpy(sc).cuda(dev), torch.from_numpy(off).cuda(dev)
    h_scores = d_scores.cpu().pin_memory()
    h_pred_off = None if d_pred_off is None else d_pred_off.cpu().pin_memory()

    # preallocated outputs + workspaces; the C ABI is called directly (no allocation in the timed region)
    f32 = dict(dtype=torch.float32, device=f"cuda:{dev}")
    o_labels, o_offsets = torch.empty((B, A_STD, C_STD), **f32), torch.empty((B, A_STD, 4), **f32)
    o_boxes, o_det = torch.empty((B, A_STD, 4), **f32), torch.empty((B, K_STD, 6), **f32)
    o_valid = torch.empty((B,), dtype=torch.int32, device=f"cuda:{dev}")
    ws_enc = torch.empty(max(sx.lib.ssdx_encode_workspace_bytes(A_STD, B, G_STD), 256), dtype=torch.uint8, device=f"cuda:{dev}")
    ws_nms = torch.empty(max(sx.lib.ssdx_nms_workspace_bytes(B, A_STD, C_STD, K_STD, K_STD), 256), dtype=torch.uint8,
                         device=f"cuda:{dev}")
    std = sx.std_array(STD)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    side = torch.cuda.Stream(dev)
    # --pipeline P: the batch is cut into P contiguous sub-batches that run encode -> decode -> NMS on alternating
    # streams, so that the memory-bound kernels of one sub-batch (decode, NMS select) overlap the compute-bound ones
    # of the other (encode match, NMS sort + suppression).  Same calls, same work, same outputs; P = 1 is one stream.
    n_pipe = max(1, min(args.pipeline, B))
    pipe_bounds = [(B * i // n_pipe, B * (i + 1) // n_pipe) for i in range(n_pipe)]
    if args.pipeline_split:  # explicit sub-batch sizes (e.g. multiples of the SM count: one-CTA-per-image kernels then
        sizes = [int(v) for v in args.pipeline_split.split(",")]   # leave no SM with an extra image)
        assert sum(sizes) == B and all(v > 0 for v in sizes), "--pipeline-split must sum to the batch"
        n_pipe = len(sizes)
        edges = [0]
        for v in sizes:
            edges.append(edges[-1] + v)
        pipe_bounds = list(zip(edges[:-1], edges[1:]))
    ws_enc_p = [ws_enc] + [torch.empty_like(ws_enc) for _ in range(n_pipe - 1)]
    ws_nms_p = [ws_nms] + [torch.empty_like(ws_nms) for _ in range(n_pipe - 1)]
    nms_split = max(1, min(args.nms_split, B))
    nms_bounds = [(B * i // nms_split, B * (i + 1) // nms_split) for i in range(nms_split)]
    ws_nms_s = [ws_nms] + [torch.empty_like(ws_nms) for _ in range(nms_split - 1)]
    n_streams = max(1, min(args.pipeline_streams, n_pipe))
    pipe_streams = [None, side] + [torch.cuda.Stream(dev) for _ in range(max(0, n_streams - 2))]  # None = step stream
    pipe_streams = pipe_streams[:n_streams]

    def run_range(lo, hi, s_ptr, ws_e, ws_n, events=None, ev_stream=None, overlap_side=None):
        n = hi - lo
        i = 0
        if overlap_side is not None:
            overlap_side.wait_stream(ev_stream)
            sx.check(sx.lib.ssdx_nms_select(d_scores[lo:hi].data_ptr(), None, n, A_STD, C_STD, K_STD, K_STD, 0.01,
                                            ws_n.data_ptr(), ws_n.numel(), dev, overlap_side.cuda_stream))
        if events is not None:
            events[i].record(ev_stream); i += 1
        if "encode" in stages:
            sx.check(sx.lib.ssdx_encode(handle, d_boxes[lo:hi].data_ptr(), d_labels[lo:hi].data_ptr(),
                                        d_counts[lo:hi].data_ptr(), n, G_STD, C_STD, 0.5, std,
                                        o_labels[lo:hi].data_ptr(), o_offsets[lo:hi].data_ptr(), None, ws_e.data_ptr(),
                                        ws_e.numel(), s_ptr))
            if events is not None:
                events[i].record(ev_stream); i += 1
        src = o_offsets if "encode" in stages else d_pred_off
        sx.check(sx.lib.ssdx_decode(dec_layer._anchors.handle, src[lo:hi].data_ptr(), n, std, 0,
                                    o_boxes[lo:hi].data_ptr(), s_ptr))
        if events is not None:
            events[i].record(ev_stream); i += 1
        if overlap_side is not None:
            ev_stream.wait_stream(overlap_side)
            sx.check(sx.lib.ssdx_nms_finish(o_boxes[lo:hi].data_ptr(), d_scores[lo:hi].data_ptr(), None, n, A_STD, C_STD,
                                            K_STD, K_STD, 0.45, 0.01, o_det[lo:hi].data_ptr(), o_valid[lo:hi].data_ptr(),
                                            None, ws_n.data_ptr(), ws_n.numel(), dev, s_ptr))
        else:
            sx.check(sx.lib.ssdx_nms(o_boxes[lo:hi].data_ptr(), d_scores[lo:hi].data_ptr(), n, A_STD, C_STD, K_STD, K_STD,
                                     0.45, 0.01, o_det[lo:hi].data_ptr(), o_valid[lo:hi].data_ptr(), None,
                                     ws_n.data_ptr(), ws_n.numel(), dev, s_ptr))
        if events is not None:
            events[i].record(ev_stream)

    def step_device(events=None, overlap=False):
        """One step through the C ABI.  events: single stream, one event after every stage (roofline pass).
        overlap: the pipelined form described above (P sub-batches on two streams); with P = 1 the NMS candidate
        selection (reads only the scores) runs on the second stream while encode and decode run on the first one,
        as the two input branches of the reference's NMS layer allow (models.py:402-414)."""
        nonlocal sp
        cur = torch.cuda.current_stream(dev)
        if events is None and args.nms_beside_decode:
            # encode; then, side by side, the decode kernel (materialises the (B, A, 4) box tensor: memory-bound) and
            # the NMS call that reads the offsets itself (ssdx_decode_nms: the few hundred boxes an image's candidates
            # touch are decoded on the fly with the decode kernel's own arithmetic - detections bit-identical to NMS of
            # the decoded tensor; its second kernel is instruction-bound)
            if "encode" in stages:
                sx.check(sx.lib.ssdx_encode(handle, d_boxes.data_ptr(), d_labels.data_ptr(), d_counts.data_ptr(), B,
                                            G_STD, C_STD, 0.5, std, o_labels.data_ptr(), o_offsets.data_ptr(), None,
                                            ws_enc.data_ptr(), ws_enc.numel(), sp))
            src = o_offsets if "encode" in stages else d_pred_off
            side.wait_stream(cur)
            sx.check(sx.lib.ssdx_decode(dec_layer._anchors.handle, src.data_ptr(), B, std, 0, o_boxes.data_ptr(), sp))
            sx.check(sx.lib.ssdx_decode_nms(dec_layer._anchors.handle, src.data_ptr(), std, d_scores.data_ptr(), None, B,
                                            C_STD, K_STD, K_STD, 0.45, 0.01, o_det.data_ptr(), o_valid.data_ptr(), None,
                                            ws_nms.data_ptr(), ws_nms.numel(), side.cuda_stream))
            cur.wait_stream(side)
            return
        if events is not None or not overlap:
            run_range(0, B, sp, ws_enc, ws_nms, events, stream if events is not None else cur)
            return
        if n_pipe == 1 and nms_split > 1:
            # encode and decode over the whole batch, then the NMS in sub-batches: selection of sub-batch j + 1 on the
            # step stream while sort + suppression of sub-batch j run on the side stream
            if "encode" in stages:
                sx.check(sx.lib.ssdx_encode(handle, d_boxes.data_ptr(), d_labels.data_ptr(), d_counts.data_ptr(), B,
                                            G_STD, C_STD, 0.5, std, o_labels.data_ptr(), o_offsets.data_ptr(), None,
                                            ws_enc.data_ptr(), ws_enc.numel(), sp))
            src = o_offsets if "encode" in stages else d_pred_off
            sx.check(sx.lib.ssdx_decode(dec_layer._anchors.handle, src.data_ptr(), B, std, 0, o_boxes.data_ptr(), sp))
            for j, (lo, hi) in enumerate(nms_bounds):
                n = hi - lo
                sx.check(sx.lib.ssdx_nms_select(d_scores[lo:hi].data_ptr(), None, n, A_STD, C_STD, K_STD, K_STD, 0.01,
                                                ws_nms_s[j].data_ptr(), ws_nms_s[j].numel(), dev, sp))
                side.wait_stream(cur)
                sx.check(sx.lib.ssdx_nms_finish(o_boxes[lo:hi].data_ptr(), d_scores[lo:hi].data_ptr(), None, n, A_STD,
                                                C_STD, K_STD, K_STD, 0.45, 0.01, o_det[lo:hi].data_ptr(),
                                                o_valid[lo:hi].data_ptr(), None, ws_nms_s[j].data_ptr(),
                                                ws_nms_s[j].numel(), dev, side.cuda_stream))
            cur.wait_stream(side)
            return
        if n_pipe == 1:
            run_range(0, B, sp, ws_enc, ws_nms, None, cur, overlap_side=side)
            return
        for st in pipe_streams[1:]:
            st.wait_stream(cur)
        for p, (lo, hi) in enumerate(pipe_bounds):
            st = pipe_streams[p % len(pipe_streams)]
            run_range(lo, hi, sp if st is None else st.cuda_stream, ws_enc_p[p], ws_nms_p[p], None,
                      cur if st is None else st)
        for st in pipe_streams[1:]:
            cur.wait_stream(st)

    # kernels of one step: fused encode (1) + decode (1) + NMS merged pipeline (select_all, fast_all)
    kernels_per_step = (1 if "encode" in stages else 0) + 1 + 2

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize(dev)
    # One step = one CUDA graph launch (the C ABI is stream-ordered and allocation-free, so it captures as is):
    # no host-side launch gaps between the kernels of a step.  --no-graph times the eager launches instead.
    graph = None
    if not args.no_graph:
        outer_sp = sp
        try:
            cap_stream = torch.cuda.Stream(dev)
            cap_stream.wait_stream(stream)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream):
                sp = torch.cuda.current_stream(dev).cuda_stream
                step_device(overlap=not args.no_overlap)
            sp = outer_sp
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize(dev)
        except Exception as exc:  # capture is an optimisation of the launch path only: fall back to eager launches
            sys.stderr.write(f"bench.py: CUDA graph capture failed ({exc}); timing eager launches\n")
            sp = outer_sp
            graph = None
            torch.cuda.synchronize(dev)
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    t_start.record(stream)
    for k in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            step_device(overlap=not args.no_overlap)
    t_end.record(stream)
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    # per-stage durations (roofline): the same K steps again, eager, with an event after every stage
    n_ev = len(stages) + 1
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(args.steps)]
    for k in range(args.steps):
        step_device(evs[k])
    barrier()
    stage_ms = {s: float(np.median([evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)]))
                for i, s in enumerate(stages)}

    # ---------------- end-to-end through the public API with host buffers ----------------
    # The batch is fed in chunks: chunk i + 1 crosses PCIe while chunk i is computed (two streams; every call below is
    # the public API).  Per step all of the step's inputs are copied from pinned host memory and all of its
    # detections are read back.
    n_chunks = max(1, min(args.e2e_chunks, B))
    bounds = [(B * i // n_chunks, B * (i + 1) // n_chunks) for i in range(n_chunks)]
    copy_stream = torch.cuda.Stream(dev)
    h_det = torch.empty((B, K_STD, 6), dtype=torch.float32).pin_memory()

    # Double-buffered across steps: the copies of step i + 1 are enqueued (copy stream) before the compute of step i
    # (step stream), so PCIe stays busy while the GPU works; the host waits for step i's detections before it
    # reuses that step's pinned result buffer.
    h_det2 = [h_det, torch.empty_like(h_det).pin_memory()]

    def e2e_stage():
        staged = []
        with torch.cuda.stream(copy_stream):
            for lo, hi in bounds:
                gb = h_boxes[lo:hi].to(f"cuda:{dev}", non_blocking=True)
                gl = h_labels[lo:hi].to(f"cuda:{dev}", non_blocking=True)
                gc = h_counts[lo:hi].to(f"cuda:{dev}", non_blocking=True)
                sc = h_scores[lo:hi].to(f"cuda:{dev}", non_blocking=True)
                po = None if "encode" in stages else h_pred_off[lo:hi].to(f"cuda:{dev}", non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                staged.append((gb, gl, gc, sc, po, ev))
        return staged

    def e2e_compute(staged, out_host):
        for (lo, hi), (gb, gl, gc, sc, po, ev) in zip(bounds, staged):
            stream.wait_event(ev)
            for t in (gb, gl, gc, sc, po):
                if t is not None:
                    t.record_stream(stream)
            if "encode" in stages:
                _, off = coder.encode_batch(gb, gl, gc)
            else:
                off = po
            yx = dec_layer(off)
            det = nms_layer(yx, sc)
            out_host[lo:hi].copy_(det, non_blocking=True)  # device -> host read of the chunk's result
        done = torch.cuda.Event()
        done.record(stream)
        return done

    def run_e2e(n_steps):
        staged = e2e_stage()
        out = None
        for i in range(n_steps):
            nxt = e2e_stage() if i + 1 < n_steps else None
            done = e2e_compute(staged, h_det2[i & 1])
            done.synchronize()  # step i's detections are on the host
            out = h_det2[i & 1]
            staged = nxt
        return out

    e2e_steps = max(3, min(args.steps, 20))
    run_e2e(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    det_host = run_e2e(e2e_steps)
    e1.record(stream)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * e2e_wall)

    # two more end-to-end figures through the public API (NumPy in, NumPy out), as the reference uses the path:
    # (a) the training-data pipeline: GT boxes -> encoded targets back on the host (datacoder.py:302-347, cell 8);
    # (b) the inference tail: raw head outputs -> detections (models.py:402-414) in one call
    e2e_more = {}
    if "encode" in stages:
        n_more = max(2, min(e2e_steps, 5))
        np_in = (boxes, labels, counts)
        for _ in range(3):  # the pinned result blocks come from torch's caching host allocator: two generations of them
            enc_l, enc_o = coder.encode_batch(*np_in)  # are alive at a time, both must exist before the timed calls
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_more):
            enc_l, enc_o = coder.encode_batch(*np_in)
        t_enc = (time.perf_counter() - t0) / n_more
        e2e_more["encode_numpy_in_numpy_out"] = {
            "value": B / t_enc, "unit": "images/s", "ms_per_step": 1e3 * t_enc,
            "h2d_bytes_per_step": int(boxes.nbytes + labels.nbytes + counts.nbytes),
            "d2h_bytes_per_step": int(enc_l.nbytes + enc_o.nbytes),
            "api": "DataEncoderDecoder.encode_batch(NumPy) -> (labels, offsets) NumPy"}
        h_off_np = o_offsets.cpu().numpy()
        h_log_np = torch.log(d_scores.clamp_min(1e-30)).cpu().numpy()
        for _ in range(2):
            nms_layer.call_on_offsets_and_logits(dec_layer, h_off_np, h_log_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_more):
            det_np = nms_layer.call_on_offsets_and_logits(dec_layer, h_off_np, h_log_np)
        t_tail = (time.perf_counter() - t0) / n_more
        e2e_more["inference_tail_numpy_in_numpy_out"] = {
            "value": B / t_tail, "unit": "images/s", "ms_per_step": 1e3 * t_tail,
            "h2d_bytes_per_step": int(h_off_np.nbytes + h_log_np.nbytes), "d2h_bytes_per_step": int(det_np.nbytes),
            "api": "NonMaximumSuppression.call_on_offsets_and_logits(decode_layer, offsets, logits) NumPy -> NumPy"}
        del h_off_np, h_log_np
    sampler.stop()

    h2d = h_scores.numel() * 4 + (h_boxes.numel() * 4 + h_labels.numel() * 4 + h_counts.numel() * 4
                                  if "encode" in stages else h_pred_off.numel() * 4)
    d2h = det_host.numel() * 4

    # ---------------- sanity: first images against the oracle (outside the timed regions) ----------------
    check = "skipped"
    if rank == 0:
        from oracle import c_oracle
        from oracle import ssd_oracle as orc

        n_chk = 2
        cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1).astype(np.float32)
        yx = o_boxes[:n_chk].cpu().numpy()
        ref_det, _, _ = c_oracle.combined_nms(yx, d_scores[:n_chk].cpu().numpy(), K_STD, K_STD, 0.45, 0.01)
        ok = np.array_equal(ref_det, o_det[:n_chk].cpu().numpy())
        if "encode" in stages:
            rl, ro, _ = c_oracle.encode_batch(anchors, labels[:n_chk], boxes[:n_chk], counts[:n_chk], C_STD, 0.5, STD)
            ok = ok and np.array_equal(rl, o_labels[:n_chk].cpu().numpy()) and np.allclose(
                ro, o_offsets[:n_chk].cpu().numpy(), rtol=1e-5, atol=1e-6)
        ok = ok and np.array_equal(det_host[:n_chk].numpy(), o_det[:n_chk].cpu().numpy())
        check = "pass" if ok else "FAIL"

    # ---------------- reduce over ranks ----------------
    if dist is not None:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)
    e2e_value = world * B * e2e_steps / (e2e_ms / 1e3)

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant KERNEL: encode and decode are one kernel each; the NMS stage is two main kernels (select_all reads
        # the scores, fast_all works out of shared memory) plus two gated fallbacks, none of them longer than encode
        single = [s_ for s_ in stages if s_ in ("encode", "decode")]
        dom = max(single, key=stage_ms.get)
        dom_bytes = BYTES_PER_IMAGE[dom] * B
        achieved = dom_bytes / (stage_ms[dom] / 1e3) / 1e9
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        stage_roof = {s: {"ms": stage_ms[s], "gbs": BYTES_PER_IMAGE[s] * B / (stage_ms[s] / 1e3) / 1e9,
                          "frac": BYTES_PER_IMAGE[s] * B / (stage_ms[s] / 1e3) / 1e9 / peak} for s in stages}
        kernels = {"encode": "ssdx_encode_grid_fused_kernel", "decode": "ssdx_decode_kernel",
                   "nms": "ssdx_nms_select_all_kernel + ssdx_nms_fast_all_kernel"}
        line = {
            "metric": "images/sec encode+decode+NMS", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": roundtrip_config(args, world), "oracle_check": check, "graph_captured": graph is not None,
            "roofline": {"bound": "hbm", "kernel": kernels[dom], "stage": dom, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "stages": stage_roof,
                         "roundtrip_frac": sum(BYTES_PER_IMAGE[s] for s in stages) * B / (ms_per_step / 1e3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "ssdseglib.datacoder.DataEncoderDecoder.encode_batch -> layers.DecodeBoxesCentroidsOffsets"
                           " -> layers.NonMaximumSuppression (inputs from pinned host memory, detections to pinned host"
                           f" memory, {n_chunks} chunks per step on two streams, copies of step i + 1 under the compute"
                           " of step i)",
                    "host_affinity": numa_note, "more": e2e_more},
            "gpu_launches": (kernels_per_step * (1 if args.no_overlap else n_pipe)
                             + (2 * (nms_split - 1) if (not args.no_overlap and n_pipe == 1) else 0)) * args.steps,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, cpu_all)
            threads = os.cpu_count() or 1
            sample = 1024
            v, secs, passes = run_cpu_sample(stages, sample, 1995, threads)
            from oracle import c_oracle

            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"{passes} passes over {sample} images of the {args.workload} workload "
                                              f"({secs:.1f} s of host work), C/OpenMP port of the reference "
                                              f"(oracle/ssd_oracle.c, gcc {' '.join(c_oracle.CFLAGS)} -fopenmp), "
                                              f"{threads} threads",
                                    "config0": cpu_config0(c_oracle, threads)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="roundtrip_b1024", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph per step")
    ap.add_argument("--overlap", action="store_true",
                    help="multi-stream forms of the step (--pipeline / --nms-split, or the NMS candidate selection on a "
                         "second stream).  Default: one stream - since encode became bandwidth-bound every "
                         "overlapped form measured slower (211 us single stream, 217-277 us overlapped)")
    ap.add_argument("--pipeline", type=int, default=1, help="sub-batches per step pipelined on two streams (1 = off)")
    ap.add_argument("--nms-split", type=int, default=1,
                    help="sub-batches of the NMS stage alone, pipelined on two streams: the candidate selection "
                         "(memory-bound) of one overlaps the sort + suppression (instruction-bound) of the other")
    ap.add_argument("--nms-beside-decode", action="store_true",
                    help="after encode, run the decode kernel and the NMS call that reads the offsets (ssdx_decode_nms) "
                         "side by side on two streams")
    ap.add_argument("--pipeline-streams", type=int, default=2)
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the CPU cores next to its GPU")
    ap.add_argument("--pipeline-split", default="", help="comma-separated sub-batch sizes instead of an even split")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="chunks per step of the end-to-end leg (copy/compute overlap)")
    args = ap.parse_args()
    args.no_overlap = not args.overlap
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun so `python bench.py --gpus N` works on its own
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload",
               args.workload] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []) + (
                   ["--no-graph"] if args.no_graph else []) + (["--overlap"] if args.overlap else []) + [
                   "--e2e-chunks", str(args.e2e_chunks), "--pipeline", str(args.pipeline), "--nms-split",
                   str(args.nms_split),
                   "--pipeline-streams", str(args.pipeline_streams)] + (
                   ["--pipeline-split", args.pipeline_split] if args.pipeline_split else []) + (
                   ["--no-numa"] if args.no_numa else [])
        raise SystemExit(subprocess.call(cmd))
    if args.workload == "eval_b4096":
        gpu_arm_eval(args, rank, world, local_rank)
    else:
        gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
