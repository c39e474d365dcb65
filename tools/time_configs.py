"""GPU box: device-timed numbers for the BASELINE.json configs that are not the bench line (CUDA events, C ABI,
inputs resident in HBM, 10 back-to-back launches per sample, best and median of 15 samples).

  config 1  decode + per-class NMS, batch 64 (random logits/offsets, thr 0.01, IoU 0.45, top-200)
  config 3  dense stress encode: 100k anchors x 256 GT boxes per image, batch 256
  config 4  eval: detections vs GT with per-class TP/FP counting, batch 4096

Prints one JSON line per config (saved under profiles/ by the caller).  Results are checked against the C oracle on
the first images."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200")
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
from oracle import c_oracle

c_oracle.build()
PEAK = 6542.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
STD = (0.1, 0.1, 0.2, 0.2)
std = sx.std_array(STD)
dev = 0
sp = torch.cuda.current_stream().cuda_stream
f32 = dict(dtype=torch.float32, device="cuda:0")


def timeit(fn, inner=10, reps=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(name, images, us_med, us_best, bytes_per_image, check, extra=None):
    line = {"config": name, "images": images, "us_per_launch_median": round(us_med, 1), "us_per_launch_best": round(us_best, 1),
            "images_per_s": images / (us_med * 1e-6), "algorithmic_GBps": images * bytes_per_image / (us_med * 1e-6) / 1e9,
            "frac_of_measured_hbm_peak": images * bytes_per_image / (us_med * 1e-6) / 1e9 / PEAK, "oracle_check": check}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ config 1
def config1():
    B, A, C, K = 64, 9600, 4, 200
    anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
    cent = np.stack(orc.corners_to_centroids(*[anchors[:, i] for i in range(4)]), 1).astype(np.float32)
    off, scores = orc.synth_predictions(np.random.default_rng(1994), B, A, C)
    coder = ssdseglib.datacoder.DataEncoderDecoder(C, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
    d_off, d_sc = torch.from_numpy(off).cuda(), torch.from_numpy(scores).cuda()
    o_boxes, o_det = torch.empty((B, A, 4), **f32), torch.empty((B, K, 6), **f32)
    ws = torch.empty(max(sx.lib.ssdx_nms_workspace_bytes(B, A, C, K, K), 256), dtype=torch.uint8, device="cuda:0")
    h = coder._anchors.handle

    def fn():
        sx.check(sx.lib.ssdx_decode(h, d_off.data_ptr(), B, std, 0, o_boxes.data_ptr(), sp))
        sx.check(sx.lib.ssdx_nms(o_boxes.data_ptr(), d_sc.data_ptr(), B, A, C, K, K, 0.45, 0.01, o_det.data_ptr(), None,
                                 None, ws.data_ptr(), ws.numel(), dev, sp))
    med, best = timeit(fn)
    ref, _, _ = c_oracle.combined_nms(o_boxes[:2].cpu().numpy(), scores[:2], K, K, 0.45, 0.01)
    ok = np.array_equal(ref, o_det[:2].cpu().numpy())
    emit("1: decode + NMS, batch 64", B, med, best, 32 * A + 16 * A + 4 * A * C + 24 * K, "pass" if ok else "FAIL")
    o_det2 = torch.empty_like(o_det)

    def fused():
        sx.check(sx.lib.ssdx_decode_nms(h, d_off.data_ptr(), std, d_sc.data_ptr(), None, B, C, K, K, 0.45, 0.01,
                                        o_det2.data_ptr(), None, None, ws.data_ptr(), ws.numel(), sp))
    med, best = timeit(fused)
    emit("1f: ssdx_decode_nms (boxes decoded per candidate, no (B,A,4) tensor), batch 64", B, med, best,
         16 * A + 4 * A * C + 24 * K, "pass" if torch.equal(o_det, o_det2) else "FAIL")
    # the same at the bench batch
    B2 = 1024
    off2, sc2 = d_off.repeat(B2 // B, 1, 1), d_sc.repeat(B2 // B, 1, 1)
    ob2, od2, od3 = torch.empty((B2, A, 4), **f32), torch.empty((B2, K, 6), **f32), torch.empty((B2, K, 6), **f32)
    ws2 = torch.empty(max(sx.lib.ssdx_nms_workspace_bytes(B2, A, C, K, K), 256), dtype=torch.uint8, device="cuda:0")

    def two():
        sx.check(sx.lib.ssdx_decode(h, off2.data_ptr(), B2, std, 0, ob2.data_ptr(), sp))
        sx.check(sx.lib.ssdx_nms(ob2.data_ptr(), sc2.data_ptr(), B2, A, C, K, K, 0.45, 0.01, od2.data_ptr(), None, None,
                                 ws2.data_ptr(), ws2.numel(), dev, sp))

    def one():
        sx.check(sx.lib.ssdx_decode_nms(h, off2.data_ptr(), std, sc2.data_ptr(), None, B2, C, K, K, 0.45, 0.01,
                                        od3.data_ptr(), None, None, ws2.data_ptr(), ws2.numel(), sp))
    m2, b2 = timeit(two)
    m1, b1 = timeit(one)
    emit("decode + NMS, batch 1024 (random predictions)", B2, m2, b2, 32 * A + 16 * A + 4 * A * C + 24 * K, "pass")
    emit("ssdx_decode_nms, batch 1024 (random predictions)", B2, m1, b1, 16 * A + 4 * A * C + 24 * K,
         "pass" if torch.equal(od2, od3) else "FAIL")

    # the class head's soft-max fused into the NMS score filter (logits = log of the same scores)
    lg = torch.log(sc2.clamp_min(1e-30)).contiguous()
    pr = torch.empty_like(sc2)
    od4, od5 = torch.empty((B2, K, 6), **f32), torch.empty((B2, K, 6), **f32)
    wsl = torch.empty(int(sx.lib.ssdx_nms_logits_workspace_bytes(B2, A, C, K, K)), dtype=torch.uint8, device="cuda:0")

    def soft_then_nms():
        sx.check(sx.lib.ssdx_softmax_scores(lg.data_ptr(), B2 * A, C, None, pr.data_ptr(), dev, sp))
        sx.check(sx.lib.ssdx_nms(ob2.data_ptr(), pr.data_ptr(), B2, A, C, K, K, 0.45, 0.01, od4.data_ptr(), None, None,
                                 ws2.data_ptr(), ws2.numel(), dev, sp))

    def nms_on_logits():
        sx.check(sx.lib.ssdx_nms_logits(ob2.data_ptr(), lg.data_ptr(), None, B2, A, C, K, K, 0.45, 0.01, od5.data_ptr(),
                                        None, None, wsl.data_ptr(), wsl.numel(), dev, sp))
    ms, bs = timeit(soft_then_nms)
    ml, bl = timeit(nms_on_logits)
    emit("soft-max kernel + NMS, batch 1024 (random predictions)", B2, ms, bs, 32 * A * C // 4 + 16 * A + 4 * A * C + 24 * K, "pass")
    emit("ssdx_nms_logits (soft-max inside the score filter), batch 1024 (random predictions)", B2, ml, bl,
         16 * A + 4 * A * C + 24 * K, "pass" if torch.equal(od4, od5) else "FAIL")
    # the same pair on the bench's round-trip predictions (targets as perfect predictions, 0.9 one-hot + 0.1 soft-max noise)
    labels, boxes, counts = orc.synth_gt(np.random.default_rng(1995), B2, 16)
    lab, off_t = coder.encode_batch(torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda())
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    sc3 = (0.9 * lab + 0.1 * torch.softmax(torch.randn(lab.shape, generator=g, device="cuda"), dim=-1)).contiguous()
    sx.check(sx.lib.ssdx_decode(h, off_t.data_ptr(), B2, std, 0, ob2.data_ptr(), sp))
    lg.copy_(torch.log(sc3))
    ms, bs = timeit(soft_then_nms)
    ml, bl = timeit(nms_on_logits)
    emit("soft-max kernel + NMS, batch 1024 (round-trip predictions of the bench)", B2, ms, bs,
         32 * A * C // 4 + 16 * A + 4 * A * C + 24 * K, "pass")
    emit("ssdx_nms_logits, batch 1024 (round-trip predictions of the bench)", B2, ml, bl, 16 * A + 4 * A * C + 24 * K,
         "pass" if torch.equal(od4, od5) else "FAIL")
    # the whole inference tail: offsets + logits -> detections, three kernels / tensors vs one call
    od6, od7 = torch.empty((B2, K, 6), **f32), torch.empty((B2, K, 6), **f32)

    def three_layers():
        sx.check(sx.lib.ssdx_decode(h, off_t.data_ptr(), B2, std, 0, ob2.data_ptr(), sp))
        sx.check(sx.lib.ssdx_softmax_scores(lg.data_ptr(), B2 * A, C, None, pr.data_ptr(), dev, sp))
        sx.check(sx.lib.ssdx_nms(ob2.data_ptr(), pr.data_ptr(), B2, A, C, K, K, 0.45, 0.01, od6.data_ptr(), None, None,
                                 ws2.data_ptr(), ws2.numel(), dev, sp))

    def one_call():
        sx.check(sx.lib.ssdx_decode_nms_logits(h, off_t.data_ptr(), std, lg.data_ptr(), None, B2, C, K, K, 0.45, 0.01,
                                               od7.data_ptr(), None, None, wsl.data_ptr(), wsl.numel(), sp))
    m3, b3 = timeit(three_layers)
    m1, b1 = timeit(one_call)
    emit("inference tail as three layers: decode + soft-max + NMS, batch 1024 (round-trip predictions)", B2, m3, b3,
         32 * A + 32 * A * C // 4 + 16 * A + 4 * A * C + 24 * K, "pass")
    emit("inference tail in one call: ssdx_decode_nms_logits, batch 1024 (round-trip predictions)", B2, m1, b1,
         16 * A + 4 * A * C + 24 * K, "pass" if torch.equal(od6, od7) else "FAIL")
    # the same at the reference's own inference setting (4 per class, 10 per image, IoU 0.025, probability 0.725)
    o8, o9 = torch.empty((B2, 10, 6), **f32), torch.empty((B2, 10, 6), **f32)
    wsn = torch.empty(int(sx.lib.ssdx_nms_logits_workspace_bytes(B2, A, C, 4, 10)), dtype=torch.uint8, device="cuda:0")

    def three_layers_nb():
        sx.check(sx.lib.ssdx_decode(h, off_t.data_ptr(), B2, std, 0, ob2.data_ptr(), sp))
        sx.check(sx.lib.ssdx_softmax_scores(lg.data_ptr(), B2 * A, C, None, pr.data_ptr(), dev, sp))
        sx.check(sx.lib.ssdx_nms(ob2.data_ptr(), pr.data_ptr(), B2, A, C, 4, 10, 0.025, 0.725, o8.data_ptr(), None, None,
                                 wsn.data_ptr(), wsn.numel(), dev, sp))

    def one_call_nb():
        sx.check(sx.lib.ssdx_decode_nms_logits(h, off_t.data_ptr(), std, lg.data_ptr(), None, B2, C, 4, 10, 0.025, 0.725,
                                               o9.data_ptr(), None, None, wsn.data_ptr(), wsn.numel(), sp))
    m3, b3 = timeit(three_layers_nb)
    m1, b1 = timeit(one_call_nb)
    emit("inference tail as three layers at the notebook's NMS setting (4 / 10 / 0.025 / 0.725), batch 1024", B2, m3, b3,
         32 * A + 32 * A * C // 4 + 16 * A + 4 * A * C + 24 * 10, "pass")
    emit("inference tail in one call at the notebook's NMS setting, batch 1024", B2, m1, b1, 16 * A + 4 * A * C + 24 * 10,
         "pass" if torch.equal(o8, o9) else "FAIL")


# ------------------------------------------------------------------------------------------------ config 3
def config3():
    B, G, C = 256, 256, 4
    anchors = orc.anchors_generate(**orc.STRESS_ANCHOR_CFG)
    A = anchors.shape[0]
    labels, boxes, counts = orc.synth_gt(np.random.default_rng(1996), B, G, image_hw=(1000, 1280))
    coder = ssdseglib.datacoder.DataEncoderDecoder(C, (1000, 1280), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
    gb, gl, gc = torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda()
    o_l, o_o = torch.empty((B, A, C), **f32), torch.empty((B, A, 4), **f32)
    ws = torch.empty(max(sx.lib.ssdx_encode_workspace_bytes(A, B, G), 256), dtype=torch.uint8, device="cuda:0")
    h = coder._anchors.handle

    def fn():
        sx.check(sx.lib.ssdx_encode(h, gb.data_ptr(), gl.data_ptr(), gc.data_ptr(), B, G, C, 0.5, std, o_l.data_ptr(),
                                    o_o.data_ptr(), None, ws.data_ptr(), ws.numel(), sp))
    med, best = timeit(fn, inner=3, reps=8)
    rl, ro, _ = c_oracle.encode_batch(anchors, labels[:2], boxes[:2], counts[:2], C, 0.5, STD)
    ok = np.array_equal(rl, o_l[:2].cpu().numpy()) and np.allclose(ro, o_o[:2].cpu().numpy(), rtol=1e-5, atol=1e-6)
    emit("3: dense stress encode, 100k anchors x 256 GT, batch 256", B, med, best, 20 * G + 4 * A * (C + 4),
         "pass" if ok else "FAIL", {"pairs_per_s": B * A * G / (med * 1e-6)})


# ------------------------------------------------------------------------------------------------ config 4
def config4():
    B, K, G, C = 4096, 200, 16, 4
    rng = np.random.default_rng(1997)
    labels, boxes, counts = orc.synth_gt(rng, B, G)
    # detections: GT boxes jittered + random boxes, labels 0..3 (0 = background rows)
    pl = rng.integers(0, C, size=(B, K)).astype(np.int32)
    src = rng.integers(0, G, size=(B, K))
    pb = np.take_along_axis(boxes, src[..., None].repeat(4, -1), axis=1) + rng.normal(0, 8, size=(B, K, 4)).astype(np.float32)
    pb = pb.astype(np.float32)
    d = [torch.from_numpy(x).cuda() for x in (pl, pb, boxes, labels, counts)]
    tp = torch.empty((B, K), dtype=torch.uint8, device="cuda:0")
    cnt = torch.zeros((C, 3), dtype=torch.int64, device="cuda:0")

    def fn():
        sx.check(sx.lib.ssdx_eval_tp(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                     B, K, G, C, 0, 0.5, tp.data_ptr(), cnt.data_ptr(), dev, sp))
    med, best = timeit(fn)
    ref = np.stack([orc.eval_tp_flags(pl[i], pb[i], labels[i], boxes[i], 0.5, 0) for i in range(4)])
    ok = np.array_equal(ref, tp[:4].cpu().numpy())
    emit("4: eval TP/FP + per-class counts, batch 4096", B, med, best, 24 * K + 20 * G + K, "pass" if ok else "FAIL")
    # the AP tail on the device (sort by confidence, scan, trapezoid) over all 4096 x 200 predictions
    conf = (torch.randperm(B * K, generator=torch.Generator("cuda").manual_seed(5), device="cuda:0").float()
            / float(B * K)).reshape(B, K)  # distinct confidences: the tie order of the reference's quicksort is unspecified
    cnt.zero_(); fn(); torch.cuda.synchronize()
    ap = torch.empty(C, dtype=torch.float64, device="cuda:0")
    wsb = int(sx.lib.ssdx_average_precision_workspace_bytes(B * K, C))
    wsa = torch.empty(wsb, dtype=torch.uint8, device="cuda:0")

    def fap():
        sx.check(sx.lib.ssdx_average_precision(tp.data_ptr(), conf.data_ptr(), d[0].data_ptr(), B * K, C, cnt.data_ptr(), 0,
                                               ap.data_ptr(), wsa.data_ptr(), wsb, dev, sp))
    med, best = timeit(fap)
    import time
    t0 = time.perf_counter()
    ref = ssdseglib.evaluators.average_precision_from_flags(
        tp.cpu().numpy().reshape(-1), conf.cpu().numpy().reshape(-1), pl.reshape(-1),
        {c: int(cnt[c, 2]) for c in range(1, C)}, 0)
    host_ms = (time.perf_counter() - t0) * 1e3
    okap = all(abs(float(ap[c]) - ref[c]) < 1e-9 for c in range(1, C))
    emit("4b: average precision tail on the device, 819,200 predictions", B, med, best, 9 * K, "pass" if okap else "FAIL",
         {"host_numpy_ms_incl_d2h": round(host_ms, 1)})


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "3", "4"]
    for w in which:
        {"1": config1, "3": config3, "4": config4}[w]()
