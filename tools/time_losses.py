"""GPU box: device-timed SSD training losses (losses.localization_loss / confidence_loss) at batch
1024 on the encoded targets of the bench generator.  CUDA events, C ABI, inputs resident, 10 back-to-back calls per
sample, median of 15.  Prints one JSON line per case; a plain-torch statement of the same losses is timed beside it
(the natural baseline a training loop would otherwise run on the GPU)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200"))
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
PEAK = 6542.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
sp = torch.cuda.current_stream().cuda_stream
def timeit(fn, inner=10, reps=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort(); return ts[len(ts) // 2]
B, A, C = 1024, 9600, 4
g = torch.Generator(device="cuda"); g.manual_seed(1)
anchors = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
coder = ssdseglib.datacoder.DataEncoderDecoder(C, (480, 640), anchors[:, 0], anchors[:, 1], anchors[:, 2], anchors[:, 3])
labels, boxes, counts = orc.synth_gt(np.random.default_rng(3), B, 16)
lab, off = coder.encode_batch(torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(counts).cuda())
p_off = off + 0.5 * torch.randn(off.shape, generator=g, device="cuda")
p_lab = torch.softmax(2.0 * lab + torch.randn(lab.shape, generator=g, device="cuda"), dim=-1).contiguous()
loss = torch.empty(B, device="cuda"); g_off = torch.empty_like(off); g_lab = torch.empty_like(lab)
wsb = int(sx.lib.ssdx_confidence_loss_workspace_bytes(B, A)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
lo, hi = float(np.float32(1e-7)), float(np.float32(1 - 1e-7))
def emit(case, us, nbytes, extra=None):
    d = {"case": case, "images": B, "us": round(us, 1), "algorithmic_GBps": nbytes / us / 1e3,
         "frac_of_measured_hbm_peak": nbytes / us / 1e3 / PEAK}
    d.update(extra or {}); print(json.dumps(d), flush=True)
def loc(grad): sx.check(sx.lib.ssdx_localization_loss(off.data_ptr(), p_off.data_ptr(), B, A, loss.data_ptr(), g_off.data_ptr() if grad else None, 0, sp))
def conf(grad): sx.check(sx.lib.ssdx_confidence_loss(lab.data_ptr(), p_lab.data_ptr(), B, A, C, lo, hi, loss.data_ptr(), g_lab.data_ptr() if grad else None, None, ws.data_ptr(), wsb, 0, sp))
emit("localization_loss, value only", timeit(lambda: loc(False)), 2 * off.numel() * 4)
emit("localization_loss, value + gradient", timeit(lambda: loc(True)), 3 * off.numel() * 4)
emit("confidence_loss, value only", timeit(lambda: conf(False)), 2 * lab.numel() * 4)
emit("confidence_loss, value + gradient", timeit(lambda: conf(True)), 3 * lab.numel() * 4)
# plain torch statements of the reference's lines (same math, eager ops), for scale
def torch_loc():
    nb = (off.abs().sum(-1) > 0).float(); d = (off - p_off); ea = d.abs()
    l = torch.where(ea < 1, d * d * 0.5, ea - 0.5).sum(-1) * nb
    return l.sum(-1) / nb.sum(-1).clamp(min=1)
def torch_conf():
    is_bg = lab[:, :, 0]; nb = (is_bg - 1).abs()
    la = -(lab * torch.log(p_lab.clamp(lo, hi))).sum(-1)
    k = min(3 * int(torch.count_nonzero(nb)), int(torch.count_nonzero(is_bg)))
    v = (la * is_bg).reshape(-1); idx = torch.topk(v, k, sorted=False).indices
    keep = torch.zeros_like(v); keep[idx] = 1
    return ((la * nb).sum(-1) + (v * keep).reshape(B, A).sum(-1)) / nb.sum(-1).clamp(min=1)
ours = loss.clone(); conf(False); torch.cuda.synchronize()
ref = torch_conf(); rel = float(((loss - ref).abs() / ref.abs().clamp(min=1e-12)).max())
emit("torch eager localization_loss (value only)", timeit(torch_loc, inner=3, reps=7), 2 * off.numel() * 4)
emit("torch eager confidence_loss (value only, topk over B*A)", timeit(torch_conf, inner=3, reps=7), 2 * lab.numel() * 4,
     {"max_rel_diff_vs_ssdx": rel})
