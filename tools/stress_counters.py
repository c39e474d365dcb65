import ctypes, os, sys
ROOT='/root/repo'
PKG=os.path.join(ROOT,"multi-task-learning-object-detection-semantic-segmentation_b200")
os.environ["SSDX_LIB"]=os.path.join(PKG,"lib","libssdx_dbg.so")
sys.path.insert(0,ROOT); sys.path.insert(0,PKG)
import numpy as np, torch
import ssdseglib
from ssdseglib import _ssdx as sx
from oracle import ssd_oracle as orc
B,G,C=64,256,4
anchors=orc.anchors_generate(**orc.STRESS_ANCHOR_CFG)
labels,boxes,counts=orc.synth_gt(np.random.default_rng(1996),B,G,image_hw=(1000,1280))
coder=ssdseglib.datacoder.DataEncoderDecoder(C,(1000,1280),anchors[:,0],anchors[:,1],anchors[:,2],anchors[:,3])
out=(ctypes.c_ulonglong*32)()
sx.lib.ssdx_debug_phase_cycles_encode.argtypes=[ctypes.c_void_p]
gb,gl,gc=torch.from_numpy(boxes).cuda(),torch.from_numpy(labels).cuda(),torch.from_numpy(counts).cuda()
for it in range(2):
    sx.lib.ssdx_debug_phase_cycles_encode(out)
    coder.encode_batch(gb,gl,gc); torch.cuda.synchronize()
sx.lib.ssdx_debug_phase_cycles_encode(out)
v=list(out)
print(f"stress per image: surviving (tile,GT) {v[25]/B:.0f} of {1563*256}  slow-path entries {v[26]/B:.0f}  divisions {v[27]/B:.0f}  GT-side updates {v[28]/B:.0f}")
