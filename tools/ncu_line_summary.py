"""Per-source-line summary of an ncu report: warp-instructions executed and stall samples per CUDA line.

usage: python tools/ncu_line_summary.py REPORT.ncu-rep KERNEL_REGEX [LIB.so] [min_pct]

The SASS page of the report (ncu --page source --csv) has no line numbers, so the kernel's SASS is also
disassembled with `nvdisasm -g` from the library's cubin and the two listings are aligned by instruction order.
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre = sys.argv[1], sys.argv[2]
lib = os.path.abspath(sys.argv[3]) if len(sys.argv) > 3 else os.path.join(
    ROOT, "multi-task-learning-object-detection-semantic-segmentation_b200", "lib", "libssdx.so")
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel instance only
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
kname = rows[start - 1][1] if start > 0 else "?"
isrc, iinst, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(r[isrc].strip(), int(r[iinst] or 0), int(r[isamp] or 0)) for r in rows[start + 1:end] if len(r) > isamp]

func = re.search(r"(ssdx_\w+)", kname).group(1)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur, infunc, seq, cands = None, False, [], []
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            if infunc and seq:
                cands.append(seq)
            infunc = func in ln
            seq = []
        elif infunc:
            m = re.search(r'//## File ".*?", line (\d+)', ln)
            if m:
                cur = int(m.group(1))
            elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                seq.append((cur, ln.split("*/", 1)[1].strip().rstrip(";").strip()))
    if infunc and seq:
        cands.append(seq)
    for seq in cands:  # template instantiations share the name: take the one with the same instruction count
        if abs(len(seq) - len(sass)) <= 2:
            lines = seq
            break
    if lines is not None:
        break
if lines is None:
    raise SystemExit(f"could not align SASS of {func} (ncu has {len(sass)} instructions)")

src_path = None
agg = collections.defaultdict(lambda: [0, 0])
for (ln, _), (_, ni, ns) in zip(lines, sass):
    agg[ln][0] += ni
    agg[ln][1] += ns
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
srcs = {}
for f in glob.glob(os.path.join(os.path.dirname(os.path.dirname(lib)), "csrc", "*")):
    srcs[os.path.basename(f)] = open(f).read().splitlines()
# guess file: the .cu whose name appears in the kernel's mangled cubin name is unknown here; print text from the
# file that has that many lines and contains the function name
cand = [k for k, v in srcs.items() if any(func in l for l in v)]
text = srcs[cand[0]] if cand else []
print(f"{func}: {ti:,} warp-instructions, {ts:,} samples ({cand[0] if cand else '?'}; lines of inlined headers map by number only)")
for ln in sorted(agg, key=lambda k: (k is None, k)):
    ni, ns = agg[ln]
    if 100 * ni / max(ti, 1) >= min_pct or 100 * ns / max(ts, 1) >= min_pct:
        t = text[ln - 1].strip()[:100] if ln and ln <= len(text) else ""
        print(f"{str(ln):>5}  inst {100 * ni / max(ti, 1):5.1f}%  samples {100 * ns / max(ts, 1):5.1f}%  {t}")
