"""One line per kernel launch of an ncu report (ncu --set full): time, DRAM bytes, warp-instructions, registers,
occupancy, issue utilisation, pipe utilisation and the four largest stall reasons.
usage: python tools/ncu_kernel_summary.py REPORT.ncu-rep [tag]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; tag = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}
def g(r, k):
    i = h.index(k)
    try: return float(r[i].replace(",", "")) * scale.get(units[i], 1)
    except Exception: return float("nan")
stall = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k]
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    st = sorted(((g(r, k), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall), reverse=True)[:4]
    print(f"{tag} {name}: {g(r, 'gpu__time_duration.sum'):.1f} us, dram read {g(r, 'dram__bytes_read.sum') / 1e6:.1f} MB write "
          f"{g(r, 'dram__bytes_write.sum') / 1e6:.1f} MB, warp-inst {g(r, 'smsp__inst_executed.sum') / 1e6:.1f} M, regs "
          f"{int(g(r, 'launch__registers_per_thread'))}, warps_active {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}%, "
          f"issue_active {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}%, alu "
          f"{g(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.0f}%, fma "
          f"{g(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.0f}%, xu "
          f"{g(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.0f}%; stalls: " + ", ".join(f"{n} {v:.2f}" for v, n in st))
