"""Compact per-kernel digest of an `ncu --set full` report: duration, instructions, issue / occupancy, pipes, memory
throughputs, DRAM bytes and the top warp-stall reasons.   python tools/ncu_digest.py gpurun_out/<tag>/full.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return default
    u = units[i].lower()
    return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)


KEYS = [("us", "gpu__time_duration.sum", 1e-3), ("Minst", "smsp__inst_executed.sum", 1e-6),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1), ("smemKB", "launch__shared_mem_per_block_dynamic", 1e-3),
        ("occ_lim_smem", "launch__occupancy_limit_shared_mem", 1), ("occ_lim_reg", "launch__occupancy_limit_registers", 1),
        ("dramRdMB", "dram__bytes_read.sum", 1e-6), ("dramWrMB", "dram__bytes_write.sum", 1e-6),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1), ("lsu_wave%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1), ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1), ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
        ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1)]
stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = []
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
    d = {"kernel": name}
    for k, m, sc in KEYS:
        d[k] = round(val(r, m) * sc, 2)
    st = sorted(((val(r, c, 0.0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stall_cols), reverse=True)[:4]
    d["stalls"] = " ".join(f"{n}={v:.1f}" for v, n in st)
    out.append(d)
keys = ["kernel"] + [k for k, _, _ in KEYS] + ["stalls"]
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
w.writerow(keys)
for d in out:
    w.writerow([d[k] for k in keys])
