"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on (-lineinfo build).
    python tools/ncu_lines.py <rep> <kernel regex> [top]   -> executed warp instructions and stall samples per line"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
fname, lines, tot_i, tot_s = "", [], 0, 0
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit():
        try:
            inst, samp = int(r[7]), int(r[6]) if r[6].isdigit() else 0
        except ValueError:
            continue
        lines.append((inst, samp, fname, int(r[0]), r[1].strip()[:110]))
        tot_i += inst
        tot_s += samp
print(f"total warp instructions {tot_i}, samples {tot_s}")
for inst, samp, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100.0 * inst / max(tot_i, 1):5.1f}% inst {100.0 * samp / max(tot_s, 1):5.1f}% smp  {f}:{ln}  {src}")
