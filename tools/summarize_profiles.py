"""Turns the scratch ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep>"""
import collections, csv, json, os, subprocess, sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
NAMES = {"k_seed_init": "seed_init", "k_assign<1>": "slic_assign_first", "k_assign<0>": "slic_assign", "k_relax": "stable_relax",
         "k_gather_depths": "slic_gather_depths", "k_newton": "slic_newton", "k_gather_points": "plane_gather_points",
         "k_fuse": "surfel_fuse", "k_init_surfels": "surfel_init", "k_repack": "repack", "k_pixel_normals": "pixel_normals",
         "k_gauss_newton": "plane_gauss_newton"}


def short(kn):
    kn = kn.replace("void ", "").split("(")[0]
    return kn


# ---- launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = short(r[4])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{tag}_launches_summary.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "bench_name", "launches", "total_us", "avg_us", "share_pct"])
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, NAMES.get(k, k), v[0], f"{v[1] / 1e3:.1f}", f"{v[1] / v[0] / 1e3:.2f}", f"{v[1] / tot * 100:.1f}"])
subprocess.run(["cp", launches, os.path.join(out, f"{tag}_launches.csv")])

# ---- full set
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(w) for w in want if w in hdr]
traffic = {}
with open(os.path.join(out, f"{tag}_ncu_full_summary.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rr[2:]:
        w.writerow([r[i] for i in idx])
        k = short(r[hdr.index("Kernel Name")])

        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic.setdefault(NAMES.get(k, k), []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out, f"{tag}_launches_summary.csv")).read())
print(json.load(open(os.path.join(out, "traffic.json"))))
