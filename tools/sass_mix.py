"""Aggregate an `ncu --page source --csv` dump: per-opcode executed instructions per warp."""
import csv, collections, sys
path, nwarps = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
seen = set(); data = []
for r in rows[2:]:
    if len(r) < 10 or not r[iE].isdigit(): continue
    if r[0] in seen: break
    seen.add(r[0]); data.append(r)
tot = sum(int(r[iE]) for r in data)
print(len(data), 'sass instrs; executed', tot, 'per warp', tot / nwarps)
ops = collections.Counter(); samp = collections.Counter()
for r in data:
    t = r[iS].strip().split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    ops[op] += int(r[iE]); samp[op] += int(r[iSm])
for op, c in ops.most_common(30):
    print(f"{op:10s} {c / nwarps:8.1f} per warp   stall samples {samp[op]}")
