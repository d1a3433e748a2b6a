"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; data = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; ie = r.index("Instructions Executed"); ns = r.index("# Samples"); continue
    if hdr is None or len(r) <= ie: continue
    if r[2] != "-": continue            # SASS rows carry an address; source rows have '-'
    try: line = int(r[0]); inst = float(r[ie]); samp = float(r[ns])
    except ValueError: continue
    a = data.setdefault((line, r[1]), [0.0, 0.0, r[1]]); a[0] += inst; a[1] += samp      # (line, text): the kernels span several files
tot = sum(v[0] for v in data.values()); stot = sum(v[1] for v in data.values())
print("total warp instructions %.3e, samples %d" % (tot, stot))
for line, v in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  L%-5d %s" % (100 * v[0] / tot, 100 * v[1] / max(stot, 1), line[0], v[2][:120]))
