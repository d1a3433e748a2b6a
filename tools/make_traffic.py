"""profiles/r02_traffic.json from the `--page raw --csv` exports of tools/ncu_capture.sh: DRAM and L2 traffic per launch and issue
utilisation of the raster kernel of each bench workload (read by bench.py for roofline.traffic)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for arg in sys.argv[1:]:
    wl, path = arg.split("=")
    rows = list(csv.reader(open(path)))
    h, r = rows[0], rows[2]
    d = {k: r[i] for i, k in enumerate(h)}
    f = lambda k: float(d[k].replace(",", "")) if d.get(k) not in (None, "") else None
    unit = lambda k: rows[1][h.index(k)]
    def to_bytes(k):
        v, u = f(k), unit(k)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    out[wl] = {"kernel": d["Kernel Name"].split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
               "lts_t_bytes": f("lts__t_sectors.sum") * 32, "kernel_us_under_ncu": f("gpu__time_duration.sum") * (1000 if unit("gpu__time_duration.sum") == "ms" else 1),
               "sm_issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "sm_cycles_active_over_elapsed": f("sm__cycles_active.avg") / f("sm__cycles_elapsed.avg"),
               "pipe_pct": {p: f(f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active") for p in ("alu", "fma", "xu", "lsu", "adu")},
               "source": "ncu --set full --clock-control none (tools/ncu_capture.sh), " + os.path.basename(path)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps({k: (round(v["traffic_bytes_per_launch"] / 1e6, 1), round(v["lts_t_bytes"] / 1e6, 1), v["sm_issue_active_pct"]) for k, v in out.items()}))
