"""Fold the key raw metrics of one kernel launch of an .ncu-rep into profiles/<name>.json under a label.

    python tools/ncu_summary.py gpurun_out/c2_full.ncu-rep profiles/r01_k_raster_summary.json c2_textured_1080p/final
"""
import csv, io, json, os, subprocess, sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_active.avg",
        "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, out, label = sys.argv[1:4]
    # an .ncu-rep, or the `--page raw --csv` export of one (tools/ncu_capture.sh)
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units, r = rows[0], rows[1], rows[2]
    d = {}
    for k in KEEP:
        if k in h:
            i = h.index(k)
            d[k] = (r[i] + " " + units[i]).strip()
    if "lts__t_sectors.sum" in d:       # L2 traffic in bytes (32-byte sectors), the figure SURVEY 8-d asks for beside dram__bytes
        d["lts_t_bytes (sectors x 32)"] = "%.3f Mbyte" % (float(d["lts__t_sectors.sum"].split()[0]) * 32 / 1e6)
    st = sorted(((float(r[i]), k) for i, k in enumerate(h) if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k), reverse=True)
    d["stalls_per_issue_top"] = {k.split("issue_stalled_")[1].split("_per_")[0]: round(v, 3) for v, k in st[:8]}
    allj = json.load(open(out)) if os.path.exists(out) else {}
    allj[label] = d
    json.dump(allj, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
