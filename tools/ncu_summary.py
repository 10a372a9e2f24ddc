"""Condenses an `ncu --page raw --csv` export into the per-kernel summary kept under profiles/.
   python tools/ncu_summary.py gpurun_out/raw_TAG.csv profiles/TAG_ncu_full_summary.json [note]"""
import csv, json, sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stall = [(i, h.split("issue_stalled_")[1].split("_per_warp_active")[0]) for i, h in enumerate(hdr)
             if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")] or \
            [(i, h.split("issue_stalled_")[1].split("_per")[0]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h]
    out, seen = [], {}
    for r in data:
        name = r[hdr.index("Kernel Name")]
        key = (name, r[hdr.index("Grid Size")])
        if seen.get(key, 0) >= 2:   # at most two launches per (kernel, grid)
            continue
        seen[key] = seen.get(key, 0) + 1
        e = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                e[k] = (r[i] + " " + units[i]).strip()
        top = sorted(((float(r[i] or 0), n) for i, n in stall), reverse=True)[:4]
        e["top_stalls_cycles_per_issue"] = {n: round(v, 2) for v, n in top}
        out.append(e)
    if len(sys.argv) > 3:
        out.append({"_note": sys.argv[3]})
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(f"{len(out)} entries -> {sys.argv[2]}")

if __name__ == "__main__":
    main()
