"""Summarise an .ncu-rep (ncu --set full) into a small JSON: python tools/ncu_summary.py report.ncu-rep out_prefix
Writes <out_prefix>_<kernel>.json for every kernel in the report (key launch facts, pipe utilisation, stall reasons, DRAM and
local-memory traffic)."""
import csv, json, re, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "smsp__warps_eligible.avg.per_cycle_active"]
raw = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = {"Kernel Name": r[hdr.index("Kernel Name")]}
    for i, h in enumerate(hdr):
        if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
            if r[i] not in ("", "n/a"):
                d[h] = "%s %s" % (r[i], units[i])
    stalls = {k: float(v.split()[0].replace(",", "")) for k, v in d.items() if "issue_stalled" in k}
    for k in list(d):
        if "issue_stalled" in k and stalls[k] < 0.15:
            del d[k]
    short = re.sub(r"[^A-Za-z0-9_]+", "_", re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("zkmsm::", ""))[:48].strip("_")
    out = "%s_%s.json" % (sys.argv[2], short)
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)
