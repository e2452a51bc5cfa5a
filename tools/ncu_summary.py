"""Key metrics of every kernel in an .ncu-rep (ncu --set full) as a markdown table: `python tools/ncu_summary.py REPORT...`.

Reads the report with `ncu -i REPORT --page raw --csv`; nothing here runs on the GPU.  Used for the tables under profiles/.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads active per instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (smem)"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out.append(f"### `{name[:110]}`  ({path.split('/')[-1]})\n")
        out.append("| metric | value |\n|---|---|")
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                out.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALLS):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            out.append("| warp stalls per issue-active cycle (top 6) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " |")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summarise(p))
