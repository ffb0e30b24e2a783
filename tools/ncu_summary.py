"""Summarise an .ncu-rep (ncu --set full) as a markdown table of the metrics DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name regex] > profiles/x.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__inst_executed_op_global_red.sum", "global RED instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if len(sys.argv) > 2:
        import re
        data = [r for r in data if re.search(sys.argv[2], r[hdr.index("Kernel Name")])]
    names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")[-60:] for r in data]
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in METRICS:
        if key in hdr:
            i = hdr.index(key)
            print(f"| {label} ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("\nWarp stall reasons (warps stalled per issued instruction; > 0.2 shown):\n")
    print("| stall | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for h, i in stall:
        vals = [float(r[i].replace(",", "")) for r in data]
        if max(vals) > 0.2:
            print("| " + h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") + " | " +
                  " | ".join(f"{v:.2f}" for v in vals) + " |")


if __name__ == "__main__":
    main()
