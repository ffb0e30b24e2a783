"""Summarise an .ncu-rep (ncu --set full) as a markdown table of the metrics DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name regex] [--json profiles/ncu_render_kernels.json] > profiles/x.md

--json also writes the record bench.py's `roofline.traffic` / `issue_slots_pct` are read from: per render kernel the dram bytes
and issue-slot utilisation of this capture, stamped with the sha1 of the kernel sources in the tree (bench.py reports the
record as stale, traffic = null, as soon as those sources change)."""
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


def _num(x):
    return float(x.replace(",", ""))


def write_json(path, rep, hdr, units, data):
    import hashlib
    import json
    import os
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha1()
    for name in ("render_fwd.cu", "render_bwd.cu", "tile_stage.cuh"):
        with open(os.path.join(repo, "langscene-x_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rec = {"capture": os.path.basename(rep), "sources_sha1": h.hexdigest()}
    for r in data:
        kn = r[hdr.index("Kernel Name")]
        key = "render_fwd" if "render_fwd_kernel" in kn else ("render_bwd" if "render_bwd_kernel" in kn else None)
        if key is None or key in rec:
            continue
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        rec[key] = {"kernel": kn.split("(")[0].replace("void ", ""),
                    "dram_bytes": int(_num(r[ir]) * scale.get(units[ir], 1) + _num(r[iw]) * scale.get(units[iw], 1)),
                    "issue_active_pct": _num(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                    "duration_ms_under_ncu": _num(r[hdr.index("gpu__time_duration.sum")]),
                    "warp_instructions": _num(r[hdr.index("smsp__inst_executed.sum")]),
                    "shared_wavefronts": _num(r[hdr.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")])}
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)


def main():
    json_out = None
    if "--json" in sys.argv:
        i = sys.argv.index("--json")
        json_out = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if len(sys.argv) > 2:
        import re
        data = [r for r in data if re.search(sys.argv[2], r[hdr.index("Kernel Name")])]
    if json_out:
        write_json(json_out, rep, hdr, units, data)
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
