"""Per-launch table (one row per captured kernel launch) of an .ncu-rep: python tools/ncu_launch_table.py x.ncu-rep > profiles/x.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("gpu__time_duration.sum","us"),("launch__grid_size","grid"),("launch__block_size","block"),("launch__registers_per_thread","regs"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active","issue %"),("dram__bytes_read.sum","DRAM rd"),("dram__bytes_write.sum","DRAM wr"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","DRAM %"),("lts__t_sector_hit_rate.pct","L2 hit %"),("sm__warps_active.avg.pct_of_peak_sustained_active","occ %")]
print(f"# ncu --set full, every captured kernel (`{rep.split('/')[-1]}`), in launch order\n")
print("| # | kernel | " + " | ".join(f"{l} ({units[hdr.index(k)]})" if l.startswith('DRAM r') or l.startswith('DRAM w') else l for k,l in cols) + " |")
print("|---|---|" + "---|"*len(cols))
for n, r in enumerate(data):
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ","").split("::")[-1]
    vals=[]
    for k,l in cols:
        v = r[hdr.index(k)].replace(",","")
        try:
            f=float(v); v = f"{f:.0f}" if f==int(f) else f"{f:.1f}"
        except: pass
        vals.append(v)
    print(f"| {n} | {name} | " + " | ".join(vals) + " |")
