"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python tools/launch_summary.py profiles/x_launches.csv > profiles/x_launches_summary.md
Times under the profiler are cold-cache and serialised: read the SHARES, not the absolutes."""
import csv, sys, collections

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    full = r["Kernel Name"]
    own = "lsx::" in full
    name = full.split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0, own])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) * 1e-6  # ns -> ms
total = sum(a[1] for a in agg.values())
# the micro-benchmarks of the peaks live in the same library but are not part of the operator
bench_only = ("ex2_kernel", "ffma_kernel", "red_kernel")
own_total = sum(a[1] for n, a in agg.items() if a[2] and not n.endswith(bench_only))
print(f"# Launch list summary of `{path.split('/')[-1]}` (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
print("Cold-cache, serialised times: use the SHARES. The list includes the in-process reference leg, the peak")
print("micro-benchmarks and torch's own small kernels.")
print(f"The new operator's own kernels sum to {own_total:.1f} ms of the {total:.1f} ms; `share of own` is relative to that sum.\n")
print("| launches | total ms | share | share of own | avg us | kernel |")
print("|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    is_own = a[2] and not n.endswith(bench_only)
    so = f"{100 * a[1] / own_total:.1f}%" if is_own else "-"
    print(f"| {a[0]} | {a[1]:.3f} | {100 * a[1] / total:.1f}% | {so} | {1e3 * a[1] / a[0]:.1f} | `{n[-70:]}` |")
