#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_v3.csv profiles/r01_launches_bench_v3 "<note>"
    python scripts/summarize_ncu.py full gpurun_out/apply_v3_full.ncu-rep profiles/r01_apply_v3_full "<note>" [tet:119]

`launches`: copies the per-launch csv (gpu__time_duration.sum, --clock-control none) and writes a
per-kernel share table. `full`: extracts the metrics the roofline argument uses from a
`ncu --set full` report (read here, without a GPU) and writes profiles/apply_traffic.json, which
bench.py reports as roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def launches(src, dst, note):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    i_unit = hdr.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        v = float(r[i_val].replace(",", ""))
        v = v / 1e3 if r[i_unit] in ("ns", "nsecond") else v
        agg[r[i_name]].append(v)
    total = sum(sum(v) for v in agg.values())
    with open(dst + ".csv", "w") as f:
        f.write(open(src).read())
    with open(dst + "_summary.txt", "w") as f:
        f.write(f"# {note}\n# cold-cache, serialised (ncu): compare SHARES, not absolutes\n")
        f.write(f"# total {total:.1f} us over {sum(len(v) for v in agg.values())} launches\n\n")
        for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{100 * sum(v) / total:6.2f}%  {len(v):4d} launches  avg {sum(v) / len(v):9.2f} us  {name[:150]}\n")
    print(open(dst + "_summary.txt").read())


def full(src, dst, note):
    # src: the report, or its raw page already exported on the GPU box (`ncu -i rep --page raw --csv > raw.csv`: the
    # report itself can exceed what a gpurun call brings back)
    out = open(src).read() if src.endswith(".csv") else \
        subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    i_name = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in FULL_METRICS if m in hdr]
    lines, traffic = [f"# {note}", "# ncu --set full --clock-control none (per launch)"], []
    for r in data:
        lines.append(f"\n{r[i_name]}")
        vals = {}
        for m, i in cols:
            lines.append(f"  {m:80s} {r[i]:>16s} {units[i]}")
            vals[m] = (r[i], units[i])
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * scale[vals["dram__bytes_read.sum"][1]]
        wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * scale[vals["dram__bytes_write.sum"][1]]
        traffic.append({"kernel": r[i_name][:120], "dram_bytes": rd + wr,
                        "duration_us": float(vals["gpu__time_duration.sum"][0].replace(",", ""))})
    open(dst + "_raw.txt", "w").write("\n".join(lines) + "\n")
    # per apply variant (the epilogue is part of the kernel name): what bench.py reports as roofline.traffic
    variants = {}
    for epi in ("EpiUY", "EpiYYandYX", "EpiXY", "NoEpi", "EpiResidual"):
        v = [t["dram_bytes"] for t in traffic if "apply_kernel_tma" in t["kernel"] and epi + ">" in t["kernel"]]
        if v:
            variants[epi] = sum(v) / len(v)
    cell, axis = (sys.argv[5].split(":") + ["119"])[:2] if len(sys.argv) > 5 else ("tet", "119")
    json.dump({"workload": {"cell": cell, "axis": int(axis)}, "variants": variants, "source": os.path.basename(dst) + "_raw.txt",
               "note": note + "; dram__bytes_read.sum + dram__bytes_write.sum per launch", "launches": traffic},
              open(os.path.join(os.path.dirname(dst), "apply_traffic.json"), "w"), indent=1)
    print("\n".join(lines))
    print("dram bytes per launch, per apply variant:", variants)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
