#!/usr/bin/env python
"""Turns a gpurun_out/<tag> directory produced by tools/gpu_profile.sh into committed evidence:
profiles/<round>_*.csv (launch lists, selected ncu raw metrics) + profiles/<round>_summary.md +
profiles/traffic.json (DRAM bytes per launch of each workload's dominant kernel, read by bench.py).
    python tools/summarize_profiles.py gpurun_out/r01p r01"""
import csv, json, os, shutil, sys

src, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
md = ["# %s: ncu evidence (B200, `--clock-control none`)\n" % rnd,
      "Source: `bash tools/gpu_profile.sh` under gpurun; raw CSV exports next to this file.",
      "Launch lists are cold-cache and serialised (compare shares, not absolutes); the bench numbers are CUDA-event timings.\n"]
traffic = {}
for w in ("cfg2", "cfg3", "cfg4"):
    lp = os.path.join(src, "launches_%s.csv" % w)
    if not os.path.exists(lp):
        continue
    shutil.copy(lp, os.path.join(dst, "%s_launches_%s.csv" % (rnd, w)))
    rows = list(csv.reader(open(lp)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = {}
    for r in rows[h + 1:]:
        name = r[4].split("(")[0].replace("void ", "").replace("b2::", "")
        a = agg.setdefault(name, {"ids": set(), "t": 0.0, "r": 0.0, "w": 0.0})
        a["ids"].add(r[0])
        v = float(r[-1].replace(",", ""))
        if r[-3] == "gpu__time_duration.sum": a["t"] += v
        if r[-3] == "dram__bytes_read.sum": a["r"] += v
        if r[-3] == "dram__bytes_write.sum": a["w"] += v
    tot = sum(a["t"] for a in agg.values())
    md.append("## %s launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`)\n" % w)
    md.append("| kernel | launches | avg us | share of step | DRAM read/launch | DRAM write/launch |\n|---|---|---|---|---|---|")
    best = None
    for k, a in agg.items():
        n = len(a["ids"])
        md.append("| `%s` | %d | %.1f | %.1f %% | %.3f GB | %.3f GB |" % (k, n, a["t"] / n / 1e3, 100 * a["t"] / tot, a["r"] / n / 1e9, a["w"] / n / 1e9))
        if best is None or a["t"] > best[1]["t"]:
            best = (k, a, n)
    traffic[w] = int((best[1]["r"] + best[1]["w"]) / best[2])
    md.append("")
    rp = os.path.join(src, "prof_%s_raw.csv" % w)
    if os.path.exists(rp):
        rows = list(csv.reader(open(rp)))
        hdr, units = rows[0], rows[1]
        out = [["metric", "unit"] + [r[hdr.index("Kernel Name")][:60] for r in rows[2:]]]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append([k, units[i]] + [r[i] for r in rows[2:]])
        with open(os.path.join(dst, "%s_ncu_full_%s.csv" % (rnd, w)), "w", newline="") as f:
            csv.writer(f).writerows(out)
        md.append("### %s `ncu --set full` (selected metrics, one column per captured kernel)\n" % w)
        md.append("| metric | unit | " + " | ".join("`%s`" % c for c in out[0][2:]) + " |")
        md.append("|---|---|" + "---|" * (len(out[0]) - 2))
        for r in out[1:]:
            md.append("| " + " | ".join(r) + " |")
        md.append("")
for w in ("cfg2", "cfg2s", "cfg3", "cfg4", "cfg1", "full", "reference"):
    bp = os.path.join(src, "bench_%s.json" % w)
    if os.path.exists(bp):
        shutil.copy(bp, os.path.join(dst, "%s_bench_%s.json" % (rnd, w)))
cp = os.path.join(src, "clocks.csv")
if os.path.exists(cp):
    lines = open(cp).read().splitlines()
    busy = [l for l in lines[1:] if l.split(",")[1].strip().split()[0].isdigit() and int(l.split(",")[1].strip().split()[0]) > 1000]
    md.append("## clocks during the bench runs (nvidia-smi -lms 200)\n")
    md.append("%d samples, %d under load; under load: %s" % (len(lines) - 1, len(busy), "; ".join(sorted(set(", ".join(x.strip() for x in l.split(",")[1:3] + l.split(",")[4:]) for l in busy)))[:600]))
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
open(os.path.join(dst, "%s_summary.md" % rnd), "w").write("\n".join(md) + "\n")
print("\n".join(md)[:3000])
