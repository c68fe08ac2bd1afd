"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

    python tools/summarize_profile.py <round tag> <launches.csv> <full.ncu-rep> ["<command the launch list was taken over>"]
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, launches_csv, rep = sys.argv[1:4]
command = sys.argv[4] if len(sys.argv) > 4 else "python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong --no-ppo-line"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and shares ------------------------------------------------------------
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 10 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += float(r[-1]) / 1e3
total_us = sum(v[1] for v in tot.values())
lines = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over "
         f"`{command}`", "",
         "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
         "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{name}` | {n} | {us:.1f} | {100 * us / total_us:.1f}% | {us / n:.1f} |")
step_us = sum(v[1] for k, v in tot.items() if k.startswith("k_step"))
ours = sum(v[1] for k, v in tot.items() if k.startswith("k_"))
lines += ["", f"`k_step` share of all device time in the run: {100 * step_us / total_us:.1f}% "
          f"(of this repo's kernels: {100 * step_us / ours:.1f}%); the rest is one-off table construction, reset and "
          "torch fills/RNG of the harness."]
open(os.path.join(out_dir, f"launches_{tag}.md"), "w").write("\n".join(lines) + "\n")

# ---- full capture of the step kernel ----------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__block_size", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
md = [f"# ncu --set full of the step kernel ({tag})", ""]
traffic = []
for r in rr[2:]:
    get = lambda k: r[hdr.index(k)] if k in hdr else None   # noqa: E731
    md += [f"## `{get('Kernel Name')}`", "", "| metric | value | unit |", "|---|---:|---|"]
    for k in want:
        if k in hdr:
            md.append(f"| {k} | {get(k)} | {units[hdr.index(k)]} |")
    stalls = sorted(((float(r[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)
    md += ["", "warp stall reasons (avg warps stalled per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]), ""]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(get("dram__bytes_read.sum")) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(get("dram__bytes_write.sum")) * scale[units[hdr.index("dram__bytes_write.sum")]]
    traffic.append(rd + wr)
open(os.path.join(out_dir, f"step_kernel_{tag}.md"), "w").write("\n".join(md) + "\n")
json.dump({"kernel": "k_step", "dram_bytes_per_launch": sum(traffic) / len(traffic), "captures": len(traffic),
           "source": os.path.basename(rep)}, open(os.path.join(out_dir, f"traffic_{tag}.json"), "w"))
print("\n".join(lines[-3:]))
print("traffic per launch", sum(traffic) / len(traffic) / 1e6, "MB")
