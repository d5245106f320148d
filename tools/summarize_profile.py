"""Summarise gpurun_out ncu artefacts into small tracked files under profiles/ (usage: python tools/summarize_profile.py TAG)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
out_tag = sys.argv[2] if len(sys.argv) > 2 else tag
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

# ---- launch list: per-kernel totals and shares
launch_csv = os.path.join(src, f"launches_{tag}.csv")
if os.path.exists(launch_csv):
    rows = [r for r in csv.reader(open(launch_csv)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hdr_i]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[hdr_i + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("ovk::", "")
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1e3 if unit in ("nsecond", "ns") else v * 1e3 if unit in ("msecond", "ms") else v
        tot[name][0] += 1
        tot[name][1] += us
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(dst, f"{out_tag}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over one L/14@224 batch-1024 forward step + loss leg\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f}% |\n")
    print("wrote launches summary,", len(tot), "kernels")

# ---- full capture of the GEMM: key metrics per captured launch
rep = os.path.join(src, f"prof_gemm_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(dst, f"{out_tag}_gemm_ncu_full.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        traffic = []
        for r in rows[2:]:
            w.writerow([r[i][:90] for i in idx])
            rd = float(r[hdr.index("dram__bytes_read.sum")])
            wr = float(r[hdr.index("dram__bytes_write.sum")])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            traffic.append(rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]])
    json.dump({"source": f"profiles/{out_tag}_gemm_ncu_full.csv", "dram_bytes_per_launch": traffic,
               "mean_dram_bytes_per_launch": sum(traffic) / max(1, len(traffic))},
              open(os.path.join(dst, "gemm_traffic.json"), "w"), indent=1)
    print("wrote gemm ncu full csv,", len(rows) - 2, "launches")


# ---- attention / loss kernels: key metrics
rep2 = os.path.join(src, f"prof_attn_loss_{tag}.ncu-rep")
if os.path.exists(rep2):
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(dst, f"{out_tag}_attention_loss_ncu_full.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:90] for i in idx])
    print("wrote attention/loss ncu csv,", len(rows) - 2, "launches")
