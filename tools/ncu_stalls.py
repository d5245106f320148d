"""Read an .ncu-rep here (no GPU): headline counters + stall samples bucketed between landmark SASS instructions.
usage: python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:3]:
    d = dict(zip(hdr, r))
    for k in ("Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k in d:
            print(f"{k} = {d[k][:100]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:   # the first kernel of the report only (a second one repeats the header)
    if r and r[0] == "Address":
        break
    if len(r) >= len(hdr):
        data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        tot[s] += int(r[idx[s]] or 0)
T = sum(tot.values())
print("total samples", T)
print("  ".join(f"{s[6:]}={100 * v / T:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
marks = ("SYNCS", "LDTM", "STTM", "BAR.", "UTMA", "UTCHMMA", "UTCBAR", "SHFL", "LDS", "STS", "STG", "EXIT", "MUFU.RCP", "MUFU.LG2", "WARPSYNC")
cum = 0
start = 0
for i, r in enumerate(data):
    n = int(r[idx["# Samples"]] or 0)
    s = r[idx["Source"]].strip()
    cum += n
    if any(m in s for m in marks) and cum >= thr:
        print(f"{start:5d}-{i:5d} samples={cum:6d} ({100 * cum / T:4.1f}%)  ends at: {s[:70]}  ex={r[idx['Instructions Executed']]}")
        cum = 0
        start = i + 1
print("tail", cum)
