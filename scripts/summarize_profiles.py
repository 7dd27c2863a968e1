"""Turns the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under
profiles/:  launch list -> per-kernel share table; full capture -> raw metric CSV + a small JSON
with DRAM traffic, tensor-pipe and issue utilisation per kernel.
usage: python scripts/summarize_profiles.py <launches.csv> <capture.ncu-rep> <round tag, e.g. r1>"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list -----------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi or r[ki] == "Kernel Name":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1][:70]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
with open(os.path.join(out_dir, f"{tag}_launch_summary.txt"), "w") as f:
    f.write(f"# ncu launch list ({os.path.basename(launches)}): 3 training steps of the bench workload "
            "(scripts/prof_step.py 4096 3), B200\n# ncu --metrics gpu__time_duration.sum --clock-control none "
            "(cold-cache, serialised: compare shares)\n        us    n  share  kernel\n")
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"{v:10.1f} {cnt[name]:4d} {100 * v / total:5.1f}%  {name}\n")
import shutil
shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches.csv"))

# ---- full capture ------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(out_dir, f"{tag}_ncu_raw_mlp_kernels.csv"), "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def get(r, key):
    for h, i in col.items():
        if h.endswith(key):
            try:
                return float(r[i].replace(",", ""))
            except ValueError:
                return None
    return None


units = rows[1]
kernels = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).split("::")[-1]
    def bytes_of(key):
        for h, i in col.items():
            if h.endswith(key):
                v = float(r[i].replace(",", ""))
                u = units[i]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return None
    kernels[name] = {
        "duration_ms_under_ncu": get(r, "gpu__time_duration.sum"),
        "dram_bytes_read": bytes_of("dram__bytes_read.sum"),
        "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
        "dram_throughput_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_active_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": get(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "lts_throughput_pct": get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": get(r, "launch__registers_per_thread"),
        "warp_instructions": get(r, "smsp__inst_executed.sum"),
    }
json.dump({"source": f"ncu --set full --clock-control none, scripts/prof_step.py 4096 2, profiles/{tag}_ncu_raw_mlp_kernels.csv",
           "kernels": kernels}, open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launch_summary.txt")).read())
print(json.dumps(kernels, indent=1))
