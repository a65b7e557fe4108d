"""Distil `ncu --set full` reports into profiles/r02_ncu_current.json, stamped with the source hash of csrc/, which is what
bench.py quotes for `roofline.traffic` and the integer-pipe percentages (and withholds when the sources have changed since).

    python scripts/ncu_to_json.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]
Run here (no GPU needed) right after the capture, with the tree in the state that was profiled."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

M = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
     "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
     "smsp__inst_executed.sum": "warp_instructions", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
     "launch__registers_per_thread": "registers", "launch__grid_size": "grid"}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}

out_path = os.path.join(ROOT, "profiles", "r02_ncu_current.json")
kernels = {}
if os.path.exists(out_path):
    old = json.load(open(out_path))
    if old.get("csrc_sha") == bench.csrc_hash():
        kernels = old.get("kernels", {})          # same build: reports add up
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        m = re.search(r"(\w+(?:<[^>]*>)?)\s*\(", name)
        key = m.group(1) if m else name
        d = {"report": os.path.basename(rep)}
        for k, short in M.items():
            if k in ix:
                v = float(r[ix[k]].replace(",", ""))
                d[short] = v * SCALE.get(units[ix[k]], 1.0)
        d["traffic"] = d.get("dram_read", 0.0) + d.get("dram_write", 0.0)
        kernels[key] = d
json.dump({"csrc_sha": bench.csrc_hash(), "note": "one launch at 2^20 envs each; bytes per launch, duration in us (under ncu: cold, serialised)",
           "kernels": kernels}, open(out_path, "w"), indent=1)
print(json.dumps(kernels, indent=1))
