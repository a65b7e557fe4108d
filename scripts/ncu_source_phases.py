"""Per-phase view of an ncu --import-source report (SASS page): executed warp-instructions and stall samples between the
barriers of the kernel, plus the instructions with the most stall samples.
    python scripts/ncu_source_phases.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_samples = sum(int(r[ix["# Samples"]]) for r in body)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(f"{len(body)} SASS instructions, {tot_inst} warp-instructions executed, {tot_samples} stall samples")
phase, acc = 0, None
def flush():
    if acc and acc["n"]:
        top = sorted(((v, k) for k, v in acc["st"].items() if v), reverse=True)[:4]
        print(f"phase {acc['id']:2d} [{acc['first']}..]: sass {acc['n']:5d}  executed {acc['inst']:10d} ({100*acc['inst']/tot_inst:5.1f}%)  "
              f"samples {acc['s']:7d} ({100*acc['s']/tot_samples:5.1f}%)  " + " ".join(f"{k[6:]}={v}" for v, k in top))
for r in body:
    if acc is None:
        acc = dict(id=phase, first=r[ix["Source"]].strip()[:28], n=0, inst=0, s=0, st={c: 0 for c in stall_cols})
    acc["n"] += 1; acc["inst"] += int(r[ix["Instructions Executed"]]); acc["s"] += int(r[ix["# Samples"]])
    for c in stall_cols:
        acc["st"][c] += int(r[ix[c]] or 0)
    if "BAR.SYNC" in r[ix["Source"]]:
        flush(); phase += 1; acc = None
flush()
print("hottest instructions by stall samples:")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"  {int(r[ix['# Samples']]):6d}  {r[ix['Source']].strip()[:70]:70s} {st}")
