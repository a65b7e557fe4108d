"""Where the warps wait: per-SASS-instruction stall samples of one kernel in an ncu report (source page) joined with the
`nvdisasm -g` line table of the built library: totals per stall reason, samples per source line, and the instructions with the
most samples.
    python scripts/ncu_stalls.py gpurun_out/x.ncu-rep <mangled-name-substring> [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lib = os.path.join(ROOT, "lib", "libtetris_piclim_sm100.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "piclim_kernels.sm_100a.cubin", lib], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, "piclim_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout
on, cur, lines = False, None, []
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
assert len(body) == len(lines), (len(body), len(lines))
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
allsmp = sum(int(r[ix["# Samples"]] or 0) for r in body)
tot = {h: sum(int(r[ix[h]] or 0) for r in body) for h in names}
print(f"{len(body)} instructions, {allsmp} samples")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {h:28s} {v:8d}  {100*v/allsmp:5.1f}%")
src = {}
def text(fl):
    if not fl: return ""
    f, ln = fl
    if f not in src:
        import glob
        path = glob.glob(os.path.join(ROOT, "*_b200", "csrc", f))
        src[f] = open(path[0]).read().splitlines() if path else []
    return src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
by_line = collections.Counter(); ex_line = collections.Counter()
for r, fl in zip(body, lines):
    by_line[fl] += int(r[ix["# Samples"]] or 0); ex_line[fl] += int(r[ix["Instructions Executed"]] or 0)
print("--- samples per source line")
for fl, v in by_line.most_common(topn):
    print(f"  {100*v/allsmp:5.1f}% {v:5d} smp {ex_line[fl]:9d} exec  {fl[0] if fl else '?'}:{fl[1] if fl else 0}  {text(fl)}")
print("--- instructions with the most samples")
base = int(body[0][ix["Address"]], 16)
for k in sorted(sorted(range(len(body)), key=lambda k: -int(body[k][ix["# Samples"]] or 0))[:topn]):
    r = body[k]
    st = {h[6:]: int(r[ix[h]] or 0) for h in names if int(r[ix[h]] or 0) > 3}
    fl = lines[k]
    print(f"  {int(r[ix['Address']],16)-base:#06x} {int(r[ix['# Samples']] or 0):4d} {r[ix['Source']][:52]:52s} {fl[0] if fl else '?'}:{fl[1] if fl else 0} {st}")
