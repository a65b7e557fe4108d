"""Executed warp-instructions and stall samples per source line: joins the SASS page of an ncu report (in kernel order)
with `nvdisasm -g` line information of the same build, then buckets lines by the function they sit in.
    python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep step_observe_kernelILi0"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
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
# function buckets: the nearest preceding line that looks like a function header
src = {}
def bucket(f, ln):
    if f not in src:
        path = glob.glob(os.path.join(ROOT, "*_b200", "csrc", f)) or glob.glob(os.path.join(ROOT, "**", f), recursive=True)
        src[f] = open(path[0]).read().splitlines() if path else []
    s = src[f]
    for k in range(min(ln, len(s)) - 1, -1, -1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline|__host__|[a-zA-Z_].*\s)\S*?\b(\w+)\s*\(", s[k])
        if m and not s[k].startswith((" ", "\t", "//", "#")):
            return m.group(1)
    return "?"
by_fn, by_line = collections.Counter(), collections.Counter()
smp_fn = collections.Counter()
tot = 0
for r, fl in zip(body, lines):
    ex, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    tot += ex
    b = bucket(*fl) if fl else "?"
    by_fn[(fl[0] if fl else "?", b)] += ex; smp_fn[(fl[0] if fl else "?", b)] += sm
    by_line[fl] += ex
print(f"total executed warp-instructions {tot}")
for (f, b), v in by_fn.most_common(25):
    print(f"  {100*v/tot:5.1f}%  {v:10d}  samples {smp_fn[(f,b)]:6d}  {f}:{b}")
print("top lines:")
for fl, v in by_line.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    f, ln = fl
    text = src.get(f, [""] * ln)[ln - 1].strip()[:110] if fl and src.get(f) else ""
    print(f"  {100*v/tot:5.1f}%  {f}:{ln}  {text}")
