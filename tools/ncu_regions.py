"""Stall samples / executed instructions of line ranges of one source file:
python tools/ncu_regions.py report.ncu-rep name=lo-hi [name=lo-hi ...]   (lines of the kernel's .cu file)"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
ranges = collections.OrderedDict()
for a in sys.argv[2:]:
    k, v = a.split("="); lo, hi = v.split("-"); ranges[k] = (int(lo), int(hi))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = []   # (file name, {line: [samples, exec, thread_exec]})
cur = None; agg = None; hdr = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; agg = collections.OrderedDict(); files.append(agg); cur = None
        ismp = hdr.index("Warp Stall Sampling (All Samples)"); iex = hdr.index("Instructions Executed")
        ith = hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
        continue
    if hdr is None or len(r) <= iex: continue
    if r[0].strip().isdigit(): cur = int(r[0]); agg.setdefault(cur, [0, 0, 0])
    if cur is None or not r[ismp].strip().isdigit(): continue
    a = agg[cur]; a[0] += int(r[ismp]); a[1] += int(r[iex] or 0)
    if ith is not None and r[ith].strip().isdigit(): a[2] += int(r[ith])
tot = sum(a[0] for f in files for a in f.values()); totx = sum(a[1] for f in files for a in f.values())
print("files", len(files), "total samples", tot, "exec", totx)
# the kernel's own file = the one with the most executed instructions
main = max(files, key=lambda f: sum(a[1] for a in f.values()))
for f in files:
    if f is not main:
        print(f"  other file: samples {100*sum(a[0] for a in f.values())/tot:5.1f}%  exec {100*sum(a[1] for a in f.values())/totx:5.1f}%")
for name, (lo, hi) in ranges.items():
    s = sum(a[0] for l, a in main.items() if lo <= l <= hi); x = sum(a[1] for l, a in main.items() if lo <= l <= hi); th = sum(a[2] for l, a in main.items() if lo <= l <= hi)
    print(f"  {name:24s} {lo:4d}-{hi:4d}: samples {100*s/tot:5.1f}%  exec {100*x/totx:5.1f}% ({x/1e6:7.0f}M)  lanes/inst {th/max(x,1):4.1f}")
