"""Stall samples and executed instructions per CUDA source line: python tools/ncu_lines.py report.ncu-rep [N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ismp = hdr.index("Warp Stall Sampling (All Samples)"); iex = hdr.index("Instructions Executed")
isrc = 1
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.OrderedDict()
cur = None
for r in rows[hi + 1:]:
    if len(r) <= iex:
        continue
    if r[0].strip():
        cur = (r[0].strip(), r[isrc].strip())
        agg.setdefault(cur, [0, 0, collections.Counter()])
    if cur is None or not r[ismp].strip().isdigit():
        continue
    a = agg[cur]
    a[0] += int(r[ismp]); a[1] += int(r[iex] or 0)
    for i in stall_cols:
        if r[i].strip().isdigit():
            a[2][hdr[i]] += int(r[i])
tot = sum(a[0] for a in agg.values()); totx = sum(a[1] for a in agg.values())
print(f"total samples {tot}, instructions executed {totx}")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    top = ", ".join(f"{k[6:]} {100*v/max(a[0],1):.0f}%" for k, v in a[2].most_common(3))
    print(f"{ln:>5} {100*a[0]/max(tot,1):5.1f}%  exec {100*a[1]/max(totx,1):5.1f}%  [{top}]  {src[:110]}")
