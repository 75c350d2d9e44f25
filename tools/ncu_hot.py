"""Top stall-sample SASS lines of a report: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > iex and r[ismp].isdigit()]
tot = sum(int(r[ismp]) for r in body); totx = sum(int(r[iex]) for r in body)
print(f"total samples {tot}, instructions executed {totx}, sass lines {len(body)}")
# per-line listing with index for context
idx = {id(r): i for i, r in enumerate(body)}
top = sorted(body, key=lambda r: -int(r[ismp]))[:N]
for r in top:
    i = idx[id(r)]
    print(f"{i:5d} {100*int(r[ismp])/max(tot,1):5.1f}%  exec {int(r[iex]):>10d}  {r[isrc].strip()}")
# instruction mix
mix = {}
for r in body:
    op = r[isrc].strip().split()[0] if not r[isrc].strip().startswith('@') else r[isrc].strip().split()[1]
    op = op.split('.')[0]
    mix[op] = mix.get(op, 0) + int(r[iex])
print("instruction mix (executed):")
for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:25]:
    print(f"  {k:10s} {v:>12d} {100*v/max(totx,1):5.1f}%")
