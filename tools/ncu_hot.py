"""Top stall-sample SASS lines from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:...`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
si, ii, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
toti = sum(int(r[ii]) for r in body)
print(f"instructions {len(body)}, samples {tot}, warp-instructions executed {toti}")
order = sorted(range(len(body)), key=lambda k: -int(body[k][si]))[:n]
for k in sorted(order):
    r = body[k]
    print(f"{k:5d} {100*int(r[si])/tot:5.1f}% exec {int(r[ii]):>9d}  {r[src].strip()[:110]}")
