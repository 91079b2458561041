"""Executed warp instructions of a kernel grouped by execution count (= by loop nest), from an `ncu --page source --csv`
export (tools/ncu_source.sh): python tools/ncu_regions.py X.source.csv [kernels].  Tells which region of the SASS the
instruction stream is spent in, with its share of the stall samples."""
import csv,sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
# the source page may hold several kernels back to back: split on header rows
hdr_idx=[i for i,r in enumerate(rows) if "# Samples" in r]
for h in hdr_idx[:int(sys.argv[2]) if len(sys.argv)>2 else 1]:
    hdr=rows[h]
    si,ii,src=hdr.index("# Samples"),hdr.index("Instructions Executed"),hdr.index("Source")
    body=[]
    for r in rows[h+1:]:
        if "# Samples" in r: break
        if len(r)>si and r[si].isdigit(): body.append(r)
    tot=sum(int(r[ii]) for r in body); ts=sum(int(r[si]) for r in body)
    print("kernel block: lines",len(body),"warp-instr",tot,"samples",ts)
    c=Counter(); cs=Counter()
    for r in body: c[int(r[ii])]+=1; cs[int(r[ii])]+=int(r[si])
    for k,v in sorted(c.items(), key=lambda kv:-kv[0]*kv[1])[:14]:
        print(f"  exec {k:>9d} x {v:4d} lines = {k*v/1e6:7.2f} M ({100*k*v/tot:4.1f}%)  samples {100*cs[k]/ts:4.1f}%")
