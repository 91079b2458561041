"""Opcode mix of one region of a kernel -- the SASS lines of an `ncu --page source --csv` export that were executed
exactly N times (tools/ncu_regions.py lists the counts): python tools/ncu_mix.py X.source.csv N."""
import csv,sys,re
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if "# Samples" in r][0]
hdr=rows[h]; si,ii,src=hdr.index("# Samples"),hdr.index("Instructions Executed"),hdr.index("Source")
want=int(sys.argv[2])
c=Counter(); s=Counter()
for r in rows[h+1:]:
    if "# Samples" in r: break
    if len(r)>si and r[si].isdigit() and int(r[ii])==want:
        t=r[src].strip()
        t=re.sub(r"^@!?U?P\d+\s+","",t)
        op=t.split()[0].split(".")[0]
        c[op]+=1; s[op]+=int(r[si])
tot=sum(c.values()); ts=sum(s.values())
for k,v in c.most_common(): print(f"{k:10s} {v:4d}  samples {100*s[k]/max(ts,1):5.1f}%")
print("total",tot)
