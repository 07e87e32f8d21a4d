#!/usr/bin/env python
"""How far ahead of its first use does ptxas place each global load of a kernel's hottest loop?

    python tools/sass_load_distance.py cardelino_b200/build/cells_sell.o Li16EtLb1ELi896ELi1ELb0ELb1

For every LDG in the DADD loop of the function whose mangled name contains the substring: the number of instructions
and of covered entries (DMULs) between the load and the first instruction that reads its destination registers.
Round 2 used it to see that the look-ahead written in cells_sell.cu is not the one that runs: ptxas sinks the loads of
one of the two group buffers to a few instructions before their use (register budget).  Needs no GPU."""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
lines = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout.split("\n")
start=[i for i,l in enumerate(lines) if 'Function' in l and pat in l][0]
end=next((i for i in range(start+1,len(lines)) if 'Function' in lines[i]),len(lines))
ins=[]
for l in lines[start:end]:
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
a2i={a:i for i,(a,_) in enumerate(ins)}
best=None
for i,(a,t) in enumerate(ins):
    m=re.search(r'BRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?0x([0-9a-f]+)',t)
    if m and int(m.group(1),16)<a and int(m.group(1),16) in a2i:
        body=ins[a2i[int(m.group(1),16)]:i+1]
        nd=sum('DADD' in x for _,x in body)
        if nd>=50 and (best is None or nd/len(body)>best[0]): best=(nd/len(body),body)
body=best[1]
n=len(body)
# loop twice to handle wraparound
seq=body+body
for i,(a,t) in enumerate(body):
    if t.startswith('LDG'):
        r=int(re.search(r'LDG\S*\s+R(\d+)',t).group(1))
        regs=[f'R{r+k}' for k in range(4)]
        dm=0
        for j in range(i+1,len(seq)):
            tt=seq[j][1]
            if 'DMUL' in tt: dm+=1
            ops=tt.split(None,1)[1] if ' ' in tt else ''
            srcs=ops.split(',',1)[1] if ',' in ops else ''
            if any(re.search(r'\b'+x+r'\b(\.|\s|,|$)',srcs) for x in regs) and not tt.startswith('LDG'):
                print('%05x %-40s first use after %3d instr, %2d entries: %s'%(a,t[:40],j-i,dm,tt[:50])); break
