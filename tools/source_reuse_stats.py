"""How many distinct source rows the down sweep needs per output row when R consecutive down states are processed together (4x4 lattice):\nthe streaming kernel loads 18.07 rows per output row (17.07 hops + its own); 16 consecutive rows share enough sources for 13.87, the\nruns of the staged kernel (states sharing the high 8 sites) 11.67.  Run: python tools/source_reuse_stats.py"""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/tools/schedule_stats.py').read().split("def schedule(")[0])
tot=sum(len(t) for t in targets)
for R in (4,8,16,32,64):
    U=0
    for b in range(0,N,R):
        s=set()
        for u in range(b,min(b+R,N)):
            s.update(targets[u]); s.add(u)
        U+=len(s)
    print('rows per group',R,'distinct sources+own per output row %.2f (of %.2f)'%(U/N, tot/N+1))
# best grouping by hop-closedness: states sharing the high 8 bits (a-blocks)
blocks={}
for i,w in enumerate(states): blocks.setdefault(w>>8,[]).append(i)
U=0
for k,v in blocks.items():
    s=set(v)
    for u in v: s.update(targets[u])
    U+=len(s)
print('a-blocks (runs sharing the high 8 sites): %.2f per row, max block %d'%(U/N, max(len(v) for v in blocks.values())))
