"""Host-side emulation of the conflict-free slot schedule of the packed up sweep (lpp_tiled.cu::schedule_conflict_free) on the 4x4\nlattice: executed slots per state / quarter-warp / warp, the edge-colouring optimum (max of the largest lane degree and the largest\nbank load of a quarter-warp), and what re-grouping the lanes by hop count would change.  Backs the numbers in profiles/README.md.\nRun: python tools/schedule_stats.py"""
import numpy as np, itertools, sys
sys.path.insert(0,'/root/repo')
from lanczosplusplus_b200 import geometry as geo
hop = geo.square(4,4,-1.0)
n=16; k=8
states=[w for w in range(1<<n) if bin(w).count('1')==k]
rank={w:i for i,w in enumerate(states)}
N=len(states)
targets=[]
for w in states:
    t=[]
    for i in range(n):
        if not (w>>i)&1: continue
        for j in range(n):
            if hop[i,j]!=0 and not (w>>j)&1:
                t.append(rank[w^(1<<i)^(1<<j)])
    targets.append(t)
cnt=np.array([len(t) for t in targets]); print('mean hops',cnt.mean(),'max',cnt.max())
def schedule(order_states, G=8):
    # order_states: list of state indices in processing order (lane assignment); positions in smem remain = state index
    out={}
    for base in range(0,N,G):
        grp=order_states[base:base+G]
        rem={u:list(targets[u]) for u in grp}
        res={u:[] for u in grp}
        while any(rem[u] for u in grp):
            order=sorted(grp,key=lambda u:-len(rem[u]))
            used=0
            for u in order:
                if not rem[u]: continue
                pick=None
                for jx,tg in enumerate(rem[u]):
                    if not (used>>(tg%G))&1: pick=jx;break
                if pick is None: res[u].append(-1); continue
                used|=1<<(rem[u][pick]%G); res[u].append(rem[u].pop(pick))
        out.update(res)
    return out
def stats(order_states, out):
    L=np.array([ (max([i for i,x in enumerate(out[u]) if x>=0])+1 if any(x>=0 for x in out[u]) else 0) for u in order_states])
    nch=(N+31)//32
    mx=np.array([L[c*32:(c+1)*32].max() for c in range(nch)])
    grp8=np.array([L[c*8:(c+1)*8].max() for c in range((N+7)//8)])
    print('  mean len per state %.2f | per-8 max %.2f | per-32 max: raw %.2f r2 %.2f r4 %.2f'%(L.mean(), grp8.mean(), mx.mean(), ((mx+1)//2*2).mean(), ((mx+3)//4*4).mean()))
ident=list(range(N))
out=schedule(ident); print('colex order'); stats(ident,out)
# optimal slots per group of 8 = max(max lane degree, max bank load)  (bipartite multigraph edge colouring, Koenig)
G=8
d=[]
for base in range(0,N,G):
    grp=ident[base:base+G]
    lane=max(len(targets[u]) for u in grp)
    bank=np.zeros(G,int)
    for u in grp:
        for t in targets[u]: bank[t%G]+=1
    d.append(max(lane,bank.max()))
d=np.array(d); print('optimal per-8 slots: mean %.2f (lane-only bound %.2f)'%(d.mean(), np.mean([max(len(targets[u]) for u in ident[b:b+G]) for b in range(0,N,G)])))
mx32=np.array([d[c*4:(c+1)*4].max() for c in range((len(d)+3)//4)]); print('optimal per-32 max %.2f r2 %.2f r4 %.2f'%(mx32.mean(), ((mx32+1)//2*2).mean(), ((mx32+3)//4*4).mean()))
def group_cost(grp):
    lane=max(len(targets[u]) for u in grp)
    bank=np.zeros(G,int)
    for u in grp:
        for t in targets[u]: bank[t%G]+=1
    return max(lane,bank.max()), lane, bank.max()
def eval_order(order, label):
    c=[group_cost(order[b:b+G]) for b in range(0,N,G)]
    c=np.array(c); 
    m32=np.array([c[i*4:(i+1)*4,0].max() for i in range((len(c)+3)//4)])
    print('%-30s per-8 mean %.2f (lane %.2f bank %.2f) per-32 max %.2f r4 %.2f'%(label,c[:,0].mean(),c[:,1].mean(),c[:,2].mean(), m32.mean(), ((m32+3)//4*4).mean()))
eval_order(ident,'colex')
bydeg=sorted(ident,key=lambda u:(len(targets[u]),u))
eval_order(bydeg,'sorted by degree')
# sorted by degree within windows of 256 colex neighbours
for W in (64,256,1024,4096):
    o=[]
    for b in range(0,N,W): o+=sorted(ident[b:b+W],key=lambda u:(len(targets[u]),u))
    eval_order(o,'degree-sorted in windows of %d'%W)
