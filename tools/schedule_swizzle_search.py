"""Companion of tools/schedule_stats.py: how far the slot count of the packed up sweep could drop if the position of a state inside\nits aligned block of 8 tile slots were chosen freely (an XOR constant per block keeps the tile fill coalesced), with and without\nsorting the lanes of a 32-state chunk by hop count.  Local search, a few minutes of CPU.  Result on the 4x4 lattice: per-quarter-warp\nslots 20.85 -> 19.94 (colex lanes) and 22.36 -> 20.13 (lanes sorted inside chunks); lane-degree bounds 19.71 and 18.09."""
import numpy as np, sys, random
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/tools/schedule_stats.py').read().split("def schedule(")[0])   # states, targets, N
G=8
deg=np.array([len(t) for t in targets])
# lane grouping options within chunks of 32 (keeps x/y accesses of a warp inside one 256-byte segment)
def groups_sorted_in_chunks(C=32):
    order=[]
    for b in range(0,N,C):
        ch=list(range(b,min(b+C,N)))
        ch.sort(key=lambda u:(-deg[u],u))
        order+=ch
    return order
def lane_bound(order):
    return np.mean([max(deg[u] for u in order[b:b+G]) for b in range(0,N,G)])
ident=list(range(N))
print('lane-only bound colex', lane_bound(ident), ' sorted within 32-chunks', lane_bound(groups_sorted_in_chunks(32)))
def cost(order, bank):
    tot=0
    for b in range(0,N,G):
        grp=order[b:b+G]
        lane=max(deg[u] for u in grp)
        bl=np.zeros(G,int)
        for u in grp:
            for t in targets[u]: bl[bank[t]]+=1
        tot+=max(lane,bl.max())
    return tot/((N+G-1)//G)
bank0=np.array([u%G for u in range(N)])
for name,order in (('colex',ident),('sorted32',groups_sorted_in_chunks(32))):
    print(name,'natural banks', cost(order,bank0))
    # local search over per-block XOR constants
    nblk=(N+7)//8
    c=np.zeros(nblk,int)
    # incremental structures: for each group, bank loads
    grp_of={}
    groups=[order[b:b+G] for b in range(0,N,G)]
    # users[t] = list of groups that have t as a target (with multiplicity)
    users=[[] for _ in range(N)]
    for gi,grp in enumerate(groups):
        for u in grp:
            for t in targets[u]: users[t].append(gi)
    lane=np.array([max(deg[u] for u in grp) for grp in groups])
    loads=np.zeros((len(groups),G),int)
    bank=bank0.copy()
    for gi,grp in enumerate(groups):
        for u in grp:
            for t in targets[u]: loads[gi,bank[t]]+=1
    def total(): return np.maximum(lane,loads.max(axis=1)).sum()
    cur=total(); rng=random.Random(1)
    for it in range(60000):
        b=rng.randrange(nblk); newc=rng.randrange(8)
        if newc==c[b]: continue
        members=[u for u in range(8*b,min(8*b+8,N))]
        touched=set()
        for t in members:
            for gi in users[t]: touched.add(gi)
        before=sum(max(lane[gi],loads[gi].max()) for gi in touched)
        # apply
        for t in members:
            ob=bank[t]; nbk=(t ^ newc)&7
            for gi in users[t]: loads[gi,ob]-=1; loads[gi,nbk]+=1
            bank[t]=nbk
        after=sum(max(lane[gi],loads[gi].max()) for gi in touched)
        if after<=before: c[b]=newc; cur+=after-before
        else:
            for t in members:
                ob=bank[t]; nbk=(t ^ c[b])&7
                for gi in users[t]: loads[gi,ob]-=1; loads[gi,nbk]+=1
                bank[t]=nbk
    print(name,'after XOR-swizzle search', cur/len(groups))
