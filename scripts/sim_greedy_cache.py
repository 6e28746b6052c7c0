"""CPU simulation of the greedy-action cache policies of the general kernel on the C4 game (8 QTable agents, 1001x101 tables):
miss rate of greedy steps with (old) invalidate-on-write, (new) a carried (max, argmax) pair that only survives raised maxima,
(top2) the exact carry the kernel uses now (row scan excludes the column the next transition writes).  DESIGN.md 4.2 quotes
the output: old 0.908, new 0.466, top2 0.008 at epsilon ~ 0.43 (epochs 200-300).  Plain numpy; takes ~2 minutes."""
import numpy as np
rng=np.random.default_rng(0)
n,S,A,T=8,1000,101,100
lo,hi=0.05,0.15; a,b=10.0,1.0
gamma,alpha=0.95,0.1
Q=[12.5/(1-gamma)+rng.standard_normal((S+1,A)) for _ in range(n)]
eps=0.5; eps_end=0.001; eps_step=0.9995
scale=lambda k: k/(A-1.0)*(hi-lo)+lo
price=rng.uniform(0,a)
# cache state per agent/row: 0 = empty, 1 = valid; policy old: invalidate on write; new: patch rule; top2: always valid after write (if carried)
valid_old=[np.zeros(S+1,bool) for _ in range(n)]
valid_new=[np.zeros(S+1,bool) for _ in range(n)]
valid_top2=[np.zeros(S+1,bool) for _ in range(n)]
stats={k:[0,0] for k in ('old','new','top2')}
E=300
for e in range(E):
    P=[price]; acts=[]
    for t in range(T):
        row=int(np.rint(np.float32(price)/np.float32(10)*np.float32(S)))
        ks=[]
        for i in range(n):
            if rng.random()<eps: k=int(rng.integers(A))
            else:
                k=int(np.argmax(Q[i][row]))
                if e>=E-100:
                    for name,v in (('old',valid_old),('new',valid_new),('top2',valid_top2)):
                        stats[name][1]+=1
                        if not v[i][row]: stats[name][0]+=1
                for v in (valid_old,valid_new,valid_top2): v[i][row]=True
            ks.append(k)
        x=[scale(k) for k in ks]; Aq=[a/b*xx for xx in x]; Qs=sum(Aq)
        price=max(0.0,a-b*Qs); P.append(price); acts.append(ks)
    rows=[int(np.rint(p/10*S)) for p in P]
    for i in range(n):
        old=[Q[i][rows[j],acts[j][i]] for j in range(T)]
        for j in range(T):
            st,k,nx=rows[j],acts[j][i],rows[j+1]
            r=P[j+1]*(a/b*scale(k))
            m=Q[i][nx].max()
            # info about row st before write available if j>0 (carried from previous iteration)
            rm=Q[i][st].max(); ra=int(np.argmax(Q[i][st]))
            nv=(1-alpha)*old[j]+alpha*(r+gamma*m)
            known = j>0
            unknown_case = (k==ra and nv<rm)
            Q[i][st,k]=nv
            valid_old[i][st]=False
            valid_new[i][st]= known and not unknown_case
            valid_top2[i][st]= known
    eps=eps_end+(eps-eps_end)*eps_step
for k,(m,t) in stats.items(): print(k,'miss rate over greedy steps %.3f'%(m/max(t,1)), t)
print('eps',eps)
