#!/usr/bin/env python3
"""Event-driven model of balancing cfg 2 (256 four-stream groups on 148 SMs x 2 slots) through a (group, time-chunk) ready queue:
a group steps in `lone` ns when its SM partner slot idles, in `pair` ns otherwise.  Prints the effective ns/step per policy
(at best 0.98 of the paired step time -- not built)."""
import random
from collections import deque
def sim(G=256, SM=148, Tsteps=2880000//8, TJ=2048, lone=166.0, pair=204.0, ovh_ns=3000.0, timeslice=False, lifo=False):
    W = 2*SM
    prog = [0]*G
    ready = deque(range(G))
    idle = deque(range(W))      # order: 0..147 first CTAs, 148.. second on each SM (sm = w % SM)
    state = ['idle']*W; grp=[None]*W; rem=[0.0]*W; ovh=[0.0]*W; tleft=[0.0]*W
    t=0.0; finished=0
    while finished < G:
        while ready and idle:
            g = ready.popleft(); w = idle.pop() if lifo else idle.popleft()
            state[w]='run'; grp[w]=g; rem[w]=min(TJ, Tsteps-prog[g]); ovh[w]=ovh_ns; tleft[w]=TJ*pair
        # rates
        dt=1e30; rate=[0.0]*W
        for w in range(W):
            if state[w]=='run':
                p=(w+SM)%W
                per = pair if state[p]=='run' else lone
                rate[w]=1.0/per
                if ovh[w]>0: d=ovh[w]
                elif timeslice: d=min(tleft[w], rem[w]*per) if False else min(tleft[w], (Tsteps-prog[grp[w]])*per)
                else: d=rem[w]*per
                dt=min(dt,d)
        t+=dt
        fin=[]
        for w in range(W):
            if state[w]=='run':
                if ovh[w]>0:
                    ovh[w]-=dt
                    continue
                adv=rate[w]*dt
                prog[grp[w]]+=adv
                if timeslice:
                    tleft[w]-=dt
                    if tleft[w]<=1e-6 or prog[grp[w]]>=Tsteps-1e-6: fin.append(w)
                else:
                    rem[w]-=adv
                    if rem[w]<=1e-6: fin.append(w)
        for w in fin:
            g=grp[w]; state[w]='idle'; grp[w]=None; idle.append(w)
            if prog[g]>=Tsteps-1e-6: finished+=1
            else: ready.append(g)
    return t/Tsteps
for ts in (False, True):
  for lifo in (False, True):
    for TJ in (1024, 4096):
        print("timeslice" if ts else "steps", "lifo" if lifo else "fifo", TJ, sim(TJ=TJ, timeslice=ts, lifo=lifo))
