#!/usr/bin/env python
"""Dense first-accept pass end to end over the number of row chunks, in the three outcomes: a
candidate accepted (download overlapped), nothing accepted (nothing downloaded), and a late
surprise (an earlier candidate accepted only by the last data set: the speculative download is
repeated).   python tools/r2_chunks.py"""
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massivedatans_b200 import synth
from massivedatans_b200.likelihood import ResidentDataset
x,y,_=synth.horns(1000000, nx=200, legacy=False, seed=1000)
ds=ResidentDataset(x,y); K=16
pts=synth.parameter_points(K, seed=7)
L=ds.loglike_batch(pts,None,0.01).copy()
srt=numpy.sort(L,axis=0); top,second=srt[-1],srt[-2]
order=numpy.argsort(numpy.bincount(numpy.argmax(L,axis=0),minlength=K),kind="stable")
pts=numpy.ascontiguousarray(pts[order]); L=L[order]
srt=numpy.sort(L,axis=0); top,second=srt[-1],srt[-2]
sure=(numpy.argmax(L,axis=0)==K-1)&(top-second>1e-6*numpy.abs(top))
far=top+1e-6*numpy.abs(top)+1e-6
Lm=numpy.where(sure,0.5*(top+second),far)
# a late surprise: candidate 3 accepted only by the very last data set
Ls=far.copy(); Ls[-1]=L[3,-1]-1e-6*abs(L[3,-1]); Ls[0]=L[9,0]-1e-6*abs(L[9,0])
for c in (1,2,3,4):
    ds.set_draw_chunks(c)
    for name,th in (('accepted',Lm),('rejected',far),('surprise',Ls)):
        ds.begin_draw(None,th)
        for _ in range(5): r=ds.draw_batch(pts,0.01)
        t0=time.perf_counter()
        for _ in range(100): r=ds.draw_batch(pts,0.01)
        dt=(time.perf_counter()-t0)/100
        want=(L>th).sum(axis=1); wk=int(numpy.nonzero(want)[0][0]) if want.any() else -1
        ok = r[0]==wk and numpy.array_equal(r[2],want) and (wk<0 or numpy.allclose(r[1],L[wk],rtol=1e-12,atol=0))
        print("chunks",c,name,"ms",round(dt*1e3,4),"k",r[0],ok, flush=True)
