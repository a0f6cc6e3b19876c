#!/usr/bin/env python
"""CPU-side evidence for the run-compressed back-projector (IONO_BP_RUNS=1): on the LOFAR-like benchmark
geometry, how many consecutive time steps does a ray of one (antenna, direction) keep touching the same
voxel?  That is the run length of consecutive ray numbers in a voxel's entry list (internal ray order:
time fastest).  Uses the oracle for the cell search; a few (antenna, direction) pairs, all 100 times.

    python tools/run_length_stats.py        # -> overall avg run length 8.6, 8.7 bytes/entry
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ionotomo_b200.ionosphere import synthetic as S
from oracle import ionotomo_oracle as O
Na,Nt,Nd,nx,ny,nz=62,100,200,256,256,128
ants=S.lofar_stations_enu_km()[:Na]
dirs_t=S.track_directions(S.directions_in_fov(Nd,4.,1234),Nt)
xv,yv,zv=S.tight_axes(ants,dirs_t,nx,ny,nz,1000.)
print('dx',xv[1]-xv[0],yv[1]-yv[0],zv[1]-zv[0])
rng=np.random.RandomState(0)
tot_entries=0; tot_runs=0; hist=np.zeros(101,int)
for trial in range(12):
    a=rng.randint(Na); d=rng.randint(Nd)
    o=np.broadcast_to(ants[a],(Nt,3)); dr=dirs_t[:,d]
    rays=O.cast_ray(o,dr,1000.,nz)   # (Nt,4,Ns)
    sets=[]
    for t in range(Nt):
        ix,_=O.find_indices(xv,rays[t,0]); iy,_=O.find_indices(yv,rays[t,1]); iz,_=O.find_indices(zv,rays[t,2])
        vox=set()
        for cx in (0,1):
            for cy in (0,1):
                for cz in (0,1):
                    vox.update((((ix+cx)*ny+(iy+cy))*nz+(iz+cz)).tolist())
        sets.append(vox)
    # runs over t per voxel
    allv=set().union(*sets)
    ent=sum(len(s) for s in sets)
    runs=0
    for t in range(Nt):
        prev=sets[t-1] if t>0 else set()
        runs+=len(sets[t]-prev)
    tot_entries+=ent; tot_runs+=runs
    print(a,d,'entries/ray',ent/Nt,'runs',runs,'avg run',ent/runs)
print('overall avg run length',tot_entries/tot_runs,'bytes/entry at 8+6/len:',8+6/(tot_entries/tot_runs))
