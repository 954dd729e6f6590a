"""Feasibility study (CPU, numpy; analysis only): how much of the vote volume could an EXACT branch-and-bound peak search skip?
Only the peak voxel and its count leave Accumulator_3D (AccumulatorSpace.py:406-419).  An upper bound per block of b^3 voxels --
the number of points whose shell (R - f, R) can touch the block -- lets every block whose bound is below the best exact count be
discarded without voting into it.  Prints, per (frame, block size): voxel votes, block "votes" of the bound pass, blocks whose
bound reaches the true peak, and the (point, block) pairs an exact pass over those blocks would still rasterise.
Results and the caveats are in DESIGN.md section 10.  Uses the oracle: test/analysis infrastructure, not product code."""
import sys, numpy as np, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from rcvpose_b200 import synth
f=(3**0.5)/4
def study(fr, kp, b):
    xyz,rl=synth.frame_to_points(fr["K"],fr["depth"],fr["radius"][kp])
    pre=oracle.prelude(xyz,rl); p,R,D=pre["p"],pre["R"],pre["D"]
    _,info=oracle.Accumulator_3D(xyz,rl,return_info=True,return_volume=True)
    vol=info["volume"]; M=info["peak"]; votes=info["votes"]
    nb=(D+b-1)//b
    lo=np.arange(nb)*b; hi=np.minimum(lo+b-1,D-1)
    bound=np.zeros((nb,nb,nb),np.int32); touched=0
    keep=R>0
    P=p[keep]; Rr=R[keep].astype(np.float64)
    for ch in range(0,len(P),256):
        q=P[ch:ch+256]; r=Rr[ch:ch+256]
        # per-axis min and max distance to each block interval
        dmin=[np.maximum(0,np.maximum(lo[None,:]-q[:,a,None], q[:,a,None]-hi[None,:])) for a in range(3)]
        dmax=[np.maximum(np.abs(q[:,a,None]-lo[None,:]), np.abs(q[:,a,None]-hi[None,:])) for a in range(3)]
        dmin2=dmin[0][:,:,None,None]**2+dmin[1][:,None,:,None]**2+dmin[2][:,None,None,:]**2
        dmax2=dmax[0][:,:,None,None]**2+dmax[1][:,None,:,None]**2+dmax[2][:,None,None,:]**2
        hit=(dmin2 < (r**2)[:,None,None,None]) & (dmax2 > ((r-f)**2)[:,None,None,None])
        bound+=hit.sum(0).astype(np.int32); touched+=int(hit.sum())
    # true block maxima
    pad=np.zeros((nb*b,)*3,np.int32); pad[:D,:D,:D]=vol
    bmax=pad.reshape(nb,b,nb,b,nb,b).max(axis=(1,3,5))
    assert (bound>=bmax).all()
    cand=(bound>=M).sum()
    # sequential best-first: process blocks by descending bound until bound < best exact so far
    order=np.argsort(-bound.ravel()); best=0; n_proc=0; work=0
    bm=bmax.ravel(); bd=bound.ravel()
    for i in order:
        if bd[i] < best: break
        n_proc+=1; work+=bd[i]; best=max(best,bm[i])
    return dict(N=len(P),D=D,b=b,votes=votes,peak=M,block_votes=touched,ratio_votes_per_blockvote=votes/touched,cand_blocks=int(cand),best_first_blocks=n_proc,
                exact_point_block_pairs=int(work), max_bound=int(bound.max()))
for fid in (5,6):
    fr=synth.config3_frame(fid)
    for b in (4,8):
        t=time.time(); print(fid, study(fr,0,b), round(time.time()-t,1))
