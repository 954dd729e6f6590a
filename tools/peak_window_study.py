"""Census for an exact peak-only mode: conservative per-box upper bounds U (sphere shell meets the b^3 box of voxel
centres) against the exact peak L of the same item; which boxes would still have to be rendered exactly (U >= L).  CPU, oracle only."""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import oracle
from rcvpose_b200 import synth
K=synth.linemod_K
for f in (3,7):
    fr=synth.config3_frame(f)
    for k in range(3):
        depth=fr["depth"]; rad=fr["radius"][k]
        xyz, r = synth.frame_to_points(K, depth, rad)
        pre=oracle.prelude(xyz/1000.0 if xyz.max()>50 else xyz, r)
        p,R,D=pre["p"],pre["R"],pre["D"]
        vol=oracle.fast_for(p,R,D,method="scatter")
        L=vol.max(); n=len(R)
        for b in (4,8):
            nb=(D+b-1)//b
            # conservative: sphere shell intersects box of voxel centres [bI, bI+b-1]
            lo=np.arange(nb)*b; hi=np.minimum(lo+b-1,D-1)
            U=np.zeros((nb,nb,nb),np.int32)
            for q in range(n):
                c=p[q]
                dmin=[np.maximum(np.maximum(lo-c[a], c[a]-hi),0) for a in range(3)]
                dmax=[np.maximum(np.abs(lo-c[a]),np.abs(hi-c[a])) for a in range(3)]
                mn=np.sqrt(dmin[0][:,None,None]**2+dmin[1][None,:,None]**2+dmin[2][None,None,:]**2)
                mx=np.sqrt(dmax[0][:,None,None]**2+dmax[1][None,:,None]**2+dmax[2][None,None,:]**2)
                U+=((mn<R[q])&(mx>R[q]-0.4331)).astype(np.int32)
            must=np.argwhere(U>=L)
            bb=[(must[:,a].min(),must[:,a].max()) for a in range(3)]
            wx=(bb[0][1]-bb[0][0]+1)*b; wy=(bb[1][1]-bb[1][0]+1)*b
            print(f"f{f} k{k} D={D} n={n} L={L} Umax={U.max()} b={b}: must boxes {len(must)} of {nb**3}, bbox {bb}, window {wx}x{wy} of {D}x{D} -> column fraction {wx*wy/(D*D):.3f}; coarse cols {(nb*nb)/(D*D):.3f}")
