import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import agrifly_b200 as agf, orc
from common import cfg_for, bit_equal
sc=agf.scenarios.offboard_scenario()
def run(O, kind, tid=0, n=8000, chunks=None, offset=None, rec=None, ref=None):
    v=O.vehicle(cfg_for(agf,sc),uwb_comm_period=0.0); v.set_state(pos=(0.2,-0.1,0),att=(1,0,0,0))
    oc=agf.offboard_cfg(5)
    parts=[v.run_offboard_ref(c,oc,ref,offset=offset,trajectory=rec) for c in (chunks or [n])]
    return np.vstack(parts), v.offboard_state()
for math in ["shared","glibc"]:
    R=orc.Oracle("ref-"+math); P=orc.Oracle("port-"+math)
    Hs=[orc.Oracle("hostsim-shared")] if math=="shared" else []
    for tid in range(6):
        ref=agf.offboard_ref(1,500000,9000000,(0,0,1.0),0.3,tid)
        a,sa=run(R,1,tid,ref=ref,offset=(0.1,0.2,0.05)); b,sb=run(P,1,tid,ref=ref,offset=(0.1,0.2,0.05),chunks=[3000,5000])
        print(math,"stages",tid,"port==ref",bit_equal(a,b),bit_equal(sa,sb), end=" ")
        for H in Hs:
            c,sc_=run(H,1,tid,ref=ref,offset=(0.1,0.2,0.05),chunks=[1,1234,6765])
            print("hostsim==ref",bit_equal(a,c),bit_equal(sa,sc_), end="")
            if not bit_equal(a,c):
                bad=np.argwhere(~((a==c)|(np.isnan(a)&np.isnan(c)))); print(" first diff",bad[0], a[bad[0][0],bad[0][1]], c[bad[0][0],bad[0][1]],end="")
        print()
    rec=agf.primitive_record((1.5,0.5,0.3),2.5,offset=(0.2,-0.1,2.0),att=(np.cos(0.2),0,0,np.sin(0.2)))
    ref=agf.offboard_ref(2,4000000,0,(0.2,-0.1,2.0),0.1)
    a,_=run(R,2,n=3400,ref=ref,rec=rec); b,_=run(P,2,n=3400,ref=ref,rec=rec,chunks=[2100,1300])
    print(math,"trajectory port==ref",bit_equal(a,b),"end",a[-1,0:3].round(3),"panic",a[-1,35], end=" ")
    for H in Hs:
        c,_=run(H,2,n=3400,ref=ref,rec=rec,chunks=[1999,1401]); print("hostsim==ref",bit_equal(a,c),end="")
        if not bit_equal(a,c):
            bad=np.argwhere(~((a==c)|(np.isnan(a)&np.isnan(c)))); print(" first diff",bad[0], a[bad[0][0],bad[0][1]], c[bad[0][0],bad[0][1]],end="")
    print()
