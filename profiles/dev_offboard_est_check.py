"""dev check: offboard loop with the MocapStateEstimator -- reference vs port vs device code on the host"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import agrifly_b200 as agf, orc
from common import cfg_for, bit_equal
s=agf.scenarios
def first_diff(a,c):
    bad=np.argwhere(~((a==c)|(np.isnan(a)&np.isnan(c))))
    return (bad[0], a[bad[0][0],bad[0][1]], c[bad[0][0],bad[0][1]]) if len(bad) else None
def run(O, sc, chunks=None, est=None, offset=None):
    v=O.vehicle(cfg_for(agf,sc),uwb_comm_period=0.0); v.set_state(pos=sc["pos"],att=sc["att"])
    v.set_offboard_estimator(est)
    oc=agf.offboard_cfg(5)
    n=sc["nticks"]
    if "ref" in sc:
        ref=agf.offboard_ref(**sc["ref"]); rec=None if sc["primitive"] is None else agf.primitive_record(**sc["primitive"])
        parts=[v.run_offboard_ref(c,oc,ref,offset=offset,trajectory=rec) for c in (chunks or [n])]
    else:
        parts=[v.run_offboard(c,oc,sc["targets"],offset=offset) for c in (chunks or [n])]
    e,c=v.offboard_estimate(0.0); e2,_=v.offboard_estimate(0.03)
    return np.vstack(parts), np.concatenate([e,e2,c])
est=agf.offboard_estimator()
for math in ["shared","glibc"]:
    R=orc.Oracle("ref-"+math); P=orc.Oracle("port-"+math)
    Hs=[orc.Oracle("hostsim-shared")] if math=="shared" else []
    for sc in [s.offboard_scenario(), s.stages_scenario(1), s.stages_scenario(4), s.tracking_scenario()]:
        a,ea=run(R,sc,est=est); b,eb=run(P,sc,est=est,chunks=[1500,sc["nticks"]-1500])
        print(math,sc["name"],"port==ref",bit_equal(a,b),bit_equal(ea,eb),"end",a[-1,0:3].round(3),"est",ea[0:3].round(3),"cnt",ea[26:], first_diff(a,b), end=" ")
        for H in Hs:
            c,ec=run(H,sc,est=est,chunks=[1,1234,sc["nticks"]-1235])
            print("hostsim==ref",bit_equal(a,c),bit_equal(ea,ec),first_diff(a,c),end="")
        print()
# measurement rejection and the forced reset: teleport the vehicle in mid-flight
def run_jump(O, sc):
    v=O.vehicle(cfg_for(agf,sc),uwb_comm_period=0.0); v.set_state(pos=sc["pos"],att=sc["att"])
    v.set_offboard_estimator(est)
    oc=agf.offboard_cfg(5)
    a=v.run_offboard(1000,oc,sc["targets"])
    p=a[-1,0:3]+np.array([3.0,0.5,0.0])
    v.set_state(pos=p, vel=a[-1,3:6], att=a[-1,6:10], ang_vel=a[-1,10:13])
    parts=[a]
    cs=[]
    for k in range(20):
        parts.append(v.run_offboard(5,oc,sc["targets"])); cs.append(v.offboard_estimate(0.0)[1])
    parts.append(v.run_offboard(900,oc,sc["targets"]))
    e,c=v.offboard_estimate(0.03)
    return np.vstack(parts), np.concatenate([e,c]+cs)
sc=s.offboard_scenario()
for math in ["shared","glibc"]:
    a,ea=run_jump(orc.Oracle("ref-"+math),sc); b,eb=run_jump(orc.Oracle("port-"+math),sc)
    print(math,"jump: port==ref",bit_equal(a,b),bit_equal(ea,eb),"rejected",ea[14],"end",a[-1,0:3].round(3),"panic",a[-1,35],first_diff(a,b),end=" ")
    if math=="shared":
        c,ec=run_jump(orc.Oracle("hostsim-shared"),sc); print("hostsim==ref",bit_equal(a,c),bit_equal(ea,ec),first_diff(a,c),end="")
    print()
