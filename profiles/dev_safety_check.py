"""dev check: SafetyNet / emergency stage -- unmodified ROS node vs restatement vs port vs device code on the host"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import agrifly_b200 as agf, orc
from common import cfg_for, bit_equal
sc=agf.scenarios.stages_scenario(3, nticks=4000)
sc["ref"]=dict(sc["ref"], desired_pos=(2.5,0.0,1.0), safety_net=True)
ref=agf.offboard_ref(**sc["ref"]); est=agf.offboard_estimator(); oc=agf.offboard_cfg(5)
def run(O, node=False, chunks=None):
    v=O.vehicle(cfg_for(agf,sc),uwb_comm_period=0.0); v.set_state(pos=sc["pos"],att=sc["att"])
    if node:
        return v.run_stages_node(4000,oc,ref,est), v.stages_node_state()
    v.set_offboard_estimator(est)
    tr=np.vstack([v.run_offboard_ref(c,oc,ref) for c in (chunks or [4000])])
    return tr, v.offboard_state()
for math in ("glibc","shared"):
    R=orc.Oracle("ref-"+math)
    a,sa=run(R,node=True); b,sb=run(R); c,sc_=run(orc.Oracle("port-"+math),chunks=[1700,2300])
    print(math,"node==restatement",bit_equal(a,b),bit_equal(sa,sb),"port==node",bit_equal(a,c),bit_equal(sa,sc_),"stage",sa[0],"end pos",a[-1,0:3].round(3),"flight state",a[-1,34],"max x",a[:,0].max().round(3))
    if math=="shared":
        d,sd=run(orc.Oracle("hostsim-shared"),chunks=[1,999,3000]); print("   hostsim==node",bit_equal(a,d),bit_equal(sa,sd))
