"""dev check: the UNMODIFIED ExampleVehicleStateMachine (ROS rates-control node) vs the restated stage logic, same loop"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import agrifly_b200 as agf, orc
from common import cfg_for, bit_equal
for math in ("glibc","shared"):
    R=orc.Oracle("ref-"+math)
    sc=agf.scenarios.stages_scenario(3, nticks=7000)
    sc["pos"]=(0.2,-0.1,0.0)
    ref=agf.offboard_ref(**sc["ref"]); est=agf.offboard_estimator(); oc=agf.offboard_cfg(5)
    a=R.vehicle(cfg_for(agf,sc),uwb_comm_period=0.0); a.set_state(pos=sc["pos"],att=sc["att"])
    ta=a.run_stages_node(7000,oc,ref,est); sa=a.stages_node_state()
    b=R.vehicle(cfg_for(agf,sc),uwb_comm_period=0.0); b.set_state(pos=sc["pos"],att=sc["att"]); b.set_offboard_estimator(est)
    tb=b.run_offboard_ref(7000,oc,ref); sb=b.offboard_state()
    d=~((ta==tb)|(np.isnan(ta)&np.isnan(tb)))
    print(math,"node == restatement:",bit_equal(ta,tb),"first diff tick",(np.argmax(d.any(1)) if d.any() else None),"end",ta[-1,0:3].round(4),tb[-1,0:3].round(4))
    print("  node state",sa.round(4)); print("  rest state",sb.round(4))
