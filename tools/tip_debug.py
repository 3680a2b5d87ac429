import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh
from tests import oracle_lib
oracle = oracle_lib.load()
nx = ny = 256
pe = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)
sim = yh.Sim(pe); sim.cross_field_ic(); sim.run(9000, tb_steps=4)
u0, v0 = (a[0] for a in sim.get_state()); sim.close()
p = oracle.params_default(nx, ny)
ref = oracle_lib.Reference(nofma=False); ref.init(p)
ours = yh.Sim(p); ours.set_state(u0, v0)
ours.run(99, tb_steps=1)
pu_o = ours.get_state()[0][0].copy()
ru, rv, _ = ref.rd_run(u0, v0, 99); pu_r = ru
ours.run(1, tb_steps=1)
ru, rv, _ = ref.rd_run(ru, rv, 1)
ou = ours.get_state()[0][0]
t_o = ours.tips()
print("ours tips", len(t_o), "ref(ref fields)", len(ref.tip(ru, pu_r)), "ref(our fields)", len(ref.tip(ou, pu_o)),
      "oracle(our fields)", len(oracle.tip_track(p, pu_o, ou)), "oracle(ref fields)", len(oracle.tip_track(p, pu_r, ru)))
print("max |ours - ref| u:", np.abs(ou - ru).max())
for t in t_o[:6]:
    i, j = int(t["x"]), int(t["y"])
    print(t, "present", ou[j:j+2, i:i+2].ravel(), "past", pu_o[j:j+2, i:i+2].ravel(), "ref present", ru[j:j+2, i:i+2].ravel())
