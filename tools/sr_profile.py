"""C3 step anatomy: a short device-resident symmetry-reduction run (for the ncu launch list)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402

nx = 512
pw = yh.default_params(nx, nx, timeIntOrder=1, lap4=0)
sim = yh.Sim(pw)
sim.cross_field_ic()
sim.run(12001, tb_steps=4)
tips = sim.tips()
u0, v0 = sim.get_state()
sim.close()
tx, ty = (float(tips[-1]["x"]), float(tips[-1]["y"])) if len(tips) else (nx / 2.0, nx / 2.0)
p = yh.default_params(nx, nx, reduce_sym=True, tipx0=tx, tipy0=ty)
sim = yh.Sim(p)
sim.set_state(u0, v0)
sim.run_sr(2, record=False)
sim.run_sr_device(int(sys.argv[1]) if len(sys.argv) > 1 else 30, record=False)
sim.close()
