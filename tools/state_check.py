import sys; sys.path.insert(0,'.')
import numpy as np, veritas_b200 as vb
from veritas_b200 import solver as S
run = vb.LaserPlasmaRun(4096, 1024, density=0.1)
run.init_device(); n=run.run_fields_phase(); print('fields steps',n)
ctx=run.ctx
def q():
    ctx.moments(); return [repr(float(np.sum(ctx.get_1d(S.CHARGES0+s)))) for s in range(2)]
f0=ctx.download_f(0,0,1).copy()
print(q())
for k in range(5):
    dt=run.calculate_dt(); run.advance(dt)
print(q(), 'dt',dt)
f1=ctx.download_f(0,0,1)
print('f change rel', np.linalg.norm(f1-f0)/np.linalg.norm(f0), 'max f', f0.max(), 'J max', np.abs(ctx.get_1d(S.J)).max(), 'E max', np.abs(ctx.get_1d(S.EFIELD)).max())
