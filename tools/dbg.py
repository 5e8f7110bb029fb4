import sys; sys.path.insert(0,'.')
import faulthandler; faulthandler.enable()
import numpy as np, fds_b200, traceback
from fds_b200 import _lib
N=1<<16; m=32
eng = fds_b200.L96(N).engine(m)
eng.set_state(np.random.rand(N)); eng.run(0.0, 10, want_J=False)
eng.orthonormal_random_tangents(seed=1)
print('ortho ok', flush=True)
eng.boundary(0.0, 1e-4, 0, 0, 1); print('b0 ok', flush=True)
eng.segment(0.0, 1e-4, 5); print('seg ok', flush=True)
c = eng.ctx
c.boundary_begin(1); print('begin', flush=True)
eng._halo(_lib.FDS_BUF_PRIMAL, 3); print('halo', flush=True)
c.boundary_dxdt(0.0, 1, 1e-4); c.sync(); print('dxdt', flush=True)
c.boundary_factor(1, 1, 1e-4); c.sync(); print('factor', flush=True)
c.boundary_finish(1, 1e-4, 1); c.sync(); print('finish', flush=True)
print(c.boundary_results(1)[1][:3], flush=True)
