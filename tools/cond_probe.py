"""Condition bound and CholeskyQR passes of every boundary of a long device-resident run (development aid)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import fds_b200

logN = int(sys.argv[1]) if len(sys.argv) > 1 else 24
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
N, m, S = 1 << logN, 32, 100
eps = 1e-6 * np.sqrt(N)
run = fds_b200.L96(N)
eng = run.engine(m)
eng.ctx.set_state(np.random.RandomState(0).rand(N))
eng.run(0.0, 200, want_J=False)
eng.orthonormal_random_tangents(seed=1)
eng.boundary(0.0, eps, 0, 0, 1)
eng.segment(0.0, eps, S)
eng.begin_deferred(K, S)
for _ in range(K):
    eng.boundary_deferred(0.0, eps, 1, 1, 1, keep_tangents=False)
    eng.segment_deferred(0.0, eps, S)
q = eng._dq
eng.end_deferred()
for i in range(K):
    R, b, Gd, gd, two, cond = eng.ctx.boundary_unpack(q['out'][i], 1)
    dg = np.abs(np.diag(R))
    print('boundary %2d: cond bound %8.2f  passes %d  diag R min %.3f max %.3f  |Gdil| %.2e' % (i, cond, 2 if two else 1, dg.min(), dg.max(), np.abs(Gd).max()))
run.close()
