import sys, ctypes, os
import numpy as np
sys.path.insert(0, '.')
os.environ['FDS_DEBUG'] = '4'
import fds_b200
from fds_b200 import _lib
N = 1 << 24; m = 32
eng = fds_b200.L96(N).engine(m)
eng.set_state(np.random.RandomState(0).rand(N))
eng.run(0.0, 50, want_J=False)
eng.orthonormal_random_tangents(seed=1)
eps = 1e-6 * np.sqrt(N)
eng.boundary(0.0, eps, 0, 0, 1)
eng.segment(0.0, eps, 5)
eng.ctx.sync()
lib = _lib.load()
buf = (ctypes.c_longlong * (64 * 16))()
lib.fds_debug_trace(buf)
t = np.array(buf[:]).reshape(64, 16)
t0 = t[0, 0]
names = ['P:wait0', 'P:wait1', 'P:issued', 'w0:wait0', 'w0:wait1', 'w0:end', 'w3:wait0', 'w3:wait1', 'w3:end', 'w13:wait0', 'w13:wait1', 'w13:end']
print('cycles relative to first stamp; one row per block')
print(' '.join('%9s' % n for n in names))
for r in t[:40]:
    print(' '.join('%9d' % (x - t0) for x in r[:12]))
d = np.diff(t[:, 5]); print('w0 block period mean', d.mean())
print('w0 wait mean', (t[:, 4] - t[:, 3]).mean(), 'compute mean', (t[:, 5] - t[:, 4]).mean())
print('w3 wait mean', (t[:, 7] - t[:, 6]).mean(), 'compute mean', (t[:, 8] - t[:, 7]).mean())
print('w13 wait mean', (t[:, 10] - t[:, 9]).mean(), 'compute mean', (t[:, 11] - t[:, 10]).mean())
print('producer: empty-wait mean', (t[:, 1] - t[:, 0]).mean(), 'issue mean', (t[:, 2] - t[:, 1]).mean())
