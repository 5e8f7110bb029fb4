"""A/B of library builds on ONE board (development aid): step-kernel time per launch, boundary time.

    python tools/ab_step.py libA.so libB.so [...]      (run under gpurun; variants alternate twice)

Boards differ by +-3 % through their power caps, so variants are only comparable inside one call.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np
import fds_b200._lib as L
L.LIB_PATH = sys.argv[1]
import fds_b200
N, m, S = 1 << 24, 32, 100
eps = 1e-6 * np.sqrt(N)
out = []
for math in sys.argv[2].split(','):
    run = fds_b200.L96(N, math=math)
    eng = run.engine(m)
    eng.ctx.set_state(np.random.RandomState(0).rand(N))
    eng.run(0.0, 100, want_J=False)
    eng.orthonormal_random_tangents(seed=1)
    eng.boundary(0.0, eps, 0, 0, 1)
    eng.segment(0.0, eps, S)
    eng.ctx.profile_enable(True)
    t0 = time.perf_counter()
    K = 3
    eng.begin_deferred(K, S)
    for _ in range(K):
        eng.boundary_deferred(0.0, eps, 1, 1, 1, keep_tangents=False)
        eng.segment_deferred(0.0, eps, S)
    eng.end_deferred()
    dt = (time.perf_counter() - t0) / K
    ms, n = eng.ctx.profile_read()
    out.append('%%s: %%.4f ms/launch (%%d), segment %%.2f ms, non-step %%.2f ms' %% (math, ms / n, n, dt * 1e3, dt * 1e3 - ms / K))
    run.close()
print(' | '.join(out))
''' % ROOT

libs = [a for a in sys.argv[1:] if a.endswith('.so')]
maths = 'exact,fast'
for rep in range(2):
    for lib in libs:
        r = subprocess.run([sys.executable, '-c', CHILD, os.path.abspath(lib), maths], capture_output=True, text=True)
        print('%-28s %s' % (os.path.basename(lib), r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
