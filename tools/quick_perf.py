"""Quick wall-clock look at the step kernel and the boundary (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import fds_b200

def main():
    logN = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    math = sys.argv[3] if len(sys.argv) > 3 else 'exact'
    scheme = sys.argv[4] if len(sys.argv) > 4 else 'rk4'
    N = 1 << logN
    eps = 1e-6 * np.sqrt(N)
    run = fds_b200.L96(N, math=math, scheme=scheme)
    eng = run.engine(m)
    eng.set_state(np.random.RandomState(0).rand(N))
    eng.run(0.0, 100, want_J=False)
    eng.orthonormal_random_tangents(seed=1)
    eng.boundary(0.0, eps, 0, 0, 1)
    eng.segment(0.0, eps, 100)
    eng.ctx.sync()
    for steps in (20, 100):
        l0 = eng.ctx.launches
        t0 = time.perf_counter()
        eng.segment(0.0, eps, steps)
        eng.ctx.sync()
        dt = time.perf_counter() - t0
        upd = N * (m + 1) * steps / dt
        gbs = 16.0 * N * (m + 2) * steps / dt / 1e9
        print(scheme + ' N=2^%d m=%d %s: %d steps %.2f ms/step  %.3e upd/s  %.0f GB/s algorithmic  (%d launches)'
              % (logN, m, math, steps, dt / steps * 1e3, upd, gbs, eng.ctx.launches - l0))
    for keep in (True, False):
        eng.segment(0.0, eps, 2)
        eng.ctx.sync()
        t0 = time.perf_counter()
        eng.boundary(0.0, eps, 1, 1, 1, keep_tangents=keep)
        eng.ctx.sync()
        dt = time.perf_counter() - t0
        print('boundary (dilation+QR+ensemble, keep_tangents=%s): %.2f ms' % (keep, dt * 1e3))

main()
