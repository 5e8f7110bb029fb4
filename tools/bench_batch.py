"""Throughput of the batched parameter sweep (BASELINE configs[4]): B independent Lorenz-96
problems of N sites, m tangents, in one device context.  python tools/bench_batch.py [B N m K]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import fds_b200


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    steps, eps = 100, 1e-4
    run = fds_b200.L96(N)
    eng = run.batch_engine(m, B)
    eng.ctx.set_batch_params(np.linspace(-2.0, 2.0, B))          # F swept over [6, 10]
    eng.set_state(np.random.RandomState(0).rand(B, N))
    eng.run(0.0, 200, want_J=False)
    eng.orthonormal_random_tangents(seed=1)
    eng.boundary(0.0, eps, 0, 0, 1)
    eng.segment(0.0, eps, steps)

    def seg():
        R, b, Gd, gd = eng.boundary(0.0, eps, 1, 1, 1, keep_tangents=False)
        J, G, g = eng.segment_sens(0.0, eps, steps)
        return R, G
    seg()
    eng.ctx.sync()
    l0 = eng.ctx.launches
    t0 = time.perf_counter()
    for _ in range(K):
        R, J = seg()
    eng.ctx.sync()
    dt = time.perf_counter() - t0
    tb0 = time.perf_counter()
    eng.boundary(0.0, eps, 1, 1, 1, keep_tangents=False)
    eng.ctx.sync()
    tb = time.perf_counter() - tb0
    assert np.isfinite(J).all() and np.isfinite(R).all() and (np.diagonal(R, axis1=1, axis2=2) > 0).all()
    upd = B * N * (m + 1) * steps * K / dt
    print('batch B=%d N=%d m=%d: %d segments x %d steps in %.3f s -> %.3e state-updates/s '
          '(%.1f GB/s algorithmic, boundary %.1f ms, %d launches)'
          % (B, N, m, K, steps, dt, upd, 16.0 * B * N * (m + 2) * steps * K / dt / 1e9, tb * 1e3,
             eng.ctx.launches - l0))


main()
