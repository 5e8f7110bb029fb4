"""Randomised check of the batched context against single-problem contexts (development aid; run under gpurun):
one segment of B stacked problems must give, problem by problem, the bits of the single-problem context."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
from fds_b200.device import Context


def main():
    rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    t0, bad, n = time.time(), 0, 0
    while time.time() - t0 < budget:
        B = int(rng.randint(2, 24))
        N = 64 * int(rng.choice([1, 2, 3, 4, 7, 16, 33, 64]))
        m = int(rng.choice([1, 2, 4, 5, 8, 16, 16, 32]))
        m = min(m, max(1, N // 3))
        hs = int(rng.choice([8, 16]))
        if 8 * hs > N:
            hs = 8
        steps = int(rng.randint(2, 3 * hs))
        scheme = 'rk4' if rng.rand() < 0.75 else 'euler'
        eps = 1e-4
        params = rng.uniform(-0.5, 0.7, B)
        U = rng.rand(B, N) * 3 + 1
        V = rng.randn(B, m, N) / np.sqrt(N)
        v = rng.randn(B, N) / np.sqrt(N)
        try:
            bc = Context(N, m, batch=B, max_halo_steps=hs, scheme=scheme)
            bc.set_batch_params(params)
            Ub, Vb, vb, Jb = bc.run_segment_host(U, V, v, 0.0, eps, steps)
            bc.close()
            one = Context(N, m, scheme=scheme, max_halo_steps=hs)
            ok = True
            for p in range(B):
                u1, V1, v1, J1 = one.run_segment_host(U[p], V[p], v[p], params[p], eps, steps)
                ok = ok and np.array_equal(Ub[p], u1) and np.array_equal(Vb[p], V1) and np.array_equal(vb[p], v1) \
                    and np.array_equal(Jb[:, p], J1)
            one.close()
            msg = ''
        except Exception as e:                        # noqa: BLE001
            ok, msg = False, repr(e)[:300]
        n += 1
        bad += 0 if ok else 1
        print('%s B=%d N=%d m=%d steps=%d halo=%d %s %s' % ('ok  ' if ok else 'FAIL', B, N, m, steps, hs, scheme, msg), flush=True)
    print('FUZZ_BATCH %d cases, %d failures' % (n, bad))
    sys.exit(1 if bad else 0)


main()
