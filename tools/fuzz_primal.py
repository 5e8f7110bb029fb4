"""Randomised parity sweep of the primal marches (run-up and u1, J = run(u0, s, steps)) against the oracle:
random sizes, step counts, halo capacities, schemes, with and without the objective; single problems and batches."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import fds_b200
from fds_b200.device import Context
from oracle import plugins


def main():
    rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    special_N = [1920, 1921, 1919, 2048, 1952, 3840, 64, 65, 63, 8, 9, 17, 1000, 4097, 1856]
    t0, bad, n = time.time(), 0, 0
    while time.time() - t0 < budget:
        N = int(special_N[n]) if n < len(special_N) else int(rng.choice([rng.randint(8, 300), rng.randint(300, 6000), rng.randint(6000, 200000)]))
        m = int(rng.choice([1, 2, 5, 16, 32]))
        m = min(m, max(1, N // 3))
        steps = int(rng.randint(1, 70))
        scheme = 'rk4' if rng.rand() < 0.6 else 'euler'
        orun = plugins.run_l96_rk4 if scheme == 'rk4' else plugins.run_l96_euler
        halo = int(rng.choice([0, 1, 3, 8, 11, 16]))
        if halo and 8 * halo > N:
            halo = 0
        s = float(rng.rand() - 0.5)
        batch = int(rng.choice([1, 1, 1, 3])) if N % 64 == 0 else 1
        try:
            if batch > 1:
                ds = rng.rand(batch) - 0.5
                U = rng.rand(batch, N) * 2 - 0.5
                bc = Context(N, m, batch=batch, scheme=scheme, max_halo_steps=halo or 8)
                bc.set_batch_params(ds)
                bc.set_state(U)
                bc.run(s, steps, want_J=False)
                Ub = bc.get_state()
                bc.close()
                good = all(np.array_equal(Ub[p], orun(U[p], s + ds[p], steps)[0]) for p in range(batch))
                msg = 'batch of %d' % batch
            else:
                u0 = rng.rand(N) * 2 - 0.5
                eng = fds_b200.L96(N, scheme=scheme, max_halo_steps=halo).engine(m)
                eng.set_state(u0)
                eng.run(s, steps, want_J=False)
                a = eng.get_state().copy()
                eng.set_state(u0)
                J = eng.run(s, steps, want_J=True)
                b = eng.get_state().copy()
                eng.close()
                u1, Jo = orun(u0, s, steps)
                good = np.array_equal(a, u1) and np.array_equal(b, u1) and np.array_equal(J, Jo)
                msg = 'state %s/%s J %s' % (np.array_equal(a, u1), np.array_equal(b, u1), np.array_equal(J, Jo))
        except Exception as e:                                      # noqa: BLE001
            good, msg = False, 'EXCEPTION %r' % (e,)
        n += 1
        if not good:
            bad += 1
            print('FAIL N=%d m=%d steps=%d %s halo=%d: %s' % (N, m, steps, scheme, halo, msg), flush=True)
    print('FUZZ-PRIMAL %d cases, %d failures' % (n, bad))
    sys.exit(1 if bad else 0)


main()
