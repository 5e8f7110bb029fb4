"""Randomised parity sweep against the oracle port (development aid; run under gpurun):
one segment bit for bit + the boundary that follows, over random sizes, subspace dimensions, step counts, halo
capacities, schemes and orders of the time-dilation stencil."""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import fds_b200
from oracle import plugins, shadowing_port as port


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def one(case, N, m, steps, scheme, halo, order):
    rng = np.random.RandomState(case)
    s, eps = 0.1, 1e-4
    orun = plugins.run_l96_rk4 if scheme == 'rk4' else plugins.run_l96_euler
    u0 = orun(rng.rand(N), s, 30)[0]
    V = np.linalg.qr(rng.rand(m, N).T)[0].T
    v = rng.randn(N) / np.sqrt(N)
    run = fds_b200.L96(N, scheme=scheme, max_halo_steps=halo)
    u1, V1, v1, J0, G, g = fds_b200.segment.run_segment(run, u0, V, v, s, 0, steps, eps)
    ru1, rV1, rv1, rJ0, rG, rg = port.run_segment(orun, u0, V, v, s, steps, eps)
    ok = np.array_equal(u1, ru1) and np.array_equal(V1, rV1) and np.array_equal(v1, rv1) and np.array_equal(J0, rJ0)
    ok = ok and np.array_equal(np.array(G), np.array(rG)) and np.array_equal(g, rg)
    # device-resident: segment then boundary (exercises the primal column written by the last pass)
    eng = run.engine(m)
    port.TimeDilation.ORDER = order
    try:
        eng.set_dxdt_weights(fds_b200.timedilation.stencil_weights(order))
        eng.set_state(u0)
        eng.set_tangents(V, v)
        eng.boundary(s, eps, from_ensemble=0, do_qr=0, build=1)        # projection only, builds the ensemble
        td0 = port.TimeDilation(orun, u0, s)
        pV, pv = td0.project(V), td0.project(v)
        eng.segment(s, eps, steps)
        R, b, Gd, gd = eng.boundary(s, eps, from_ensemble=1, do_qr=1, build=0)
        a1, aV, av, _, _, _ = port.run_segment(orun, u0, pV, pv, s, steps, eps)
        td = port.TimeDilation(orun, a1, s)
        eG = maxrel(Gd, td.contribution(aV))
        lss = port.LssTangent()
        lss.checkpoint(td.project(aV), td.project(av))
        sgn = np.sign(np.diag(lss.Rs[0]))
        eR = maxrel(R, sgn[:, None] * lss.Rs[0])
    finally:
        port.TimeDilation.ORDER = 3
        run.close()
    good = ok and eG < 1e-9 and eR < 1e-9
    return good, 'bitwise %s  G_dil %.1e  R %.1e' % (ok, eG, eR)


def main():
    rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    special_N = [2012, 2013, 2011, 4024, 2048, 1952, 64, 65, 8, 9, 17, 1000, 4097]
    t0, bad, n = time.time(), 0, 0
    while time.time() - t0 < budget:
        N = int(special_N[n]) if n < len(special_N) else int(rng.choice([rng.randint(8, 300), rng.randint(300, 6000), rng.randint(6000, 30000)]))
        m = int(rng.choice([1, 2, 3, 5, 8, 16, 16, 32, 32, 33, 40]))
        m = min(m, max(1, N // 3))
        steps = int(rng.randint(2, 24))
        scheme = 'rk4' if rng.rand() < 0.7 else 'euler'
        halo = int(rng.choice([0, 3, 4, 8, 16]))
        order = int(rng.choice([1, 2, 3, 3, 3, 4, 5]))
        if halo and (8 * halo > N or halo < order):
            halo = 0
        if N < 64 * max(3, order):          # default halo capacity is n / 64 (>= 3): keep the stencil inside it
            order = 3
        try:
            good, msg = one(n, N, m, steps, scheme, halo, order)
        except Exception as e:                                      # noqa: BLE001
            good, msg = False, 'EXCEPTION %r' % (e,)
        n += 1
        if not good:
            bad += 1
        print('%s N=%d m=%d steps=%d %s halo=%d order=%d: %s' % ('ok  ' if good else 'FAIL', N, m, steps, scheme, halo, order, msg), flush=True)
    print('FUZZ %d cases, %d failures' % (n, bad))
    sys.exit(1 if bad else 0)


main()
