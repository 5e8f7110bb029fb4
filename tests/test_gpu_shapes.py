"""Edge shapes of the hot path against the oracle port (which is pinned against the reference):
ragged N (not a multiple of 8 / 64), the smallest and a large subspace dimension, the minimum
number of steps, odd step counts (two-steps-per-pass kernel + one trailing single step), every
kernel family (tensor-core boundary for m = 16 / 32, register-blocked and generic ones otherwise)."""
import numpy as np
import pytest

import fds_b200
from oracle import plugins, shadowing_port as port

gpu = pytest.mark.gpu
TOL = 1e-11


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@gpu
@pytest.mark.parametrize('N,m,steps', [(4099, 7, 5), (37, 1, 2), (1024, 64, 3), (520, 16, 9), (8200, 32, 4),
                                       (64, 3, 2), (100, 48, 7)])
def test_segment_and_boundary_against_the_port(N, m, steps):
    rng = np.random.RandomState(N + m)
    s, eps = 0.1, 1e-4
    u0 = plugins.run_l96_rk4(rng.rand(N), s, 40)[0]
    V = np.linalg.qr(rng.rand(m, N).T)[0].T
    v = rng.randn(N) / np.sqrt(N)
    run = fds_b200.L96(N)
    # one segment through the host-buffer entry point: bit for bit
    u1, V1, v1, J0, G, g = fds_b200.segment.run_segment(run, u0, V, v, s, 0, steps, eps)
    ru1, rV1, rv1, rJ0, rG, rg = port.run_segment(plugins.run_l96_rk4, u0, V, v, s, steps, eps)
    assert np.array_equal(u1, ru1) and np.array_equal(V1, rV1) and np.array_equal(v1, rv1)
    assert np.array_equal(J0, rJ0)
    if steps >= 2:
        assert np.array_equal(np.array(G), np.array(rG)) and np.array_equal(g, rg)
    # the boundary that follows, device resident, vs time dilation + projection + QR of the port
    eng = run.engine(m)
    eng.set_dxdt_weights(fds_b200.timedilation.stencil_weights(3))
    eng.set_state(u1)
    eng.set_tangents(V1, v1)
    R, b, Gd, gd = eng.boundary(s, eps, from_ensemble=0, do_qr=1, build=0)
    Q, vq = eng.get_tangents()
    td = port.TimeDilation(plugins.run_l96_rk4, ru1, s)
    lss = port.LssTangent()
    rQ, rvq = lss.checkpoint(td.project(rV1), td.project(rv1))
    sgn = np.sign(np.diag(lss.Rs[0]))
    assert maxrel(Gd, td.contribution(rV1)) < TOL and abs(gd - td.contribution(rv1)) <= TOL * abs(gd) + 1e-15
    assert maxrel(R, sgn[:, None] * lss.Rs[0]) < TOL and maxrel(b, sgn * lss.bs[0]) < 1e-10
    assert maxrel(Q, sgn[:, None] * rQ) < TOL * 10 and maxrel(vq, rvq) < 1e-10
    assert np.abs(Q @ Q.T - np.eye(m)).max() < 1e-13
    eng.close()
