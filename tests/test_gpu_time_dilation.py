"""Time-dilation variants ("the analytic run_ddt path done right").  The reference's
TimeDilationExact (fds/timedilation.py:54-61) raises NameError upstream, so there is no reference output to
pin against; what can be pinned: the analytic derivative on the device equals the same formula on the host
bit for bit, a callable run_ddt reproduces it through the public API, both agree with the finite-difference
default to the order of the stencil, and dJ/ds stays within the reference's error bar."""
import numpy as np
import pytest

import fds_b200
from oracle import plugins
from conftest import golden

gpu = pytest.mark.gpu
DT = 0.01


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def ddt_host(u, s):
    """du per step = dt * f(u), same operation order as the device kernel."""
    return DT * plugins.l96_rhs_factored(np.asarray(u), s + plugins.L96_F_OFFSET)


def first_boundary(N, m, mode, u, V, v, s=0.1, eps=1e-4):
    eng = fds_b200.L96(N).engine(m)
    eng.set_dxdt_weights(fds_b200.timedilation.stencil_weights(3))
    eng.set_time_dilation(mode)
    eng.set_state(u)
    eng.set_tangents(V, v)
    R, b, Gd, gd = eng.boundary(s, eps, from_ensemble=0, do_qr=1, build=0)
    Q, vq = eng.get_tangents()
    eng.close()
    return R, b, Gd, gd, Q, vq


@gpu
@pytest.mark.parametrize('N,m', [(40, 5), (1000, 16)])
def test_analytic_equals_callable_and_is_close_to_finite_differences(N, m):
    rng = np.random.RandomState(N)
    u = plugins.run_l96_rk4(rng.rand(N), 0.1, 300)[0]
    V = np.linalg.qr(rng.rand(m, N).T)[0].T
    v = rng.randn(N) / np.sqrt(N)
    an = first_boundary(N, m, 'analytic', u, V, v)
    cb = first_boundary(N, m, ddt_host, u, V, v)
    fd = first_boundary(N, m, 1, u, V, v)
    for a, c in zip(an, cb):                                   # device formula == host formula: same bits
        assert np.array_equal(np.asarray(a), np.asarray(c))
    # 3rd-order one-sided differences vs the exact derivative: O((dt |f'|)^3) ~ 1e-3 relative at |x| ~ 10
    assert maxrel(an[2], fd[2]) < 3e-3 and abs(an[3] - fd[3]) < 3e-3 * abs(fd[3]) + 1e-6
    assert maxrel(an[0], fd[0]) < 3e-3 and maxrel(an[4], fd[4]) < 3e-3
    # projection: the tangents are orthogonal to the analytic direction
    d = ddt_host(u, 0.1)
    assert np.abs(an[4] @ d).max() < 1e-12 * np.linalg.norm(d) and abs(an[5] @ d) < 1e-10 * np.linalg.norm(d)


@gpu
def test_cfg3_with_analytic_and_callable_time_dilation():
    g = golden('cfg3_l96rk4_N40_m16_K100')
    out = {}
    for key, mode in (('analytic', 'analytic'), ('callable', ddt_host)):
        np.random.seed(int(g['seed']))
        cp = fds_b200.shadowing(fds_b200.L96(40), g['u0'], float(g['s']), int(g['m']), int(g['K']), int(g['steps']),
                                int(g['runup']), epsilon=float(g['eps']), run_ddt=mode, return_checkpoint=True)
        out[key] = cp
    a, c = out['analytic'], out['callable']
    assert np.array_equal(np.array(a.J), np.array(c.J)) and np.array_equal(np.array(a.lss.Rs), np.array(c.lss.Rs))
    assert np.array_equal(np.array(a.G_dil), np.array(c.G_dil))
    assert np.array_equal(np.array(a.J)[0], g['J'][0])          # the primal does not depend on the dilation
    s1, s2 = fds_b200.lss_gradient_series(a)
    G, Gerr = fds_b200.timeseries.mean_std(s1 + s2)
    assert (np.abs(G - g['G_ref']) <= g['G_err'] + Gerr).all(), (G, g['G_ref'], g['G_err'], Gerr)


@gpu
def test_batch_and_host_solver_accept_the_variants():
    N, m = 64, 4
    u0 = plugins.run_l96_rk4(np.random.RandomState(0).rand(N), 0.0, 200)[0]
    np.random.seed(0)
    Ja, Ga = fds_b200.shadowing_batch(fds_b200.L96(N), u0, np.array([0.0, 0.2]), m, 3, 20, 0, 1e-4, run_ddt='analytic')
    np.random.seed(0)
    Jf, Gf = fds_b200.shadowing_batch(fds_b200.L96(N), u0, np.array([0.0, 0.2]), m, 3, 20, 0, 1e-4)
    assert np.array_equal(Ja, Jf) and maxrel(Ga, Gf) < 1e-2
    np.random.seed(0)
    Jh, Gh = fds_b200.shadowing(fds_b200.HostSolver(plugins.run_l96_rk4), u0, 0.0, m, 3, 20, 0, epsilon=1e-4,
                                run_ddt=ddt_host)
    np.random.seed(0)
    Jd, Gd = fds_b200.shadowing(fds_b200.L96(N), u0, 0.0, m, 3, 20, 0, epsilon=1e-4, run_ddt='analytic')
    assert np.array_equal(Jh, Jd) and maxrel(Gh, Gd) < 1e-6
    with pytest.raises(fds_b200.FdsError, match='analytic'):
        fds_b200.shadowing(fds_b200.Lorenz63(), np.array([1.0, 1.0, 28.0]), 28.0, 2, 2, 100, 0, run_ddt='analytic')


@gpu
@pytest.mark.parametrize('N,m,scheme', [(40, 5, 'rk4'), (5000, 16, 'rk4'), (1 << 16, 32, 'rk4'), (3000, 8, 'euler')])
def test_fused_time_dilation_equals_one_kernel_per_step(monkeypatch, N, m, scheme):
    """The one-pass kernel (u_1, u_2, u_3 and d in shared memory tiles) against the cross-check path (three
    single-column solver steps + the stencil kernel, FDS_DILATION_SPLIT): same arithmetic, same bits in everything
    that depends on d."""
    rng = np.random.RandomState(N)
    step = plugins.run_l96_rk4 if scheme == 'rk4' else plugins.run_l96_euler
    u = step(rng.rand(N), 0.1, 100)[0]
    V, v = rng.rand(m, N), rng.randn(N)
    res = []
    for split in (False, True):
        if split:
            monkeypatch.setenv('FDS_DILATION_SPLIT', '1')
        else:
            monkeypatch.delenv('FDS_DILATION_SPLIT', raising=False)
        eng = fds_b200.L96(N, scheme=scheme).engine(m)
        eng.set_state(u)
        eng.set_tangents(V, v)
        R, b, Gd, gd = eng.boundary(0.1, 1e-4, from_ensemble=0, do_qr=1, build=0)
        res.append((R, b, Gd, gd) + tuple(eng.get_tangents()))
        eng.close()
    for a, c in zip(*res):
        assert np.array_equal(np.asarray(a), np.asarray(c))


@gpu
@pytest.mark.parametrize('order', [2, 4, 6])
def test_other_orders_of_accuracy_match_the_oracle(order):
    """set_order_of_accuracy (reference fds/timedilation.py:8-12): `order` look-ahead steps and order + 1 weights;
    contribution and projection against the oracle's TimeDilation with the same order."""
    from oracle import shadowing_port as port
    N, m, s = 777, 6, 0.1
    rng = np.random.RandomState(order)
    u = plugins.run_l96_rk4(rng.rand(N), s, 200)[0]
    V, v = rng.rand(m, N), rng.randn(N)
    old_port, old_dev = port.TimeDilation.ORDER, fds_b200.timedilation.ORDER_OF_ACCURACY
    try:
        port.TimeDilation.ORDER = order
        td = port.TimeDilation(plugins.run_l96_rk4, u, s)
        fds_b200.timedilation.set_order_of_accuracy(order)
        eng = fds_b200.L96(N).engine(m)
        eng.set_dxdt_weights(fds_b200.timedilation.stencil_weights(fds_b200.timedilation.ORDER_OF_ACCURACY))
        assert eng.ctx.dilation_order == order
        eng.set_state(u)
        eng.set_tangents(V, v)
        _, _, Gd, gd = eng.boundary(s, 1e-4, from_ensemble=0, do_qr=0, build=0)
        Vp, vp = eng.get_tangents()
        eng.close()
    finally:
        port.TimeDilation.ORDER = old_port
        fds_b200.timedilation.ORDER_OF_ACCURACY = old_dev
    assert maxrel(Gd, td.contribution(V)) < 1e-11 and abs(gd - td.contribution(v)) < 1e-11 * abs(gd) + 1e-14
    assert maxrel(Vp, td.project(V)) < 1e-11 and maxrel(vp, td.project(v)) < 1e-11
