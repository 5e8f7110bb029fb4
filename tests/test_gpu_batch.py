"""BASELINE configs[4]: batched parameter sweep -- B independent Lorenz-96 shadowing problems in
one device context (problems stacked along the site axis, one streaming sweep per time step).
Oracle: a Python loop over single-problem runs, which is all the reference can do
(tests/test_lorenz.py:35-41, apps/lorenz_demo.py:86-89)."""
import numpy as np
import pytest

import fds_b200
from fds_b200.device import Context
from oracle import plugins, shadowing_port as port

gpu = pytest.mark.gpu


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@gpu
@pytest.mark.parametrize('B,N,m,hs', [(3, 64, 4, 8), (5, 256, 16, 16), (2, 4096, 16, 16)])
def test_batched_segment_is_bitwise_the_single_problem_segment(B, N, m, hs):
    """State, tangents and objective sums of one segment, problem by problem, against the
    single-problem context AND the numpy oracle: same bits (exact mode)."""
    rng = np.random.RandomState(B * 1000 + N)
    steps, eps = 2 * hs + 3, 1e-4
    params = np.linspace(-0.5, 0.7, B)
    U = np.stack([plugins.run_l96_rk4(rng.rand(N), s, 30)[0] for s in params])
    V = rng.randn(B, m, N) / np.sqrt(N)
    v = rng.randn(B, N) / np.sqrt(N)
    bc = Context(N, m, batch=B, max_halo_steps=hs)
    bc.set_batch_params(params)
    Ub, Vb, vb, Jb = bc.run_segment_host(U, V, v, 0.0, eps, steps)
    assert Jb.shape == (steps, B, m + 2, 2)
    one = Context(N, m)
    for p in range(B):
        u1, V1, v1, J1 = one.run_segment_host(U[p], V[p], v[p], params[p], eps, steps)
        assert np.array_equal(Ub[p], u1) and np.array_equal(Vb[p], V1) and np.array_equal(vb[p], v1)
        assert np.array_equal(Jb[:, p], J1)
    if N <= 256:
        ru, rV, rv, rJ0, rG, rg = port.run_segment(plugins.run_l96_rk4, U[0], V[0], v[0], params[0], steps, eps)
        assert np.array_equal(Ub[0], ru) and np.array_equal(Vb[0], rV) and np.array_equal(Jb[:, 0, 0] / N, rJ0)


@gpu
def test_shadowing_batch_equals_loop_of_single_runs():
    B, N, m, K, steps, eps = 4, 128, 6, 4, 25, 1e-5
    params = np.array([-0.3, 0.0, 0.1, 0.45])
    rng = np.random.RandomState(5)
    U = np.stack([plugins.run_l96_rk4(rng.rand(N), s, 300)[0] for s in params])
    run = fds_b200.L96(N)
    np.random.seed(11)
    Jb, Gb, H = fds_b200.shadowing_batch(run, U, params, m, K, steps, 0, epsilon=eps, return_histories=True)
    assert Jb.shape == (B, 2) and Gb.shape == (B, 2) and H['Rs'].shape == (B, K, m, m)
    for p in range(B):
        np.random.seed(11)
        np.random.rand(p * m * N)                    # the batch drew its tangents problem after problem
        cp = fds_b200.shadowing(run, U[p], params[p], m, K, steps, 0, epsilon=eps, return_checkpoint=True)
        J = np.array(cp.J)
        assert np.array_equal(H['J'][p, 0], J[0])    # first segment: same initial tangents, same bits
        # later boundaries use differently blocked Gram sums (1e-16), amplified by eps^-1 and chaos
        assert maxrel(H['J'][p], J) < 1e-9
        assert maxrel(H['Rs'][p], np.array(cp.lss.Rs)) < 1e-7
        assert maxrel(H['G_dil'][p], np.array(cp.G_dil)) < 1e-6
        assert maxrel(Jb[p], J.mean((0, 1))) < 1e-10
        assert maxrel(Gb[p], fds_b200.lss_gradient(cp)) < 1e-5


@gpu
def test_shadowing_batch_against_the_oracle_port():
    B, N, m, K, steps, eps = 3, 64, 5, 3, 20, 1e-4
    params = np.array([0.0, 0.1, 0.3])
    u0 = plugins.run_l96_rk4(np.random.RandomState(2).rand(N), 0.1, 500)[0]
    np.random.seed(4)
    Jb, Gb = fds_b200.shadowing_batch(fds_b200.L96(N), u0, params, m, K, steps, 50, epsilon=eps)
    for p in range(B):
        np.random.seed(4)
        np.random.rand(p * m * N)
        Jr, Gr = port.shadowing(plugins.run_l96_rk4, u0, params[p], m, K, steps, 50, epsilon=eps)
        assert maxrel(Jb[p], Jr) < 1e-9
        assert maxrel(Gb[p], Gr) < 1e-2       # QR sign convention: FD in the opposite direction, O(eps)


@gpu
def test_lss_solve_batch_equals_single_solves():
    rng = np.random.RandomState(0)
    B, K, m = 7, 12, 5
    R = rng.randn(B, K, m, m) + 2 * np.eye(m)
    b = rng.randn(B, K, m)
    a = fds_b200.device.lss_solve(R, b)
    for p in range(B):
        assert np.array_equal(a[p], fds_b200.device.lss_solve(R[p], b[p]))


@gpu
def test_batch_rejects_what_it_cannot_do():
    with pytest.raises(fds_b200.FdsError, match='batch'):
        Context(100, 4, batch=3)               # n not a multiple of 64
    with pytest.raises(fds_b200.FdsError):
        Context(3, 2, batch=3, scheme='lorenz63')


@gpu
def test_batched_euler_segment_bitwise():
    """The reference's own Lorenz-96 solver (forward Euler, tests/solvers/mock_fun3d/fun3d:48-57) in a batch."""
    B, N, m, steps, eps = 3, 128, 5, 21, 1e-4
    rng = np.random.RandomState(3)
    params = np.array([0.0, 0.1, -0.2])
    U = np.stack([plugins.run_l96_euler(rng.rand(N), s, 30)[0] for s in params])
    V, v = rng.randn(B, m, N) / np.sqrt(N), rng.randn(B, N) / np.sqrt(N)
    bc = Context(N, m, batch=B, scheme='euler', max_halo_steps=8)
    bc.set_batch_params(params)
    Ub, Vb, vb, Jb = bc.run_segment_host(U, V, v, 0.0, eps, steps)
    for p in range(B):
        ru, rV, rv, rJ0, rG, rg = port.run_segment(plugins.run_l96_euler, U[p], V[p], v[p], params[p], steps, eps)
        assert np.array_equal(Ub[p], ru) and np.array_equal(Vb[p], rV) and np.array_equal(vb[p], rv)
        assert np.array_equal(Jb[:, p, 0] / N, rJ0)


@gpu
def test_sweep_sensitivities_are_consistent_with_the_slope_of_the_averages():
    """Physics check of the whole batched loop (the point of shadowing): over a sweep of forcings the
    mean dJ/ds must match the slope of the long-time averages J(F) themselves."""
    B, N, m, K, steps = 256, 64, 32, 50, 100
    params = np.linspace(-1.0, 1.0, B)
    np.random.seed(1)
    Jb, Gb = fds_b200.shadowing_batch(fds_b200.L96(N), np.random.rand(N), params, m, K, steps, 2000, 1e-4)
    assert np.isfinite(Jb).all() and np.isfinite(Gb).all()
    for q, tol in ((1, 0.2), (0, 0.5)):                  # <x^2> is the well-resolved objective, <x> is noisier
        slope = np.polyfit(params, Jb[:, q], 1)[0]
        assert abs(Gb[:, q].mean() - slope) < tol * abs(slope), (q, Gb[:, q].mean(), slope)


@gpu
def test_batched_fast_math_is_close_to_exact():
    B, N, m, steps, eps = 3, 256, 6, 10, 1e-4
    rng = np.random.RandomState(8)
    params = np.array([0.0, 0.3, -0.3])
    U = np.stack([plugins.run_l96_rk4(rng.rand(N), s, 30)[0] for s in params])
    V, v = rng.randn(B, m, N) / np.sqrt(N), rng.randn(B, N) / np.sqrt(N)
    out = []
    for math in ('exact', 'fast'):
        bc = Context(N, m, batch=B, math=math)
        bc.set_batch_params(params)
        out.append(bc.run_segment_host(U, V, v, 0.0, eps, steps))
        bc.close()
    (ue, Ve, ve, Je), (uf, Vf, vf, Jf) = out
    assert maxrel(uf, ue) < 1e-13 and maxrel(Jf, Je) < 1e-13 and not np.array_equal(uf, ue)
    assert maxrel(Vf, Ve) < 1e-6            # FD tangents amplify rounding by 1 / eps


@gpu
@pytest.mark.parametrize('scheme,run', [('rk4', plugins.run_l96_rk4), ('euler', plugins.run_l96_euler)])
@pytest.mark.parametrize('B,N,steps,hs', [(3, 64, 29, 8), (7, 4096, 50, 16), (2, 192, 5, 3)])
def test_batched_run_up_is_the_oracle_run_of_every_problem(scheme, run, B, N, steps, hs):
    """Run-up of a sweep (shadowing_batch(..., runup_steps)): all problems in one fused multi-step march over the
    stacked sites, ghost cells included; problem p ends where the oracle's run at parameter s + ds[p] ends."""
    rng = np.random.RandomState(B + N)
    params = np.linspace(-0.3, 0.6, B)
    U = rng.rand(B, N) * 2 - 0.5
    bc = Context(N, 2, batch=B, scheme=scheme, max_halo_steps=hs)
    bc.set_batch_params(params)
    bc.set_state(U)
    bc.run(0.05, steps, want_J=False)
    Ub = bc.get_state()
    bc.close()
    for p in range(B):
        assert np.array_equal(Ub[p], run(U[p], 0.05 + params[p], steps)[0])
