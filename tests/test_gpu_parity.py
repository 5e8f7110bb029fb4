"""GPU parity tests: the CUDA hot path (through the C ABI) against the oracle and the
golden fixtures generated from the reference (oracle/gen_golden.py).

Tolerances: time stepping, finite-difference tangents and objective
sums are BIT-IDENTICAL in exact mode; Gram-based quantities (R, b, Q, G_dil, g_dil,
alpha) agree to <= 1e-11 relative after sign normalisation of the QR.
"""
import os
import re

import numpy as np
import pytest

import fds_b200
from fds_b200 import _lib
from oracle import plugins, shadowing_port as port
from conftest import golden, ROOT

gpu = pytest.mark.gpu
SEG_RK4 = ['seg_l96rk4_N40_m16', 'seg_l96rk4_N37_m3', 'seg_l96rk4_N1024_m8', 'seg_l96rk4_N2048_m32']
SEG_ALL = SEG_RK4 + ['seg_l96euler_N40_m16']
TOL = 1e-11          # north star: single segment within 1e-11 relative


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def scheme_of(name):
    return 'euler' if 'euler' in name else 'rk4'


# --------------------------------------------------------------------------- #
# CPU: the C-ABI library loads and exports every symbol the header declares
# --------------------------------------------------------------------------- #
def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'fds_b200.h')).read()
    declared = set(re.findall(r'\b(fds_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b'sm_100a' in lib.fds_version()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'fds_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


# --------------------------------------------------------------------------- #
# the solver plugin: u1, J = run(u0, s, steps)
# --------------------------------------------------------------------------- #
@gpu
@pytest.mark.parametrize('scheme', ['rk4', 'euler'])
@pytest.mark.parametrize('N', [4, 37, 40, 1024, 4099])
def test_run_plugin_bitwise(scheme, N):
    rng = np.random.RandomState(N)
    u0 = rng.rand(N) * 4 - 1
    run = fds_b200.L96(N, scheme=scheme)
    ref = plugins.run_l96_rk4 if scheme == 'rk4' else plugins.run_l96_euler
    for steps in (1, 2, 3, 40):
        u1, J = run(u0, 0.1, steps)
        ur, Jr = ref(u0, 0.1, steps)
        assert np.array_equal(u1, ur), (scheme, N, steps)
        assert np.array_equal(J, Jr), (scheme, N, steps)
    run.close()


@gpu
def test_run_plugin_long_runup_crosses_halo_refresh():
    N = 1024
    u0 = np.random.RandomState(3).rand(N)
    run = fds_b200.L96(N, max_halo_steps=7)
    u1, J = run(u0, 0.0, 100)
    ur, Jr = plugins.run_l96_rk4(u0, 0.0, 100)
    assert np.array_equal(u1, ur) and np.array_equal(J, Jr)


# --------------------------------------------------------------------------- #
# one segment (reference run_segment): bit-identical
# --------------------------------------------------------------------------- #
@gpu
@pytest.mark.parametrize('simple', [False, True])
@pytest.mark.parametrize('name', SEG_ALL)
def test_run_segment_bitwise(name, simple):
    g = golden(name)
    N, m = g['u0'].shape[0], g['V'].shape[0]
    run = fds_b200.L96(N, scheme=scheme_of(name), simple_kernels=simple)
    u1, V1, v1, J0, G, gg = fds_b200.segment.run_segment(
        run, g['u0'], g['V'], g['v'], float(g['s']), 0, int(g['steps']), float(g['eps']))
    for got, key in ((u1, 'u1'), (V1, 'V1'), (v1, 'v1'), (J0, 'J0'), (np.array(G), 'G'), (gg, 'g')):
        assert np.array_equal(got, g[key]), (name, key, maxrel(got, g[key]))
    run.close()


@gpu
@pytest.mark.parametrize('hs', [1, 3, 16])
def test_segment_independent_of_halo_refresh_interval(hs):
    g = golden('seg_l96rk4_N1024_m8')
    run = fds_b200.L96(1024, max_halo_steps=hs)
    u1, V1, v1, J0, G, gg = fds_b200.segment.run_segment(
        run, g['u0'], g['V'], g['v'], float(g['s']), 0, int(g['steps']), float(g['eps']))
    assert np.array_equal(u1, g['u1']) and np.array_equal(V1, g['V1']) and np.array_equal(J0, g['J0'])


@gpu
def test_segment_fast_math_close_not_identical():
    g = golden('seg_l96rk4_N1024_m8')
    run = fds_b200.L96(1024, math='fast')
    u1, V1, v1, J0, G, gg = fds_b200.segment.run_segment(
        run, g['u0'], g['V'], g['v'], float(g['s']), 0, int(g['steps']), float(g['eps']))
    assert maxrel(u1, g['u1']) < 1e-12
    assert maxrel(J0, g['J0']) < 1e-13
    # FD tangents amplify rounding by 1/eps: FMA mode is only statistically equivalent
    assert maxrel(V1, g['V1']) < 1e-6


# --------------------------------------------------------------------------- #
# one boundary: time dilation, projection, QR
# --------------------------------------------------------------------------- #
def _check_boundary(g, R, b, Gd, gd, Q, vq):
    sgn = np.sign(np.diag(g['R']))
    assert (np.diag(R) > 0).all()
    assert maxrel(R, sgn[:, None] * g['R']) < TOL
    assert maxrel(b, sgn * g['b']) < TOL
    assert maxrel(Gd, g['G_dil']) < TOL and abs(gd - g['g_dil']) <= TOL * abs(g['g_dil'])
    assert maxrel(Q, sgn[:, None] * g['Q']) < TOL
    assert maxrel(vq, g['vq']) < TOL


@gpu
@pytest.mark.parametrize('name', SEG_ALL)
def test_boundary_from_explicit_tangents(name):
    g = golden(name)
    N, m = g['u0'].shape[0], g['V'].shape[0]
    eng = fds_b200.L96(N, scheme=scheme_of(name)).engine(m)
    eng.ctx.set_dxdt_weights(fds_b200.timedilation.stencil_weights(3))
    eng.set_state(g['u1'])
    eng.set_tangents(g['V1'], g['v1'])
    R, b, Gd, gd = eng.boundary(float(g['s']), float(g['eps']), from_ensemble=0, do_qr=1, build=0)
    Q, vq = eng.get_tangents()
    _check_boundary(g, R, b, Gd, gd, Q, vq)
    assert np.array_equal(eng.get_state(), g['u1'])
    eng.close()


@gpu
@pytest.mark.parametrize('name', SEG_ALL)
def test_segment_then_boundary_device_resident(name):
    """The production sequence: ensemble -> segment -> boundary straight from the ensemble
    (tangents never materialised), then the next perturbed ensemble."""
    g = golden(name)
    N, m = g['u0'].shape[0], g['V'].shape[0]
    s, eps, steps = float(g['s']), float(g['eps']), int(g['steps'])
    eng = fds_b200.L96(N, scheme=scheme_of(name)).engine(m)
    eng.ctx.set_dxdt_weights(fds_b200.timedilation.stencil_weights(3))
    eng.set_time_dilation(False)
    eng.set_state(g['u0'])
    eng.set_tangents(g['V'], g['v'])
    eng.boundary(s, eps, from_ensemble=0, do_qr=0, build=1)      # identity (no dilation): just builds the ensemble
    eng.set_time_dilation(True)
    J = eng.segment(s, eps, steps)
    J0, G, gg = fds_b200.segment.sensitivities(J, eps)
    assert np.array_equal(J0, g['J0']) and np.array_equal(np.array(G), g['G']) and np.array_equal(gg, g['g'])
    R, b, Gd, gd = eng.boundary(s, eps, from_ensemble=1, do_qr=1, build=1)
    Q, vq = eng.get_tangents()
    _check_boundary(g, R, b, Gd, gd, Q, vq)
    assert np.array_equal(eng.get_state(), g['u1'])
    # the next ensemble is u + eps * [Q; vq]: advance it one step and compare with the oracle
    J1 = eng.segment(s, eps, 1)
    U = np.vstack([g['u1'], g['u1'] + eps * Q, g['u1'] + eps * vq])
    F = s + 8.0 + np.r_[np.zeros(m + 1), eps]
    step = plugins.l96_euler_step if 'euler' in name else plugins.l96_rk4_step
    Xo, Jo = plugins.run_l96_ensemble(step, U, F, 1)
    assert maxrel(J1, Jo) < 1e-13
    eng.close()


@gpu
@pytest.mark.parametrize('scheme', ['rk4', 'euler'])
@pytest.mark.parametrize('H,steps', [(3, 7), (5, 11), (6, 13), (8, 8), (128, 10)])
def test_mixed_pass_depths_bitwise(scheme, H, steps):
    """Temporal blocking under a small halo capacity: a segment is cut into halo chunks of at most H steps, and a
    chunk into passes of four (forward Euler), two and one time steps -- every mixture gives the oracle's ensemble
    and objective sums bit for bit."""
    N, m, s, eps = 4096, 4, 0.1, 1e-4
    rng = np.random.RandomState(11)
    u0 = plugins.run_l96_rk4(rng.rand(N), s, 40)[0]
    V = np.linalg.qr(rng.rand(m, N).T)[0].T
    v = rng.randn(N) / np.sqrt(N)
    eng = fds_b200.L96(N, scheme=scheme, max_halo_steps=H).engine(m)
    eng.set_time_dilation(False)
    eng.set_state(u0)
    eng.set_tangents(V, v)
    eng.boundary(s, eps, from_ensemble=0, do_qr=0, build=1)      # no dilation, no QR: just builds u0 + eps [V; v]
    J = eng.segment(s, eps, steps)
    U = np.vstack([u0, u0 + eps * V, u0 + eps * v])
    F = np.r_[np.full(m + 1, s + 8.0), (s + eps) + 8.0]         # fds/segment.py:64: parameter + epsilon, then F = s + 8
    step = plugins.l96_euler_step if scheme == 'euler' else plugins.l96_rk4_step
    Xo, Jo = plugins.run_l96_ensemble(step, U, F, steps)
    assert np.array_equal(J, Jo)
    assert np.array_equal(eng.get_state(), Xo[0])
    Vd, vd = eng.get_tangents()
    assert np.array_equal(Vd, (Xo[1:m + 1] - Xo[0]) / eps) and np.array_equal(vd, (Xo[m + 1] - Xo[0]) / eps)
    eng.close()


@gpu
def test_boundary_projection_only_matches_reference():
    g = golden('seg_l96rk4_N1024_m8')
    eng = fds_b200.L96(1024).engine(8)
    eng.ctx.set_dxdt_weights(fds_b200.timedilation.stencil_weights(3))
    eng.set_state(g['u1'])
    eng.set_tangents(g['V1'], g['v1'])
    _, _, Gd, gd = eng.boundary(float(g['s']), float(g['eps']), from_ensemble=0, do_qr=0, build=0)
    Vp, vp = eng.get_tangents()
    assert maxrel(Vp, g['Vp']) < TOL and maxrel(vp, g['vp']) < TOL
    assert maxrel(Gd, g['G_dil']) < TOL


@gpu
def test_rank_deficient_tangents_fail_loudly():
    eng = fds_b200.L96(64).engine(4)
    eng.set_state(np.random.RandomState(0).rand(64))
    V = np.random.RandomState(1).rand(4, 64)
    V[3] = V[2]
    eng.set_tangents(V, None)
    with pytest.raises(fds_b200.FdsError, match='Cholesky'):
        eng.boundary(0.0, 1e-4, from_ensemble=0, do_qr=1, build=0)


# --------------------------------------------------------------------------- #
# LssTangent.solve
# --------------------------------------------------------------------------- #
@gpu
def test_lss_solve_matches_reference_superlu():
    g = golden('kkt_K50_m16')
    alpha = fds_b200.device.lss_solve(g['R'], g['b'])
    assert maxrel(alpha, g['alpha']) < TOL
    # known-answer property: the constraint alpha_{i+1} = R_i alpha_i + b_i
    res = np.einsum('kij,kj->ki', g['R'][:-1], alpha[:-1]) + g['b'][:-1] - alpha[1:]
    assert np.abs(res).max() < 1e-10


@gpu
@pytest.mark.parametrize('K,m', [(1, 1), (3, 2), (200, 32), (20, 64)])
def test_lss_solve_shapes(K, m):
    rng = np.random.RandomState(K * 100 + m)
    # upper-triangular blocks of a uniformly hyperbolic system (half expanding, half contracting,
    # weak coupling), so the Schur complement is well conditioned independently of K
    diag = np.where(np.arange(m) < (m + 1) // 2, 2.0 + rng.rand(K, m), 0.2 + 0.2 * rng.rand(K, m))
    R = np.triu(0.03 * rng.randn(K, m, m), 1) + diag[:, :, None] * np.eye(m)
    b = rng.randn(K, m)
    alpha = fds_b200.device.lss_solve(R, b)
    assert maxrel(alpha, port.lss_solve(R, b)) < 1e-10


@gpu
def test_lss_solve_on_reference_history():
    g = golden('cfg3_l96rk4_N40_m16_K100')
    sgn = np.sign(np.diagonal(g['Rs'], axis1=1, axis2=2))
    alpha = fds_b200.device.lss_solve(g['Rs'], g['bs'])
    assert maxrel(alpha, g['alpha_ref']) < 1e-9      # kappa of the 100-block Schur complement


# --------------------------------------------------------------------------- #
# adaptive CholeskyQR: one certified pass vs CholeskyQR2
# --------------------------------------------------------------------------- #
@gpu
@pytest.mark.parametrize('name', ['seg_l96rk4_N40_m16', 'seg_l96rk4_N2048_m32'])
def test_single_pass_qr_agrees_with_choleskyqr2(name):
    g = golden(name)
    N, m = g['u0'].shape[0], g['V'].shape[0]
    out = {}
    for passes in (0, 2):
        eng = fds_b200.L96(N).engine(m)
        eng.ctx.set_qr_passes(passes)
        eng.ctx.set_dxdt_weights(fds_b200.timedilation.stencil_weights(3))
        eng.set_state(g['u1'])
        eng.set_tangents(g['V1'], g['v1'])
        launches = eng.ctx.launches
        R, b, Gd, gd = eng.boundary(float(g['s']), float(g['eps']), from_ensemble=0, do_qr=1, build=0)
        out[passes] = (R, b, Gd, gd) + eng.get_tangents() + (eng.ctx.last_qr_passes, eng.ctx.last_condition)
        eng.close()
    one, two = out[0], out[2]
    assert one[7] >= 1.0
    if name == 'seg_l96rk4_N2048_m32':
        assert one[7] < 29.5 and one[6] == 1 and two[6] == 2    # certified: the single pass was taken
    for a, b in zip(one[:6], two[:6]):
        assert maxrel(a, b) < 1e-12
    Q = one[4]
    assert np.abs(Q @ Q.T - np.eye(m)).max() < 1e-13
    _check_boundary(g, *one[:6])


@gpu
def test_ill_conditioned_block_falls_back_to_two_passes():
    """Nearly parallel tangents (cond ~ 1e3): cond^2 u is not negligible, the single pass is refused
    and CholeskyQR2 still orthonormalises to rounding."""
    N, m = 512, 8
    eng = fds_b200.L96(N).engine(m)
    eng.set_state(np.random.RandomState(0).rand(N))
    rng = np.random.RandomState(1)
    W = rng.rand(N) + 2e-3 * rng.rand(m, N)
    eng.set_tangents(W, np.zeros(N))
    eng.set_time_dilation(False)
    R, b, _, _ = eng.boundary(0.0, 1e-4, from_ensemble=0, do_qr=1, build=0)
    assert eng.ctx.last_condition > 30 and eng.ctx.last_qr_passes == 2
    Q, _ = eng.get_tangents()
    assert np.abs(Q @ Q.T - np.eye(m)).max() < 1e-13
    assert maxrel(R.T @ Q, W) < 1e-12
    eng.close()


@gpu
@pytest.mark.parametrize('N,m,cond', [(5, 5, 1e9), (300, 6, 1e11), (4096, 12, 1e11)])
def test_numerically_singular_gram_takes_shifted_choleskyqr3(N, m, cond):
    """A tangent block whose Gram is singular in fp64 (cond_2 of the block far beyond 1e8: CholeskyQR2 breaks down)
    -- what m = N with a wide Lyapunov spectrum produces -- is still factored: shifted CholeskyQR3 agrees with
    Householder QR (np.linalg.qr, what the reference calls at pascal_lite/operators/linalg.py:21) on R to
    cond * u and orthonormalises to rounding."""
    rng = np.random.RandomState(m)
    U = np.linalg.qr(rng.randn(N, m))[0]
    W = (np.linalg.qr(rng.randn(m, m))[0] * np.logspace(0, -np.log10(cond), m)) @ U.T         # (m, N), singular values 1 .. 1/cond
    eng = fds_b200.L96(max(N, 4)).engine(m) if N >= 4 else None
    eng.set_state(rng.rand(N))
    eng.set_tangents(W, np.zeros(N))
    eng.set_time_dilation(False)
    R, b, _, _ = eng.boundary(0.0, 1e-4, from_ensemble=0, do_qr=1, build=0)
    assert eng.ctx.last_qr_passes == 3
    Q, _ = eng.get_tangents()
    assert np.abs(Q @ Q.T - np.eye(m)).max() < 1e-12
    assert maxrel(R.T @ Q, W) < 1e-12                                 # backward error of the factorisation
    Rr = np.linalg.qr(W.T)[1]
    sgn = np.sign(np.diag(Rr))
    assert maxrel(R, sgn[:, None] * Rr) < 1e-4 * cond * 1e-12 + 1e-9   # forward error ~ cond * u
    eng.close()


@gpu
@pytest.mark.parametrize('scheme,run', [('rk4', plugins.run_l96_rk4), ('euler', plugins.run_l96_euler)])
@pytest.mark.parametrize('N,steps,m', [(40, 1, 2), (40, 37, 2), (1000, 8, 6), (4099, 9, 6), (50000, 23, 1),
                                       (50000, 300, 16), (1 << 20, 11, 8)])
def test_primal_march_is_the_oracle_run(scheme, run, N, steps, m):
    """The solver call ``u1, J = run(u0, s, steps)`` and the run-up (fds/fds.py:207, J dropped) go through the fused
    multi-step single-column kernel, up to 8 steps per pass, the recorded one with the 64-site leaf sums of the
    objective of every step: the state and J are the oracle's bit for bit."""
    s = 0.1
    u0 = np.random.RandomState(N + steps).rand(N) * 2 - 0.5
    eng = fds_b200.L96(N, scheme=scheme).engine(m)
    eng.set_state(u0)
    eng.run(s, steps, want_J=False)
    a = eng.get_state().copy()
    eng.set_state(u0)
    J = eng.run(s, steps, want_J=True)
    b = eng.get_state().copy()
    eng.close()
    u1, Jo = run(u0, s, steps)
    assert np.array_equal(a, b) and np.array_equal(a, u1)
    assert np.array_equal(J, Jo)
