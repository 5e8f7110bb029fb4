"""CPU: the oracle (numpy port + plugins) against the fixtures generated from
the real reference by oracle/gen_golden.py.  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import plugins, shadowing_port as port
from conftest import golden

SEG = ['seg_l96rk4_N40_m16', 'seg_l96rk4_N37_m3', 'seg_l96rk4_N1024_m8',
       'seg_l96rk4_N2048_m32', 'seg_l96euler_N40_m16']


def _run_of(name):
    return plugins.run_l96_euler if 'euler' in name else plugins.run_l96_rk4


def maxrel(a, b):
    return float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize('name', SEG)
def test_port_single_segment_bitwise(name):
    g = golden(name)
    u1, V1, v1, J0, G, gg = port.run_segment(_run_of(name), g['u0'], g['V'], g['v'],
                                             float(g['s']), int(g['steps']), float(g['eps']))
    for got, key in ((u1, 'u1'), (V1, 'V1'), (v1, 'v1'), (J0, 'J0'), (np.array(G), 'G'), (gg, 'g')):
        assert np.array_equal(got, g[key]), key


@pytest.mark.parametrize('name', SEG)
def test_port_boundary(name):
    g = golden(name)
    td = port.TimeDilation(_run_of(name), g['u1'], float(g['s']))
    assert np.array_equal(td.dxdt, g['dxdt'])
    assert maxrel(td.contribution(g['V1']), g['G_dil']) < 1e-14
    assert maxrel(td.contribution(g['v1']), g['g_dil']) < 1e-14
    Vp, vp = td.project(g['V1']), td.project(g['v1'])
    assert maxrel(Vp, g['Vp']) < 1e-14 and maxrel(vp, g['vp']) < 1e-14
    lss = port.LssTangent()
    Q, vq = lss.checkpoint(Vp, vp)
    assert maxrel(Q, g['Q']) < 1e-13 and maxrel(vq, g['vq']) < 1e-13
    assert maxrel(lss.Rs[0], g['R']) < 1e-13 and maxrel(lss.bs[0], g['b']) < 1e-13


def test_ensemble_helper_matches_per_trajectory_runs():
    rng = np.random.RandomState(0)
    U = rng.rand(5, 24) * 3
    F = np.array([8.0, 8.0, 8.0, 8.0, 8.0 + 1e-4])
    X, J = plugins.run_l96_ensemble(plugins.l96_rk4_step, U, F, 9)
    for k in range(5):
        x, j = plugins.run_l96_rk4(U[k], F[k] - 8.0, 9)
        assert np.array_equal(x, X[k]) and np.array_equal(j, J[:, k])


def test_tree_sum_partition_independent():
    rng = np.random.RandomState(1)
    x = rng.randn(3, 4096)
    full = plugins.tree_sum(x)
    parts = np.stack([plugins.tree_sum(x[:, i * 512:(i + 1) * 512]) for i in range(8)], -1)
    assert np.array_equal(plugins.tree_sum(parts), full)
    assert np.allclose(full, x.sum(-1), rtol=1e-12)
    y = rng.randn(37)
    assert plugins.tree_sum(y) == plugins.tree_sum(np.r_[y, np.zeros(27)])


def test_kkt_known_answer():
    g = golden('kkt_K50_m16')
    assert np.array_equal(port.lss_solve(g['R'], g['b']), g['alpha'])
    bt = port.lss_solve_block_thomas(g['R'], g['b'])
    assert maxrel(bt, g['alpha']) < 1e-11
    # the solution satisfies the constraint alpha_{i+1} = R_i alpha_i + b_i
    a = g['alpha']
    res = a[1:] - (np.einsum('kij,kj->ki', g['R'][:-1], a[:-1]) + g['b'][:-1])
    assert np.abs(res).max() < 1e-10


def test_dxdt_monomials():
    """reference tests/test_timedilation.py:13-16 (order-k stencil kills x^k/k!)."""
    from math import factorial
    for k in range(2, 10):
        x = np.arange(k + 1) ** k / factorial(k)
        assert abs(port.dxdt_from_states(list(x))) < 1e-12


@pytest.mark.parametrize('name,run', [
    ('cfg1_lorenz63_m2_K20', plugins.run_lorenz63),
    ('circular_m1_K5', plugins.run_circular),
    ('fun3d_l96euler_N40_m16_K20', plugins.run_l96_euler)])
def test_port_full_runs_match_reference(name, run):
    g = golden(name)
    np.random.seed(int(g['seed']))
    J, G = port.shadowing(run, g['u0'], float(g['s']), int(g['m']), int(g['K']),
                          int(g['steps']), int(g['runup']), epsilon=float(g['eps']))
    assert maxrel(J, g['J_ref']) < 1e-12
    assert maxrel(G, g['G_ref']) < 1e-9


def test_reference_windows():
    """The correctness windows the reference's own tests assert."""
    g = golden('fun3d_l96euler_N40_m16_K20')      # tests/test_fun3d.py:131-134
    J, G = g['J_ref'], g['G_ref']
    assert 1 < J[0] < 4 and 10 < J[1] < 40 and -0.1 < G[0] < 0.5 and 1 < G[1] < 8
    g = golden('cfg2_vanderpol_m1_K20')           # tests/test_vanderpol.py:35-36
    assert 1.5 < g['J_ref'][0] < 2.5 and 0.025 < g['G_ref'][0] < 0.035
    g = golden('circular_m1_K5')                  # tests/test_circular.py:41-42 (s=1)
    assert abs(g['J_ref'][0] - 2) < 1e-3 and abs(g['G_ref'][0] - 1) < 0.01
    g = golden('cfg1_lorenz63_m2_K20')            # tests/test_lorenz.py:42-45,65-67
    assert abs(g['J_ref'][1] - 100) < 1e-12 and abs(g['G_ref'][1]) < 1e-12
    L = np.log(abs(np.diagonal(g['Rs'], axis1=1, axis2=2))).mean(0) / (1000 * 0.001)
    assert 0.5 < L[0] < 1.5 and -15 < L[1] < -5


def test_port_restart_equivalence():
    """reference tests/test_restart.py:42-64: 10 segments + continue to 20 equals
    20 straight (J exactly, G allclose)."""
    u0 = np.array([1.0, 0.5])
    np.random.seed(0)
    st = port.shadowing(plugins.run_circular, u0, 1.0, 2, 10, 100, 0, return_state=True)
    st = port.continue_shadowing(plugins.run_circular, 1.0, st, 20, 100)
    J2, G2 = np.array(st.J).mean((0, 1)), port.lss_gradient(st)
    np.random.seed(0)
    J1, G1 = port.shadowing(plugins.run_circular, u0, 1.0, 2, 20, 100, 0)
    assert np.array_equal(J1, J2) and np.allclose(G1, G2)
