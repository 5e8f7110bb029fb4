"""Host-side logic that needs no GPU: post-processing formulas, pinned-pool bookkeeping,
checkpoint naming.  The device solve is replaced by the oracle's numpy solve."""
import numpy as np
import pytest

import fds_b200
from fds_b200 import checkpoint as ckpt, device, fds as core
from fds_b200.lsstan import LssTangent
from oracle import shadowing_port as port
from conftest import golden


def oracle_solve(R, b, dev=0):
    R, b = np.asarray(R), np.asarray(b)
    if b.ndim == 3:
        return np.stack([oracle_solve(R[p], b[p]) for p in range(b.shape[0])])
    lss = port.LssTangent()
    lss.Rs, lss.bs = list(R), list(b)
    return lss.solve()


def random_histories(rng, B, K, m, steps, q):
    return dict(Rs=rng.randn(B, K, m, m) + 2 * np.eye(m), bs=rng.randn(B, K, m), G_lss=rng.randn(B, K, m, q),
                g_lss=rng.randn(B, K, q), J=rng.randn(B, K, steps, q), G_dil=rng.randn(B, K, m), g_dil=rng.randn(B, K))


def test_lss_gradient_batch_is_the_vectorised_lss_gradient(monkeypatch):
    monkeypatch.setattr(device, 'lss_solve', oracle_solve)
    rng = np.random.RandomState(0)
    B, K, m, steps, q = 5, 7, 3, 11, 2
    H = random_histories(rng, B, K, m, steps, q)
    Gb = core.lss_gradient_batch(H['Rs'], H['bs'], H['G_lss'], H['g_lss'], H['J'], H['G_dil'], H['g_dil'])
    assert Gb.shape == (B, q)
    for p in range(B):
        cp = ckpt.Checkpoint(None, None, None, LssTangent(H['Rs'][p], H['bs'][p]), list(H['G_lss'][p]),
                             list(H['g_lss'][p]), list(H['J'][p]), list(H['G_dil'][p]), list(H['g_dil'][p]))
        assert np.allclose(Gb[p], fds_b200.lss_gradient(cp), rtol=1e-13, atol=1e-13)


def test_lss_gradient_matches_the_oracle_port(monkeypatch):
    """fds_b200.lss_gradient against the restated reference formula (fds/fds.py:29-51), incl. segment_range."""
    monkeypatch.setattr(device, 'lss_solve', oracle_solve)
    rng = np.random.RandomState(1)
    K, m, steps, q = 9, 4, 6, 2
    H = {k: v[0] for k, v in random_histories(rng, 1, K, m, steps, q).items()}
    cp = ckpt.Checkpoint(None, None, None, LssTangent(H['Rs'], H['bs']), list(H['G_lss']), list(H['g_lss']),
                         list(H['J']), list(H['G_dil']), list(H['g_dil']))
    st = port.State(None, None, None, port.LssTangent(), list(H['G_lss']), list(H['g_lss']), list(H['J']),
                    list(H['G_dil']), list(H['g_dil'])) if hasattr(port, 'State') else None
    if st is not None:
        st.lss.Rs, st.lss.bs = list(H['Rs']), list(H['bs'])
        assert np.allclose(fds_b200.lss_gradient(cp), port.lss_gradient(st), rtol=1e-13)
    a, b = fds_b200.lss_gradient_series(cp)
    assert a.shape == (K, q) and b.shape == (K, q)
    assert np.allclose(fds_b200.lss_gradient(cp), a.mean(0) + b.mean(0))
    half = fds_b200.lss_gradient(cp, 5)
    assert np.allclose(half, fds_b200.lss_gradient(cp, (0, 5)))


def test_trapez_mean_and_sensitivities():
    rng = np.random.RandomState(2)
    J = rng.randn(10, 6, 2)
    w = np.r_[4.0, 1.0, 2.0 * np.ones(7), 1.0] / 20          # weights [4,1,2,...,2,1]/(2n)
    assert np.allclose(fds_b200.segment.trapez_mean(J[:, 0], 0), (w[:, None] * J[:, 0]).sum(0))
    assert np.array_equal(fds_b200.segment.trapez_mean(J, 0), port.trapez_mean(J, 0))
    J0, G, g = fds_b200.segment.sensitivities(J, 1e-3)
    assert len(G) == 4 and np.allclose(G[2], fds_b200.segment.trapez_mean(J[:, 3] - J[:, 0], 0) / 1e-3)
    assert np.allclose(g, fds_b200.segment.trapez_mean(J[:, 5] - J[:, 0], 0) / 1e-3)


def test_shadowing_batch_rejects_host_plugins_and_single_problems():
    with pytest.raises(ValueError):
        fds_b200.shadowing_batch(fds_b200.L96(64), np.zeros(64), [0.1], 2, 2, 10, 0)
    with pytest.raises(TypeError, match='device plugin'):
        fds_b200.shadowing_batch(lambda u, s, n: (u, None), np.zeros(64), [0.1, 0.2], 2, 2, 10, 0)


def test_device_plugins_are_picklable_like_the_reference_requires():
    import pickle
    for run in (fds_b200.L96(128, scheme='euler'), fds_b200.Lorenz63(), fds_b200.VanDerPol(), fds_b200.Circular()):
        r2 = pickle.loads(pickle.dumps(run))
        assert type(r2) is type(run) and r2.N == run.N and r2.kw == run.kw and r2._engines == {}


@pytest.mark.parametrize('tag', ['cfg1', 'cfg3'])
def test_covariant_lyapunov_vectors_match_the_reference_values(tag):
    """lyapunov_covariant_vectors / ..._magnitude_and_sin_angle / lyapunov_exponents (reference
    fds/lsstan.py:58-85) on the R history of two reference runs: VALUES against the reference's own outputs
    (fixture written by oracle/gen_golden.py clv), not just shapes.  Host-only post-processing."""
    from fds_b200.lsstan import LssTangent
    g = golden('clv_from_reference')
    src = golden(str(g[tag + '_src']))
    lss = LssTangent(list(src['Rs']), list(src['bs']))
    v = lss.lyapunov_covariant_vectors()
    mag, sin = lss.lyapunov_covariant_magnitude_and_sin_angle()
    assert v.shape == g[tag + '_v'].shape
    scale = np.abs(g[tag + '_v']).max()
    assert np.abs(v - g[tag + '_v']).max() <= 1e-12 * scale
    assert np.abs(mag - g[tag + '_mag']).max() <= 1e-12 * np.abs(g[tag + '_mag']).max()
    # sin = sqrt(1 - cos^2) amplifies rounding where the vectors are nearly parallel: compare sin^2
    assert np.allclose(sin ** 2, g[tag + '_sin'] ** 2, rtol=0, atol=1e-10, equal_nan=True)
    assert np.array_equal(lss.lyapunov_exponents(), g[tag + '_exponents'])
    assert np.array_equal(lss.lyapunov_exponents((2, 11)), g[tag + '_exponents_range'])
