"""CPU: the per-thread body of the streaming CUDA step kernel (l96_march.cuh,
l96_core.cuh), compiled for the host, against the numpy oracle -- bit for bit.
This validates ring indexing, warm-up, strip planning, halo shrinkage and the
objective reduction tree without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import plugins

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'fds_b200', 'csrc', 'emul_host.cpp')
SO = os.path.join(ROOT, 'tests', '_emul.so')
c_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope='module')
def emul():
    deps = [SRC] + [os.path.join(ROOT, 'fds_b200', 'csrc', f)
                    for f in ('l96_core.cuh', 'l96_march.cuh', 'l96_march2.cuh', 'l96_marchk.cuh', 'common.cuh')]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++',
                               SRC, '-o', SO])
    lib = ctypes.CDLL(SO)
    lib.emul_tree_reduce.restype = ctypes.c_double
    return lib


def ptr(a, off=0):
    return ctypes.cast(a.ctypes.data + 8 * off, c_dp)


LL = ctypes.c_longlong


def plan_of(emul, N, H, max_strips):
    out = (LL * 8)()
    emul.emul_plan(LL(N), LL(H), LL(max_strips), out)
    return dict(zip(('a0', 'ls', 'nstrips', 'nblk', 'alo', 'ahi', 'node_shift', 'nnode'), out))


def padded(U, alo, ahi, ghost_l, ghost_r):
    """[site][M] array for sites [alo, ahi): owned sites + periodic ghosts [-ghost_l, 0) and
    [N, N + ghost_r); everything else NaN, as cells outside the valid halo may hold anything."""
    M, N = U.shape
    buf = np.full((ahi - alo, M), np.nan)
    for g in range(-ghost_l, N + ghost_r):
        buf[g - alo] = U[:, g % N]
    return buf


@pytest.mark.parametrize('N,M,H,max_strips,exact', [
    (200, 5, 3, 7, 1), (40, 18, 1, 4, 1), (37, 3, 2, 64, 1), (1024, 4, 4, 16, 1),
    (4096, 2, 2, 37, 1), (333, 1, 3, 1, 1), (200, 5, 3, 7, 0), (4, 2, 3, 5, 1), (640, 3, 5, 3, 1), (5000, 2, 2, 2, 1), (3000, 1, 1, 1, 1), (70000, 1, 2, 5, 1)])
@pytest.mark.parametrize('scheme', ['rk4', 'euler'])
def test_streaming_march_matches_oracle(emul, N, M, H, max_strips, exact, scheme):
    step_fn = emul.emul_l96_rk4_step if scheme == 'rk4' else emul.emul_l96_euler_step
    oracle_step = plugins.l96_rk4_step if scheme == 'rk4' else plugins.l96_euler_step
    rng = np.random.RandomState(N + M)
    U = rng.rand(M, N) * 6 - 1
    F = 8.0 + np.r_[np.zeros(M - 1), 1e-4] if M > 1 else np.array([8.1])
    pl = plan_of(emul, N, H, max_strips)
    alo, ahi = pl['alo'], pl['ahi']
    g = 1 << pl['node_shift']
    assert pl['nstrips'] <= max_strips and pl['ls'] % g == 0 and pl['a0'] % g == 0 and g >= 64
    assert pl['nnode'] == (N + g - 1) // g
    assert alo <= -8 * H and ahi >= N + 4 * H and alo % 8 == 4 and ahi % 8 == 4
    assert pl['a0'] + pl['nstrips'] * pl['ls'] >= N + 4 * (H - 1)
    bufA = padded(U, alo, ahi, 8 * H, 4 * H)
    bufB = np.full_like(bufA, np.nan)
    nleaf = pl['nnode']
    Xo, Jo = plugins.run_l96_ensemble(oracle_step, U, F, H)
    for j in range(H):
        jpart = np.full((nleaf, M, 2), np.nan)
        with np.errstate(all='ignore'):
            step_fn(ptr(bufA, -alo * M), ptr(bufB, -alo * M), M, LL(N), LL(H), ptr(F),
                                   ctypes.c_double(0.01), exact, LL(max_strips), ptr(jpart))
        assert not np.isnan(jpart).any()
        # still-valid range after step j+1: [-8(H-1-j), N + 4(H-1-j))
        lo, hi = -8 * (H - 1 - j), N + 4 * (H - 1 - j)
        assert not np.isnan(bufB[lo - alo:hi - alo]).any()
        if exact:
            for k in range(M):
                for q in range(2):
                    col = np.ascontiguousarray(jpart[:, k, q])
                    got = emul.emul_tree_reduce(ptr(col), LL(nleaf), LL(1))
                    assert got / N == Jo[j, k, q]
        bufA, bufB = bufB, bufA
    got = bufA[-alo:-alo + N].T
    if exact:
        assert np.array_equal(got, Xo)
    else:
        assert np.abs(got - Xo).max() < 1e-13 and (scheme == 'euler' or not np.array_equal(got, Xo))


@pytest.mark.parametrize('scheme', ['rk4', 'euler'])
@pytest.mark.parametrize('N,M,H,max_strips,exact', [
    (200, 5, 4, 7, 1), (40, 18, 2, 4, 1), (37, 3, 2, 64, 1), (1024, 4, 4, 16, 1), (4096, 2, 2, 37, 1),
    (333, 1, 3, 1, 1), (200, 5, 4, 7, 0), (4, 2, 3, 5, 1), (640, 3, 6, 3, 1), (5000, 2, 2, 2, 1),
    (64, 2, 2, 1, 1), (128, 1, 2, 2, 1), (70000, 1, 2, 5, 1)])
def test_two_steps_per_pass_march_matches_oracle(emul, N, M, H, max_strips, exact, scheme):
    """l96_march2.cuh: two chained pipes (RK4, or the reference's forward Euler), both steps' objective
    partials -- bit for bit two oracle steps."""
    step_fn = plugins.l96_rk4_step if scheme == 'rk4' else plugins.l96_euler_step
    emul_fn = emul.emul_l96_rk4_step2 if scheme == 'rk4' else emul.emul_l96_euler_step2
    rng = np.random.RandomState(N + M + 1)
    U = rng.rand(M, N) * 6 - 1
    F = 8.0 + np.r_[np.zeros(M - 1), 1e-4] if M > 1 else np.array([8.1])
    pl = plan_of(emul, N, H, max_strips)
    alo, ahi = pl['alo'], pl['ahi']
    assert alo <= pl['a0'] - 20 and ahi >= pl['a0'] + pl['nstrips'] * pl['ls'] + 20
    bufA = padded(U, alo, ahi, 8 * H, 4 * H)
    bufB = np.full_like(bufA, np.nan)
    nleaf = pl['nnode']
    npass = H // 2
    Xo, Jo = plugins.run_l96_ensemble(step_fn, U, F, 2 * npass)
    for j in range(npass):
        jp = [np.full((nleaf, M, 2), np.nan) for _ in range(2)]
        with np.errstate(all='ignore'):
            emul_fn(ptr(bufA, -alo * M), ptr(bufB, -alo * M), M, LL(N), LL(H), ptr(F),
                    ctypes.c_double(0.01), exact, LL(max_strips), ptr(jp[0]), ptr(jp[1]))
        done = 2 * (j + 1)
        lo, hi = -8 * (H - done), N + 4 * (H - done)
        assert not np.isnan(bufB[lo - alo:hi - alo]).any()
        for st in range(2):
            assert not np.isnan(jp[st]).any()
            if exact:
                for k in range(M):
                    for q in range(2):
                        col = np.ascontiguousarray(jp[st][:, k, q])
                        assert emul.emul_tree_reduce(ptr(col), LL(nleaf), LL(1)) / N == Jo[2 * j + st, k, q]
        bufA, bufB = bufB, bufA
    got = bufA[-alo:-alo + N].T
    if exact:
        assert np.array_equal(got, Xo)
    else:
        assert np.abs(got - Xo).max() < 1e-13


@pytest.mark.parametrize('scheme,K', [('euler', 4), ('rk4', 2)])
@pytest.mark.parametrize('N,M,H,max_strips,exact', [
    (200, 5, 4, 7, 1), (40, 18, 4, 4, 1), (37, 3, 4, 64, 1), (1024, 4, 8, 16, 1), (4096, 2, 4, 37, 1),
    (333, 1, 5, 1, 1), (200, 5, 4, 7, 0), (8, 2, 4, 5, 1), (640, 3, 8, 3, 1), (5000, 2, 4, 2, 1),
    (64, 2, 4, 1, 1), (128, 1, 4, 2, 1), (70000, 1, 4, 5, 1)])
def test_k_steps_per_pass_march_matches_oracle(emul, N, M, H, max_strips, exact, scheme, K):
    """l96_marchk.cuh: a chain of K pipes (four forward-Euler steps per pass; the generic march with K = 2
    on the RK4 pipe), all K steps' objective partials -- bit for bit K oracle steps."""
    step_fn = plugins.l96_rk4_step if scheme == 'rk4' else plugins.l96_euler_step
    emul_fn = emul.emul_l96_rk4_stepk2 if scheme == 'rk4' else emul.emul_l96_euler_step4
    rng = np.random.RandomState(N + M + 1)
    U = rng.rand(M, N) * 6 - 1
    F = 8.0 + np.r_[np.zeros(M - 1), 1e-4] if M > 1 else np.array([8.1])
    pl = plan_of(emul, N, H, max_strips)
    alo, ahi = pl['alo'], pl['ahi']
    assert alo <= pl['a0'] - 12 - 8 * (K - 1) and ahi >= pl['a0'] + pl['nstrips'] * pl['ls'] + 12 + 8 * (K - 1)
    bufA = padded(U, alo, ahi, 8 * H, 4 * H)
    bufB = np.full_like(bufA, np.nan)
    nleaf = pl['nnode']
    npass = H // K
    Xo, Jo = plugins.run_l96_ensemble(step_fn, U, F, K * npass)
    for j in range(npass):
        jp = [np.full((nleaf, M, 2), np.nan) for _ in range(K)]
        jpp = (ctypes.POINTER(ctypes.c_double) * K)(*[ptr(a) for a in jp])
        with np.errstate(all='ignore'):
            emul_fn(ptr(bufA, -alo * M), ptr(bufB, -alo * M), M, LL(N), LL(H), ptr(F),
                    ctypes.c_double(0.01), exact, LL(max_strips), jpp)
        done = K * (j + 1)
        lo, hi = -8 * (H - done), N + 4 * (H - done)
        assert not np.isnan(bufB[lo - alo:hi - alo]).any()
        for st in range(K):
            assert not np.isnan(jp[st]).any()
            if exact:
                for k in range(M):
                    for q in range(2):
                        col = np.ascontiguousarray(jp[st][:, k, q])
                        assert emul.emul_tree_reduce(ptr(col), LL(nleaf), LL(1)) / N == Jo[K * j + st, k, q]
        bufA, bufB = bufB, bufA
    got = bufA[-alo:-alo + N].T
    if exact:
        assert np.array_equal(got, Xo)
    else:
        assert np.abs(got - Xo).max() < 1e-13


def test_point_forms_match_oracle(emul):
    rng = np.random.RandomState(5)
    N, M = 64, 3
    U = rng.rand(M, N) * 6 - 1
    F = np.array([8.0, 8.0, 8.0001])
    alo, ahi = -12, N + 12
    buf = padded(U, alo, ahi, 8, 4)
    out = np.zeros_like(buf)
    emul.emul_l96_rk4_point(ptr(buf, -alo * M), ptr(out, -alo * M), M, ctypes.c_longlong(0),
                            ctypes.c_longlong(N), ptr(F), ctypes.c_double(0.01), 1)
    assert np.array_equal(out[-alo:-alo + N].T, plugins.l96_rk4_step(U, F[:, None]))
    emul.emul_l96_euler_point(ptr(buf, -alo * M), ptr(out, -alo * M), M, ctypes.c_longlong(0),
                              ctypes.c_longlong(N), ptr(F), ctypes.c_double(0.01), 1)
    assert np.array_equal(out[-alo:-alo + N].T, plugins.l96_euler_step(U, F[:, None]))
