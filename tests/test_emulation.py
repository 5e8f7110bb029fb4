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
                    for f in ('l96_core.cuh', 'l96_march.cuh', 'common.cuh')]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++',
                               SRC, '-o', SO])
    lib = ctypes.CDLL(SO)
    lib.emul_tree_reduce.restype = ctypes.c_double
    return lib


def ptr(a, off=0):
    return ctypes.cast(a.ctypes.data + 8 * off, c_dp)


def padded(U, H):
    """[site][M] array for sites [alo, ahi) with periodic ghosts; returns (buf, alo, ahi)."""
    M, N = U.shape
    alo = -(8 * H + 12)
    ahi = (N + 4 * H + 4 + 7) // 8 * 8 + 4
    buf = np.zeros((ahi - alo, M))
    idx = np.arange(alo, ahi) % N
    buf[:] = U.T[idx]
    return buf, alo, ahi


@pytest.mark.parametrize('N,M,H,max_strips,exact', [
    (200, 5, 3, 7, 1), (40, 18, 1, 4, 1), (37, 3, 2, 64, 1), (1024, 4, 4, 16, 1),
    (4096, 2, 2, 37, 1), (333, 1, 3, 1, 1), (200, 5, 3, 7, 0)])
def test_streaming_march_matches_oracle(emul, N, M, H, max_strips, exact):
    rng = np.random.RandomState(N + M)
    U = rng.rand(M, N) * 6 - 1
    F = 8.0 + np.r_[np.zeros(M - 1), 1e-4] if M > 1 else np.array([8.1])
    bufA, alo, ahi = padded(U, H)
    bufB = np.zeros_like(bufA)
    nleaf = (N + 63) // 64
    Xo, Jo = plugins.run_l96_ensemble(plugins.l96_rk4_step, U, F, H)
    plan = (ctypes.c_longlong * 3)()
    for j in range(H):
        lo, hi = -8 * (H - 1 - j), N + 4 * (H - 1 - j)
        jpart = np.full((nleaf, M, 2), np.nan)
        emul.emul_l96_rk4_step(ptr(bufA, -alo * M), ptr(bufB, -alo * M), M,
                               ctypes.c_longlong(N), ctypes.c_longlong(lo), ctypes.c_longlong(hi),
                               ctypes.c_longlong(alo), ctypes.c_longlong(ahi), ptr(F),
                               ctypes.c_double(0.01), exact, ctypes.c_longlong(max_strips),
                               ptr(jpart), plan)
        assert plan[2] <= max_strips and plan[1] % 64 == 0 and plan[0] % 64 == 0
        assert not np.isnan(jpart).any()
        if exact:
            for k in range(M):
                for q in range(2):
                    col = np.ascontiguousarray(jpart[:, k, q])
                    got = emul.emul_tree_reduce(ptr(col), ctypes.c_longlong(nleaf), ctypes.c_longlong(1))
                    assert got / N == Jo[j, k, q]
        bufA, bufB = bufB, bufA
    got = bufA[-alo:-alo + N].T
    if exact:
        assert np.array_equal(got, Xo)
    else:
        assert np.abs(got - Xo).max() < 1e-13 and not np.array_equal(got, Xo)


def test_point_forms_match_oracle(emul):
    rng = np.random.RandomState(5)
    N, M = 64, 3
    U = rng.rand(M, N) * 6 - 1
    F = np.array([8.0, 8.0, 8.0001])
    buf, alo, ahi = padded(U, 1)
    out = np.zeros_like(buf)
    emul.emul_l96_rk4_point(ptr(buf, -alo * M), ptr(out, -alo * M), M, ctypes.c_longlong(0),
                            ctypes.c_longlong(N), ptr(F), ctypes.c_double(0.01), 1)
    assert np.array_equal(out[-alo:-alo + N].T, plugins.l96_rk4_step(U, F[:, None]))
    emul.emul_l96_euler_point(ptr(buf, -alo * M), ptr(out, -alo * M), M, ctypes.c_longlong(0),
                              ctypes.c_longlong(N), ptr(F), ctypes.c_double(0.01), 1)
    assert np.array_equal(out[-alo:-alo + N].T, plugins.l96_euler_step(U, F[:, None]))
