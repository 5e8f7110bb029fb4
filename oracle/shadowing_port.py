"""Eager-numpy restatement of the reference's shadowing segment loop.
TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by fds_b200.

The reference evaluates the same numpy operations through a lazy graph
(pascal_lite); here they are executed eagerly, in the same order, so results
agree with the reference to rounding (pinned by oracle/gen_golden.py against
/root/reference; fixtures in tests/golden/).  Citations are relative to
/root/reference.
"""
from __future__ import annotations

import multiprocessing
from dataclasses import dataclass, field

import numpy as np
from scipy import sparse
from scipy.sparse import linalg as splinalg


# --------------------------------------------------------------------------- #
# small utilities
# --------------------------------------------------------------------------- #
def trapez_mean(J, axis=0):
    """fds/segment.py:9-12 -- trapezoid time mean with the extrapolated sample
    J[-1] = 2 J[0] - J[1] in front."""
    J = np.moveaxis(np.asarray(J), axis, 0)
    ghost = 2 * J[0] - J[1]
    return (J.sum(0) + J[:-1].sum(0) + ghost) / (2 * J.shape[0])


def mean_std(x):
    """fds/timeseries.py:3-7 -- mean and the std of 8 batch means / sqrt(8)."""
    x = np.asarray(x)
    edges = np.array(np.linspace(0, len(x), 9), int)
    batches = np.array([x[a:b].mean(0) for a, b in zip(edges[:-1], edges[1:])])
    return x.mean(0), batches.std(0) / np.sqrt(len(batches))


def windowed_mean(a):
    """fds/timeseries.py:9-12 -- the window is all ones: a plain mean."""
    a = np.asarray(a)
    w = np.ones(a.shape[0])
    return (a * w[:, None]).sum(0) / w.sum()


def fd_stencil(order):
    """fds/timedilation.py:16-19 -- one-sided first-derivative weights."""
    A = np.array([np.arange(order + 1) ** i for i in range(order + 1)])
    rhs = np.zeros(order + 1)
    rhs[1] = 1
    return np.linalg.solve(A, rhs)


def dxdt_from_states(states):
    """fds/timedilation.py:14-35 -- d = sum_i c_i u_i, accumulated left to right
    from Python's integer 0 (0 + x is exact)."""
    c = fd_stencil(len(states) - 1)
    acc = 0
    for ci, ui in zip(c, states):
        acc = acc + ci * ui
    return acc


def rowdot(X, y):
    """pascal_lite/operators/linalg.py:36 -- (x*y).sum(-1) (numpy pairwise)."""
    return (X * y).sum(-1)


# --------------------------------------------------------------------------- #
# time dilation, fds/timedilation.py:37-82
# --------------------------------------------------------------------------- #
class TimeDilation:
    ORDER = 3  # fds/timedilation.py:64

    def __init__(self, run, u0, parameter, pool=None):
        calls = [(u0, parameter, k) for k in range(1, self.ORDER + 1)]
        if pool is None:
            ends = [run(*c)[0] for c in calls]
        else:
            ends = [r[0] for r in pool.starmap(run, calls)]
        self.dxdt = dxdt_from_states([u0] + list(ends))
        self.dxdt_normalized = self.dxdt / rowdot(self.dxdt, self.dxdt) ** 0.5

    def contribution(self, V):               # fds/timedilation.py:38-45
        return rowdot(V, self.dxdt) / rowdot(self.dxdt, self.dxdt)

    def project(self, V):                    # fds/timedilation.py:47-52
        c = np.asarray(rowdot(V, self.dxdt_normalized))
        return V - c[..., None] * self.dxdt_normalized


class NoTimeDilation:
    """run_ddt=0 path (fds/timedilation.py:59-61, 39-43, 48-49)."""
    dxdt = None

    def contribution(self, V):
        return rowdot(V, V[0] * 0) if V.ndim > 1 else rowdot(V, V * 0)

    def project(self, V):
        return V


# --------------------------------------------------------------------------- #
# LssTangent, fds/lsstan.py
# --------------------------------------------------------------------------- #
@dataclass
class LssTangent:
    Rs: list = field(default_factory=list)
    bs: list = field(default_factory=list)

    def K_segments(self):
        return len(self.Rs)

    def m_modes(self):
        return self.Rs[0].shape[0]

    def checkpoint(self, V, v):
        """fds/lsstan.py:20-34 -> pascal_lite/operators/linalg.py:15-26:
        V.T = Q0 R (LAPACK Householder, no sign normalisation), Q = Q0.T."""
        Q0, R = np.linalg.qr(V.T)
        Q = Q0.T
        b = rowdot(Q, v)
        v = v - (b[:, None] * Q).sum(0)
        self.Rs.append(R)
        self.bs.append(b)
        return Q, v

    def solve(self):
        return lss_solve(np.array(self.Rs), np.array(self.bs))

    def lyapunov_exponents(self):            # fds/lsstan.py:58-64
        R = np.array(self.Rs)
        i = np.arange(self.m_modes())
        return np.log(abs(R[:, i, i]))


def lss_solve(R, b):
    """fds/lsstan.py:36-50 -- min |alpha|^2 s.t. alpha_{i+1} = R_i alpha_i + b_i
    via the Schur complement B B^T (scipy SuperLU)."""
    K, m = b.shape
    eyes = np.eye(m) * np.ones([K, 1, 1])
    shape = (m * K, m * (K + 1))
    I = sparse.bsr_matrix((eyes, np.r_[1:K + 1], np.r_[:K + 1]))
    D = sparse.bsr_matrix((R, np.r_[:K], np.r_[:K + 1]), shape=shape)
    B = (D - I).tocsr()
    alpha = -(B.T * splinalg.spsolve(B * B.T, np.ravel(b)))
    return alpha.reshape([K + 1, -1])[:-1]


def lss_solve_block_thomas(R, b):
    """Independent dense restatement used as a known-answer cross-check of
    lss_solve (the reference has no test that isolates solve()): block
    tridiagonal Cholesky of S = B B^T with S_ii = R_i R_i^T + I,
    S_{i+1,i} = -R_{i+1};  alpha_i = lambda_{i-1} - R_i^T lambda_i."""
    K, m = b.shape
    C = np.zeros((K, m, m))      # diagonal Cholesky factors
    E = np.zeros((K, m, m))      # sub-diagonal blocks L_{i,i-1}
    y = np.zeros((K, m))
    for i in range(K):
        S = R[i] @ R[i].T + np.eye(m)
        rhs = b[i].copy()
        if i:
            E[i] = np.linalg.solve(C[i - 1], -R[i].T).T   # (-R_i) C_{i-1}^{-T}
            S = S - E[i] @ E[i].T
            rhs = rhs - E[i] @ y[i - 1]
        C[i] = np.linalg.cholesky(S)
        y[i] = np.linalg.solve(C[i], rhs)
    lam = np.zeros((K, m))
    for i in reversed(range(K)):
        rhs = y[i].copy()
        if i + 1 < K:
            rhs = rhs - E[i + 1].T @ lam[i + 1]
        lam[i] = np.linalg.solve(C[i].T, rhs)
    alpha = np.zeros((K, m))
    for i in range(K):
        alpha[i] = -(R[i].T @ lam[i])
        if i:
            alpha[i] += lam[i - 1]
    return alpha


# --------------------------------------------------------------------------- #
# one segment, fds/segment.py:14-82
# --------------------------------------------------------------------------- #
def run_segment(run, u0, V, v, parameter, steps, epsilon, pool=None):
    m = len(V)
    calls = [(u0, parameter, steps)]
    calls += [(u0 + V[j] * epsilon, parameter, steps) for j in range(m)]
    calls += [(u0 + v * epsilon, parameter + epsilon, steps)]
    if pool is None:
        res = [run(*c) for c in calls]
    else:
        res = pool.starmap(run, calls)
    res = [(u, np.array(J).reshape([steps, -1])) for u, J in res]   # fds/fds.py:86
    u0p, J0 = res[0]
    Vn = np.empty_like(V)
    G = []
    for j in range(m):
        u1p, J1 = res[1 + j]
        Vn[j] = (u1p - u0p) / epsilon
        G.append(trapez_mean(J1 - J0, 0) / epsilon)
    u1p, J1 = res[m + 1]
    vn = (u1p - u0p) / epsilon
    g = trapez_mean(J1 - J0, 0) / epsilon
    return u0p, Vn, vn, J0, G, g


# --------------------------------------------------------------------------- #
# driver, fds/fds.py:29-51, 92-220
# --------------------------------------------------------------------------- #
@dataclass
class State:
    """Mirror of fds/checkpoint.py:6 (u0 V v lss G_lss g_lss J G_dil g_dil)."""
    u0: np.ndarray
    V: np.ndarray
    v: np.ndarray
    lss: LssTangent
    G_lss: list
    g_lss: list
    J: list
    G_dil: list
    g_dil: list


def lss_gradient_series(st: State):
    """Per-segment series whose mean is lss_gradient (fds/fds.py:44-51)."""
    alpha = st.lss.solve()
    grad_lss = (alpha[:, :, None] * np.array(st.G_lss)).sum(1) + np.array(st.g_lss)
    J = np.array(st.J)
    dJ = trapez_mean(J.mean(0), 0) - J[:, -1]
    dil = ((alpha * np.array(st.G_dil)).sum(1) + np.array(st.g_dil)) / J.shape[1]
    return grad_lss, dil[:, None] * dJ


def lss_gradient(st: State):
    a, b = lss_gradient_series(st)
    return windowed_mean(a) + windowed_mean(b)


def initial_state(u0, m, rng=np.random):
    """fds/fds.py:21-27 + fds/compute.py:38-42: V = Q of qr(rand(m,N)), v = 0."""
    W = rng.rand(m, u0.shape[0])
    V = np.linalg.qr(W.T)[0].T
    return State(u0, V, np.zeros(u0.shape[0]), LssTangent(), [], [], [], [], [])


def continue_shadowing(run, parameter, st: State, num_segments, steps,
                       epsilon=1e-6, run_ddt=None, pool=None):
    """fds/fds.py:92-176 (without the per-segment print / pickling)."""
    def dilation(u):
        return NoTimeDilation() if run_ddt == 0 else TimeDilation(run, u, parameter, pool)

    td = dilation(st.u0)
    V, v = td.project(st.V), td.project(st.v)
    u0, V, v, J0, G, g = run_segment(run, st.u0, V, v, parameter, steps, epsilon, pool)
    st.J.append(J0), st.G_lss.append(G), st.g_lss.append(g)
    for i in range(st.lss.K_segments() + 1, num_segments + 1):
        td = dilation(u0)
        st.G_dil.append(td.contribution(V))
        st.g_dil.append(td.contribution(v))
        V, v = td.project(V), td.project(v)
        V, v = st.lss.checkpoint(V, v)
        if i < num_segments:
            u0, V, v, J0, G, g = run_segment(run, u0, V, v, parameter, steps, epsilon, pool)
            st.J.append(J0), st.G_lss.append(G), st.g_lss.append(g)
    st.u0, st.V, st.v = u0, V, v
    return st


def shadowing(run, u0, parameter, m, num_segments, steps, runup_steps,
              epsilon=1e-6, run_ddt=None, workers=None, return_state=False):
    """fds/fds.py:178-220.  ``workers``: size of the process pool the m+2 solver
    runs of a segment are farmed to (fds/segment.py:37); None = serial."""
    pool = multiprocessing.Pool(workers) if workers else None
    try:
        u0 = np.asarray(u0, dtype=np.float64)
        if runup_steps > 0:
            u0 = run(u0, parameter, runup_steps)[0]
        st = initial_state(u0, m)
        st = continue_shadowing(run, parameter, st, num_segments, steps,
                                epsilon, run_ddt, pool)
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    if return_state:
        return st
    return np.array(st.J).mean((0, 1)), lss_gradient(st)
