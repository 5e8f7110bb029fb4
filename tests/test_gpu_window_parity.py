"""Oracle parity AT the BASELINE size (Lorenz-96 N = 2^24, m = 32): Lorenz-96 is a local stencil, so any window
of sites of the device result can be reproduced exactly by the numpy oracle on the same window padded by
8 sites per step on the left and 4 on the right (the dependency cone of one RK4 step, oracle/plugins.py
l96_rk4_step; reference tests/solvers/mock_fun3d/fun3d:48-57 has the same stencil).  The windows sit where the
production strip plan has its seams: the periodic wrap at site 0 / N, a strip seam, a strip that straddles two
CTAs of the packed (strip, column) lane layout, the last objective node, and a random interior stretch.
The objective of the last step is checked against the oracle's tree sum over the whole downloaded column."""
import numpy as np
import pytest

import fds_b200
from oracle import plugins

gpu = pytest.mark.gpu


def _periodic_rows(ctx, a, b, N):
    """Rows [a, b) of the ensemble with periodic site indices (only owned sites are read)."""
    out, i = [], a
    while i < b:
        lo = i % N
        cnt = min(b - i, N - lo)
        out.append(ctx.ensemble_rows(lo, cnt))
        i += cnt
    return np.concatenate(out, axis=0)


def _windows(plan, N, M, width=64):
    a0, ls, ns = plan['a0'], plan['strip_sites'], plan['strips']
    half = width // 2
    w = {'wrap_left': (0, width), 'wrap_right': (N - width, N)}
    r = ns // 2
    w['strip_seam'] = (a0 + r * ls - half, a0 + r * ls + half)
    lanes = plan['lanes_per_cta']
    if lanes:
        rr = next(q for q in range(ns // 3, ns) if (q * M) // lanes != (q * M + M - 1) // lanes)
        w['cta_straddling_strip_start'] = (a0 + rr * ls - half, a0 + rr * ls + half)
        w['cta_straddling_strip_middle'] = (a0 + rr * ls + ls // 2 - half, a0 + rr * ls + ls // 2 + half)
    node = 1 << plan['node_shift']
    w['last_node'] = (N - node - half, N - node + half)
    i = int(np.random.RandomState(4).randint(N // 8, N // 2))
    w['interior'] = (i, i + width + 13)            # not aligned to anything
    return {k: v for k, v in w.items() if 0 <= v[0] and v[1] <= N}


@gpu
@pytest.mark.parametrize('logn,steps', [(24, 2), (24, 9), (26, 2)])     # 2^26 x 34 = 2.3e9 elements: 64-bit indexing
def test_baseline_size_windows_bitwise_vs_oracle(logn, steps):
    N, m = 1 << logn, 32
    M = m + 2
    eps, s = 1e-6 * np.sqrt(N), 0.1
    run = fds_b200.L96(N)
    eng = run.engine(m)
    ctx = eng.ctx
    eng.set_state(np.random.RandomState(0).rand(N))
    eng.run(s, 40, want_J=False)
    eng.orthonormal_random_tangents(seed=11)
    eng.boundary(s, eps, 0, 0, 1)                       # ensemble u + eps [W; w] (fds/segment.py:43-51)
    plan = ctx.strip_plan()
    assert plan['steps_per_pass'] == 2                  # the production kernel: two RK4 steps per pass
    wins = _windows(plan, N, M)
    assert len(wins) >= 6
    gl, gr = 8 * steps, 4 * steps
    inputs = {k: _periodic_rows(ctx, a - gl, b + gr, N) for k, (a, b) in wins.items()}
    J = eng.segment(s, eps, steps)
    F = np.full(M, s + 8.0)
    F[-1] = (s + eps) + 8.0                             # fds/segment.py:64: parameter + epsilon
    for k, (a, b) in wins.items():
        got = ctx.ensemble_rows(a, b - a)               # (sites, M)
        want, _ = plugins.run_l96_ensemble(plugins.l96_rk4_step, inputs[k].T, F, steps)
        want = want[:, gl:gl + (b - a)].T
        assert np.array_equal(got, want), (k, float(np.abs(got - want).max()))
    # objective of the last step = the oracle's tree sum over the whole column (primal and the s + eps run)
    for col in (0, m + 1):
        x = np.asarray(ctx.ensemble_column(col))
        assert np.array_equal(J[-1, col], plugins.l96_objective(x)), col
    run.close()


@gpu
def test_ensemble_rows_and_column_agree_with_state_download():
    """The introspection calls themselves: rows / column against get_state and the explicit tangents."""
    N, m = 4096, 5
    eps = 1e-5
    rng = np.random.RandomState(3)
    run = fds_b200.L96(N)
    eng = run.engine(m)
    u, V, v = rng.rand(N), rng.rand(m, N), rng.rand(N)
    eng.set_state(u)
    eng.set_tangents(V, v)
    eng.ctx.boundary_finish(0, eps, 3)                  # u + eps [V; v], tangents kept
    E = eng.ctx.ensemble_rows(0, N)
    assert np.array_equal(E[:, 0], u)
    assert np.array_equal(E[:, 1:m + 1], (u[None] + V * eps).T) and np.array_equal(E[:, m + 1], u + v * eps)
    for k in (0, 3, m + 1):
        assert np.array_equal(eng.ctx.ensemble_column(k), E[:, k])
    with pytest.raises(fds_b200.FdsError, match='range'):
        eng.ctx.ensemble_rows(0, 10 ** 9)
    run.close()
