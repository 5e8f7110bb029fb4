"""The enqueue-only loop (device-gated adaptive CholeskyQR, results fetched at the end) against the
step-by-step loop with its host round trips: same kernels in the same order, so every history is bit-identical.
Reference loop: fds/fds.py:128-170."""
import numpy as np
import pytest

import fds_b200
from fds_b200 import _lib

gpu = pytest.mark.gpu


def _histories(cp):
    return [np.array(x) for x in (cp.lss.Rs, cp.lss.bs, cp.G_lss, cp.g_lss, cp.J, cp.G_dil, cp.g_dil, cp.u0, cp.V, cp.v)]


@gpu
@pytest.mark.parametrize('N,m,scheme', [(2048, 16, 'rk4'), (4096, 32, 'rk4'), (1024, 16, 'euler'), (777, 5, 'rk4')])
def test_deferred_loop_equals_stepwise_loop(tmp_path, N, m, scheme):
    u0 = np.random.RandomState(1).rand(N)
    out = []
    for path in (None, str(tmp_path)):            # a checkpoint path (never written: interval) forces the stepwise loop
        run = fds_b200.L96(N, scheme=scheme)
        cp = fds_b200.shadowing(run, u0, 0.05, m, 6, 20, 100, epsilon=1e-5, return_checkpoint=True, tangent_seed=4,
                                checkpoint_path=path, checkpoint_interval=10 ** 9)
        out.append(_histories(cp))
        run.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@gpu
def test_gated_two_pass_equals_host_decided_two_pass(monkeypatch):
    """CholeskyQR2 forced (fds_set_qr_passes(2)): the device-gated sequence and the host-driven one
    (FDS_GATED=0) run the same kernels; and the adaptive device decision agrees with the forced single pass
    to rounding on a well-conditioned block."""
    N, m, eps = 4096, 16, 1e-5
    rng = np.random.RandomState(2)
    u0, W = rng.rand(N), rng.rand(m, N)
    res = {}
    for gated, passes in (('1', 2), ('0', 2), ('1', 0)):
        monkeypatch.setenv('FDS_GATED', gated)
        run = fds_b200.L96(N)
        eng = run.engine(m)
        eng.ctx.set_qr_passes(passes)
        eng.set_state(u0)
        eng.run(0.0, 50, want_J=False)
        eng.orthonormal_random_tangents(block=W)
        eng.boundary(0.0, eps, 0, 0, 1)
        eng.segment(0.0, eps, 25)
        R, b, Gd, gd = eng.boundary(0.0, eps, 1, 1, 1)
        V, v = eng.get_tangents()
        res[(gated, passes)] = (R, b, Gd, V, v, eng.ctx.ensemble_rows(0, 64))
        run.close()
    for x, y in zip(res[('1', 2)], res[('0', 2)]):
        assert np.array_equal(x, y)
    for x, y in zip(res[('1', 2)][:3], res[('1', 0)][:3]):
        assert np.abs(x - y).max() <= 1e-11 * np.abs(x).max()


@gpu
def test_enqueue_api_reports_two_pass_flag_and_condition():
    """fds_boundary_enqueue's result block: R, b, G_dil, g_dil, {status, two_passes}, cond -- nearly parallel
    tangents make the device take CholeskyQR2 on its own, orthonormal ones keep the single pass."""
    N, m, eps = 2048, 16, 1e-5
    rng = np.random.RandomState(3)
    run = fds_b200.L96(N)
    eng = run.engine(m)
    ctx = eng.ctx
    eng.set_state(rng.rand(N))
    for block, want_two in ((rng.rand(m, N), False), (1.0 + 1e-5 * rng.rand(m, N), True)):
        eng.set_tangents(np.linalg.qr(block.T)[0].T if not want_two else block, np.zeros(N))
        out = _lib.pinned_empty(ctx.boundary_out_doubles)
        ctx.set_time_dilation(0)
        ctx.boundary_enqueue(0.0, eps, 0, 1, 0, 0, out)
        ctx.sync()
        R, b, Gd, gd, two, cond = ctx.boundary_unpack(out, 1)
        assert two == want_two, cond
        assert (cond > 1e3) == want_two
        Q, _ = eng.get_tangents()
        assert np.abs(Q @ Q.T - np.eye(m)).max() < 1e-12
    run.close()


@gpu
def test_rank_deficient_block_fails_at_the_end_of_the_deferred_loop():
    """A Cholesky pivot failure inside an enqueued boundary surfaces when the results are read."""
    N, m, eps = 1024, 16, 1e-5
    rng = np.random.RandomState(5)
    run = fds_b200.L96(N)
    eng = run.engine(m)
    eng.set_state(rng.rand(N))
    W = rng.rand(m, N)
    W[3] = W[2]                                   # two identical tangents: rank deficient
    eng.set_tangents(W, np.zeros(N))
    eng.ctx.set_time_dilation(0)
    eng.begin_deferred(1, 10)
    eng.boundary_deferred(0.0, eps, 0, 1, 0)
    with pytest.raises(fds_b200.FdsError, match='rank deficient'):
        eng.end_deferred()
    run.close()
