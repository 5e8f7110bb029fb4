"""The C ABI fails loudly: wrong arguments and protocol violations come back as a non-zero status
with a message (raised as FdsError by the Python side), never as silent garbage or a CPU path."""
import numpy as np
import pytest

import fds_b200
from fds_b200 import _lib
from fds_b200.device import Context

gpu = pytest.mark.gpu


@gpu
def test_create_rejects_bad_configurations():
    for kw in (dict(n_local=2, m=2), dict(n_local=64, m=0), dict(n_local=64, m=200)):
        with pytest.raises(fds_b200.FdsError):
            Context(kw['n_local'], kw['m'])
    with pytest.raises(fds_b200.FdsError, match='state dimension'):
        Context(5, 2, scheme='lorenz63')                 # Lorenz-63 has 3 components
    with pytest.raises(fds_b200.FdsError, match='halo'):
        Context(64, 2, rank=0, world=2, n_global=128, max_halo_steps=64)     # 8 * 64 sites of halo > 64 owned


@gpu
def test_protocol_violations_are_reported():
    c = Context(256, 3, max_halo_steps=4)
    c.set_state(np.random.RandomState(0).rand(256))
    c.set_tangents(np.random.RandomState(1).rand(3, 256), None)
    with pytest.raises(fds_b200.FdsError, match='halo'):
        c.segment_advance(0.0, 1e-4, 0, 2)               # no halo refresh before stepping
    with pytest.raises(fds_b200.FdsError, match='halo_steps'):
        c.halo_wrap_ensemble(99)
    c.halo_wrap_ensemble(4)
    with pytest.raises(fds_b200.FdsError, match='at most'):
        c.segment_advance(0.0, 1e-4, 0, 5)               # more steps than the halo is valid for
    with pytest.raises(fds_b200.FdsError, match='0 .adaptive. or 2'):
        c.set_qr_passes(1)
    with pytest.raises(fds_b200.FdsError, match='steps'):
        c.segment_sensitivities(10 ** 6, 1e-4)
    with pytest.raises(fds_b200.FdsError, match='primal_history'):
        c.set_primal_history(4, np.zeros(256))
    c.close()


@gpu
def test_host_solver_context_has_no_device_solver():
    c = Context(10, 2, scheme='host')
    c.set_state(np.arange(10.0))
    with pytest.raises(fds_b200.FdsError, match='host'):
        c.run(0.0, 3)
    with pytest.raises(fds_b200.FdsError, match='host'):
        c.segment_advance(0.0, 1e-4, 0, 1)
    c.close()


@gpu
def test_lss_solve_reports_indefinite_systems():
    R = np.full((3, 2, 2), np.nan)
    with pytest.raises(fds_b200.FdsError, match='positive definite'):
        fds_b200.device.lss_solve(R, np.zeros((3, 2)))


@gpu
def test_python_side_shape_checks():
    c = Context(64, 2)
    with pytest.raises(AssertionError):
        c.set_state(np.zeros(63))
    with pytest.raises(AssertionError):
        c.set_tangents(np.zeros((3, 64)))
    c.close()
    assert _lib.load().fds_last_error(None) is not None


@gpu
def test_empty_run_and_minimum_sizes():
    """Degenerate but legal inputs: zero steps (reference RunWrapper reshapes J to (0, n_qoi)), the
    smallest ring the stencil allows, a single tangent."""
    u0 = np.array([1.0, 2.0, 3.0, 4.0])
    run = fds_b200.L96(4)
    u1, J = run(u0, 0.1, 0)
    assert np.array_equal(u1, u0) and J.shape == (0, 2)
    u1, J = run(u0, 0.1, 1)
    assert J.shape == (1, 2) and np.isfinite(u1).all()
    np.random.seed(0)
    Jm, G = fds_b200.shadowing(run, u0, 0.1, 1, 2, 2, 0, epsilon=1e-5)      # m = 1, 2 segments of 2 steps
    assert Jm.shape == (2,) and G.shape == (2,) and np.isfinite(G).all()
