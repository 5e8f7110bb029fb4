"""Host-solver mode: the reference's plugins -- plain Python run(u0, s, steps) callables --
through fds_b200.shadowing with the algebra on the device, against the reference's own outputs.
This is the other half of the cross-check matrix of SURVEY 8b (our driver running THEIR plugin)."""
import numpy as np
import pytest

import fds_b200
from fds_b200 import checkpoint as ckpt
from oracle import plugins, shadowing_port as port
from conftest import golden

gpu = pytest.mark.gpu


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@gpu
@pytest.mark.parametrize('name,run,tol', [
    ('cfg1_lorenz63_m2_K20', plugins.run_lorenz63, 1e-3),
    ('circular_m1_K5', plugins.run_circular, 1e-6),
    ('fun3d_l96euler_N40_m16_K20', plugins.run_l96_euler, None)])
def test_reference_plugins_through_the_device_algebra(name, run, tol):
    g = golden(name)
    np.random.seed(int(g['seed']))
    cp = fds_b200.shadowing(fds_b200.HostSolver(run), g['u0'], float(g['s']), int(g['m']), int(g['K']),
                            int(g['steps']), int(g['runup']), epsilon=float(g['eps']), return_checkpoint=True)
    J = np.array(cp.J)
    assert np.array_equal(J[0], g['J'][0])                 # the solver is theirs: identical primal
    G = fds_b200.lss_gradient(cp)
    if tol is not None:
        assert maxrel(J.mean((0, 1)), g['J_ref']) < 1e-6
        assert abs(G[0] - g['G_ref'][0]) < tol * abs(g['G_ref'][0])
    else:                                                  # chaotic, 20 segments: the reference's windows
        Jm = J.mean((0, 1))
        assert 1 < Jm[0] < 4 and 10 < Jm[1] < 40 and -0.1 < G[0] < 0.5 and 1 < G[1] < 8
        assert (np.abs(G - g['G_ref']) <= 3 * g['G_err']).all()


@gpu
def test_host_solver_equals_device_plugin():
    """Same system advanced by the numpy plugin on the host and by the CUDA plugin: the two engines
    differ only in who integrates -- and both integrators are bit-identical."""
    N, m, K, steps, eps, s = 64, 5, 3, 20, 1e-4, 0.1
    u0 = plugins.run_l96_rk4(np.random.RandomState(0).rand(N), s, 300)[0]
    np.random.seed(3)
    a = fds_b200.shadowing(fds_b200.HostSolver(plugins.run_l96_rk4), u0, s, m, K, steps, 10, epsilon=eps,
                           return_checkpoint=True)
    np.random.seed(3)
    b = fds_b200.shadowing(fds_b200.L96(N), u0, s, m, K, steps, 10, epsilon=eps, return_checkpoint=True)
    assert np.array_equal(np.array(a.J), np.array(b.J))
    assert np.array_equal(a.u0, b.u0) and maxrel(a.V, b.V) < 1e-9
    assert maxrel(np.array(a.lss.Rs), np.array(b.lss.Rs)) < 1e-9
    assert maxrel(fds_b200.lss_gradient(a), fds_b200.lss_gradient(b)) < 1e-7


@gpu
def test_host_solver_restart_and_plugin_signatures(tmp_path):
    """reference tests/test_restart.py:42-64 with a 4-argument plugin (run_id) and a thread pool."""
    seen = []

    def run(u0, s, steps, run_id=None):
        seen.append(run_id)
        return plugins.run_circular(u0, s, steps)

    hs = fds_b200.HostSolver(run, simultaneous_runs=3)
    u0 = np.array([1.0, 0.5])
    np.random.seed(0)
    fds_b200.shadowing(hs, u0, 1.0, 1, 10, 100, 0, checkpoint_path=str(tmp_path))
    cp = ckpt.load_last_checkpoint(str(tmp_path), 1)
    assert cp.lss.K_segments() == 10
    J2, G2 = fds_b200.continue_shadowing(hs, 1.0, cp, 20, 100)
    np.random.seed(0)
    J1, G1 = fds_b200.shadowing(fds_b200.HostSolver(run), u0, 1.0, 1, 20, 100, 0)
    assert np.array_equal(J1, J2) and np.allclose(G1, G2)
    assert 'segment00_baseline' in seen and 'segment10_param_perturb001' in seen
    assert any(r and r.startswith('time_dilation_') and r.endswith('_3steps') for r in seen)


@gpu
def test_bare_callable_is_still_rejected():
    with pytest.raises(TypeError, match='HostSolver'):
        fds_b200.shadowing(plugins.run_l96_rk4, np.zeros(40), 0.0, 2, 2, 10, 0)


def _lorenz_vdp_run(u, s, nsteps, dt=0.05):
    """A user-style pure-Python plugin: the Lorenz-63 + Van der Pol system of the reference's
    tests/test_lyapunov_vector.py:13-30, integrated with scipy's odeint, objective [z, x0^2]."""
    from scipy.integrate import odeint

    def dudt(w, t):
        x, y, z, x0, y0 = w
        return [10 * (y - x), x * (s - z) - y, x * y - 8. / 3 * z, y0, (z - x0 ** 2) * y0 - x0 * (2 * np.pi) ** 2]
    J = np.empty([nsteps, 2])
    for i in range(nsteps):
        u = odeint(dudt, u, [0, dt])[1]
        J[i] = u[2], u[3] ** 2
    return u, J


@gpu
def test_lyapunov_spectrum_of_a_python_plugin():
    """The reference's tests/test_lyapunov_vector.py:33-48 through HostSolver: time dilation off (run_ddt=0), five
    exponents from log|diag R| on the device, asserted against the reference's literal values within 3 sigma (its
    own criterion) AND against the spectrum the reference itself computes (fixture, oracle/gen_golden.py)."""
    g = golden('lyapunov_lorenz_vdp_from_reference')
    m, steps, K, dt = int(g['m']), int(g['steps']), int(g['K']), float(g['dt'])
    np.random.seed(11)
    cp = fds_b200.shadowing(fds_b200.HostSolver(_lorenz_vdp_run), np.ones(5), float(g['s']), m, K, steps,
                            int(g['runup']), run_ddt=0, return_checkpoint=True)
    L = cp.lss.lyapunov_exponents() / (steps * dt)
    L, L_err = fds_b200.timeseries.mean_std(L)
    for lit, lo, hi in zip((0.90, 0.25, 0.0, -4.98, -6.40), L - 3 * L_err, L + 3 * L_err):
        assert lo < lit < hi
    assert (np.abs(L - g['L']) < 3 * (L_err + g['L_err'])).all()
    assert (np.abs(L - g['L']) < 0.05 * np.abs(g['L']).max()).all()      # same trajectory, same seed: far closer than the bars
    v = cp.lss.lyapunov_covariant_vectors()
    assert v.shape == (m, K + 1, m) == tuple(g['clv_shape'])
    assert all(len(x) == K for x in (cp.G_dil, cp.g_dil)) and np.abs(np.array(cp.G_dil)).max() == 0.0


@gpu
def test_file_mode_of_the_reference_mpi_path(tmp_path):
    """The reference's tests/test_fun3d_mpi.py:82-154 with its mock solver (forward-Euler Lorenz-96): the initial
    state is the NAME of an HDF5 file, ``solve`` reads ``/field`` of the file it is given and answers with the
    name of the ``output.h5`` it wrote into ``get_host_dir(run_id)``; the perturbed initial conditions arrive as
    ``<get_host_dir(run_id)>/input.h5`` (fds/segment.py:46,51).  Same numbers as the in-memory mode."""
    import os
    from fds_b200 import h5field
    base = str(tmp_path)
    inputs = {}

    def get_host_dir(run_id):
        return os.path.join(base, run_id)

    def solve(u0, s, nsteps, run_id, interprocess):
        assert isinstance(u0, str)
        inputs[run_id] = u0
        u1, J = plugins.run_l96_euler(h5field.read_field(u0), s, nsteps)
        os.makedirs(get_host_dir(run_id), exist_ok=True)
        path = os.path.join(get_host_dir(run_id), 'output.h5')
        h5field.write_field(path, u1)
        return path, J

    N, m, K, steps, eps, s = 40, 4, 3, 25, 1e-4, 0.1
    u0 = plugins.run_l96_euler(np.random.RandomState(5).rand(N), s, 200)[0]
    u0_file = os.path.join(base, 'u0.h5')
    h5field.write_field(u0_file, u0)

    np.random.seed(11)
    a = fds_b200.shadowing(fds_b200.HostSolver(solve, simultaneous_runs=2), u0_file, s, m, K, steps, 10,
                           epsilon=eps, checkpoint_path=base, get_host_dir=get_host_dir, return_checkpoint=True)
    np.random.seed(11)
    b = fds_b200.shadowing(fds_b200.HostSolver(plugins.run_l96_euler), u0, s, m, K, steps, 10, epsilon=eps,
                           return_checkpoint=True)
    assert np.array_equal(np.array(a.J), np.array(b.J)) and np.array_equal(a.u0, b.u0)
    assert np.array_equal(a.V, b.V) and np.array_equal(np.array(a.lss.Rs), np.array(b.lss.Rs))
    assert np.array_equal(fds_b200.lss_gradient(a), fds_b200.lss_gradient(b))
    # who got which file: the run-up the user's file, the baseline the previous baseline's output, every
    # perturbed trajectory its own input.h5
    assert inputs['runup'] == u0_file
    assert inputs['segment00_baseline'] == os.path.join(base, 'runup', 'output.h5')
    assert inputs['segment01_baseline'] == os.path.join(base, 'segment00_baseline', 'output.h5')
    assert inputs['time_dilation_01_2steps'] == os.path.join(base, 'segment00_baseline', 'output.h5')
    for rid in ('segment00_init_perturb003', 'segment02_param_perturb004'):
        assert inputs[rid] == os.path.join(base, rid, 'input.h5')
    w = h5field.read_field(inputs['segment00_init_perturb000'])
    assert w.shape == (N,) and w.dtype == np.float64

    # continue from the checkpoint on disk (arrays): the solver still gets files
    cp = ckpt.load_last_checkpoint(base, m)
    inputs.clear()
    c = fds_b200.continue_shadowing(fds_b200.HostSolver(solve, field_files=True), s, cp, K + 2, steps,
                                    epsilon=eps, get_host_dir=get_host_dir, return_checkpoint=True)
    d = fds_b200.continue_shadowing(fds_b200.HostSolver(plugins.run_l96_euler), s, b, K + 2, steps, epsilon=eps,
                                    return_checkpoint=True)
    assert np.array_equal(np.array(c.J), np.array(d.J)) and np.array_equal(c.V, d.V)
    assert inputs['segment03_baseline'] == os.path.join(base, 'segment03_baseline', 'input.h5')
    assert inputs['segment04_baseline'] == os.path.join(base, 'segment03_baseline', 'output.h5')
