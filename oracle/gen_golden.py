"""Generate tests/golden/*.npz from the REAL reference (/root/reference) and pin
oracle/shadowing_port.py against it.  Runs in the build container only (the
GPU box has no /root/reference); the fixtures it writes are committed.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

Every fixture stores its inputs next to the reference's outputs, so the tests
never need the reference again.
"""
import os
import sys
import io
import contextlib

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

import fds as ref                      # noqa: E402  the reference itself
import pascal_lite as pascal           # noqa: E402
from fds.timedilation import TimeDilation as RefTimeDilation   # noqa: E402
from fds.lsstan import LssTangent as RefLss                    # noqa: E402
from fds.segment import run_segment as ref_run_segment        # noqa: E402
from fds.fds import RunWrapper                                 # noqa: E402
from multiprocessing import Manager                            # noqa: E402

from oracle import plugins, shadowing_port as port            # noqa: E402


def quiet(fn, *a, **k):
    """The reference prints lss_gradient every segment (fds/fds.py:161)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ref_checkpoint_arrays(cp):
    return dict(Rs=np.array(cp.lss.Rs), bs=np.array(cp.lss.bs),
                G_lss=np.array(cp.G_lss), g_lss=np.array(cp.g_lss),
                J=np.array(cp.J), G_dil=np.array(cp.G_dil),
                g_dil=np.array(cp.g_dil))


def port_state_arrays(st):
    return dict(Rs=np.array(st.lss.Rs), bs=np.array(st.lss.bs),
                G_lss=np.array(st.G_lss), g_lss=np.array(st.g_lss),
                J=np.array(st.J), G_dil=np.array(st.G_dil),
                g_dil=np.array(st.g_dil))


def maxrel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# --------------------------------------------------------------------------- #
def full_run(name, run, u0, s, m, K, steps, runup, eps, seed, pin_tol, workers=None):
    """Reference shadowing() vs the port on the same seed; save both."""
    np.random.seed(seed)
    cp = quiet(ref.shadowing, run, u0, s, m, K, steps, runup, epsilon=eps,
               return_checkpoint=True, simultaneous_runs=workers)
    G_ref = ref.lss_gradient(cp)
    A = ref_checkpoint_arrays(cp)
    J_ref = A['J'].mean((0, 1))
    alpha_ref = cp.lss.solve()

    np.random.seed(seed)
    st = port.shadowing(run, u0, s, m, K, steps, runup, epsilon=eps, return_state=True)
    P = port_state_arrays(st)
    G_port = port.lss_gradient(st)
    errs = {k: maxrel(P[k], A[k]) for k in A}
    errs['G'] = maxrel(G_port, G_ref)
    print(name, 'port-vs-reference max rel err:', {k: '%.1e' % v for k, v in errs.items()})
    # first-segment quantities must agree to rounding; later ones only as far
    # as chaos allows, so pin on the first two segments + the final gradient
    for k in ('J', 'G_lss', 'g_lss', 'Rs', 'bs', 'G_dil', 'g_dil'):
        e = maxrel(P[k][:2], A[k][:2])
        assert e < pin_tol, (name, k, e)

    # series + error bar as the survey defines it (mean_std of the per-segment series)
    class _S:  # adapt reference checkpoint to the port's series helper
        pass
    stt = port.State(None, None, None, port.LssTangent(list(A['Rs']), list(A['bs'])),
                     list(A['G_lss']), list(A['g_lss']), list(A['J']),
                     list(A['G_dil']), list(A['g_dil']))
    a, b = port.lss_gradient_series(stt)
    G_mean, G_err = port.mean_std(a + b)
    assert maxrel(G_mean, G_ref) < 1e-9
    if A['J'].size > 200000:      # keep fixtures small: drop the raw J history
        A = dict(A, J_last=A['J'][:, -1], J_tmean=A['J'].mean(0)[::100])
        del A['J']
    np.savez_compressed(
        os.path.join(GOLD, name + '.npz'),
        u0=u0, s=s, m=m, K=K, steps=steps, runup=runup, eps=eps, seed=seed,
        J_ref=J_ref, G_ref=G_ref, G_err=G_err, alpha_ref=alpha_ref,
        u_final=np.asarray(cp.u0.field if hasattr(cp.u0, 'field') else cp.u0),
        **A)
    print(name, 'J', J_ref, 'G', G_ref, '+-', G_err)


# --------------------------------------------------------------------------- #
def single_segment(name, run, N, m, steps, eps, s, seed, with_boundary=True):
    """Reference run_segment / TimeDilation / LssTangent.checkpoint called
    directly on numpy-backed symbolic arrays (one segment, one boundary)."""
    rng = np.random.RandomState(seed)
    u0 = run(rng.rand(N) * 2 + 1, s, 300)[0]               # on the attractor
    V = np.linalg.qr(rng.rand(m, N).T)[0].T
    v = rng.randn(N) / np.sqrt(N)
    wrapped = RunWrapper(run)
    mgr = Manager()
    ip = (mgr.Lock(), mgr.dict())

    sa = lambda x: pascal.symbolic_array(np.shape(x)[:-1], field=np.array(x))
    np.random.seed(seed)
    u1, V1, v1, J0, G, g = ref_run_segment(wrapped, sa(u0), sa(V), sa(v), s, 0, steps,
                                           eps, None, ip)
    ref.compute.run_compute([V1, v1])
    out = dict(u0=u0, V=V, v=v, s=s, eps=eps, steps=steps,
               u1=u1.field, V1=V1.field, v1=v1.field, J0=J0,
               G=np.array(G), g=np.array(g))

    # the port on the same inputs (pin: identical operations => tiny error)
    pu1, pV1, pv1, pJ0, pG, pg = port.run_segment(run, u0, V, v, s, steps, eps)
    assert np.array_equal(pu1, out['u1']) and np.array_equal(pV1, out['V1'])
    assert np.array_equal(pv1, out['v1']) and np.array_equal(pJ0, out['J0'])
    assert np.array_equal(np.array(pG), out['G']) and np.array_equal(pg, out['g'])

    if with_boundary:
        # boundary at the END of that segment: time dilation at u1, then QR
        td = RefTimeDilation(wrapped, sa(out['u1']), s, 'td', None, ip)
        Vs, vs = sa(out['V1']), sa(out['v1'])
        Gd, gd = td.contribution(Vs), td.contribution(vs)
        Vp, vp = td.project(Vs), td.project(vs)
        lss = RefLss()
        Q, vq = lss.checkpoint(Vp, vp)
        ref.compute.run_compute([Gd, gd, Vp, vp, Q, vq, lss.Rs[0], lss.bs[0], td.dxdt])
        out.update(dxdt=td.dxdt.field, G_dil=Gd.field, g_dil=gd.field,
                   Vp=Vp.field, vp=vp.field, Q=Q.field, vq=vq.field,
                   R=lss.Rs[0].field, b=lss.bs[0].field)
        ptd = port.TimeDilation(run, out['u1'], s)
        plss = port.LssTangent()
        assert np.array_equal(ptd.dxdt, out['dxdt'])
        for got, key in ((ptd.contribution(out['V1']), 'G_dil'),
                         (ptd.contribution(out['v1']), 'g_dil'),
                         (ptd.project(out['V1']), 'Vp'), (ptd.project(out['v1']), 'vp')):
            assert maxrel(got, out[key]) < 1e-14, (key, maxrel(got, out[key]))
        pQ, pvq = plss.checkpoint(ptd.project(out['V1']), ptd.project(out['v1']))
        for got, key in ((pQ, 'Q'), (pvq, 'vq'), (plss.Rs[0], 'R'), (plss.bs[0], 'b')):
            assert maxrel(got, out[key]) < 1e-13, (key, maxrel(got, out[key]))
    mgr.shutdown()
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print(name, 'ok; N=%d m=%d steps=%d' % (N, m, steps))


# --------------------------------------------------------------------------- #
def kkt_known_answer():
    """LssTangent.solve on a random well-conditioned problem (the survey's
    known-answer test) + the dense block-Thomas cross-check."""
    rng = np.random.RandomState(7)
    K, m = 50, 16
    R = rng.randn(K, m, m) / np.sqrt(m) + np.eye(m)
    R = np.triu(R)
    b = rng.randn(K, m)
    lss = RefLss()
    lss.Rs, lss.bs = list(R), list(b)
    alpha = lss.solve()
    bt = port.lss_solve_block_thomas(R, b)
    e = maxrel(bt, alpha)
    print('kkt: block-Thomas vs reference solve: %.2e' % e)
    assert e < 1e-11
    assert maxrel(port.lss_solve(R, b), alpha) == 0.0
    np.savez_compressed(os.path.join(GOLD, 'kkt_K50_m16.npz'), R=R, b=b, alpha=alpha)


def clv_fixture():
    """Covariant Lyapunov vectors (fds/lsstan.py:66-85; consumed by apps/**/drawCLV.py, shape-checked
    upstream in tests/test_lyapunov_vector.py:47-48): the reference's LssTangent fed with the R history of
    two reference runs already stored as fixtures -> vectors, magnitudes and sines of the angles."""
    out = {}
    for tag, name in (('cfg1', 'cfg1_lorenz63_m2_K20'), ('cfg3', 'cfg3_l96rk4_N40_m16_K100')):
        g = np.load(os.path.join(GOLD, name + '.npz'))
        lss = RefLss()
        lss.Rs, lss.bs = list(g['Rs']), list(g['bs'])
        v = lss.lyapunov_covariant_vectors()
        mag, sin = lss.lyapunov_covariant_magnitude_and_sin_angle()
        assert v.shape == (g['Rs'].shape[1], g['Rs'].shape[0] + 1, g['Rs'].shape[1])
        out[tag + '_src'] = name
        out[tag + '_v'], out[tag + '_mag'], out[tag + '_sin'] = v, mag, sin
        out[tag + '_exponents'] = lss.lyapunov_exponents()
        out[tag + '_exponents_range'] = lss.lyapunov_exponents((2, 11))
    np.savez_compressed(os.path.join(GOLD, 'clv_from_reference.npz'), **out)
    print('clv fixture:', {k: np.shape(v) for k, v in out.items()})


def reference_checkpoint_fixture():
    """A checkpoint FILE written by the reference (dill pickle of lazy pascal_lite graphs, fds/checkpoint.py:8-18),
    what the reference itself reads out of it (V and v evaluated by its own executor), and the reference's
    continuation from it (restart protocol of tests/test_restart.py:52-64): Lorenz-96 Euler N=40, m=4, 3 segments
    written, continued to 6.  The file is committed as it is (15 KB): fds_b200/refcheckpoint.py must read it
    without the reference on the path."""
    import shutil
    import tempfile
    from fds.checkpoint import load_last_checkpoint
    from fds.compute import run_compute
    tmp = tempfile.mkdtemp()
    s, m, K0, K1, steps, eps = 0.1, 4, 3, 6, 20, 1e-4
    np.random.seed(3)
    u0 = np.random.rand(40)
    quiet(ref.shadowing, plugins.run_l96_euler, u0, s, m, K0, steps, 50, epsilon=eps, checkpoint_path=tmp)
    name = 'm{0}_segment{1}'.format(m, K0)
    shutil.copy(os.path.join(tmp, name), os.path.join(GOLD, 'refcp_l96euler_' + name))
    cp = load_last_checkpoint(tmp, m)
    run_compute([cp.V, cp.v])                          # what the reference's executor makes of the stored graphs
    out = dict(u0=cp.u0.field, V=cp.V.field, v=cp.v.field, s=s, m=m, K0=K0, K1=K1, steps=steps, eps=eps)
    out.update({'file_' + k: v for k, v in ref_checkpoint_arrays(cp).items()})
    cp = load_last_checkpoint(tmp, m)
    cp2 = quiet(ref.continue_shadowing, plugins.run_l96_euler, s, cp, K1, steps, epsilon=eps, return_checkpoint=True)
    out.update({'cont_' + k: v for k, v in ref_checkpoint_arrays(cp2).items()})
    out['cont_G'] = ref.lss_gradient(cp2)
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(GOLD, 'refcp_l96euler_expected.npz'), **out)
    print('reference checkpoint fixture: file', name, 'continued to', cp2.lss.K_segments(), 'segments, G =', out['cont_G'])


def reference_graph_ops_fixture():
    """A reference-format checkpoint whose lazy V and v go through every operator family of pascal_lite
    (neg, transpose, sum over an axis, pow, div, outer, getitem, setitem, reshape / ravel, norm, dot, qr_transpose;
    pascal_lite/operators/*.py), dill-dumped as the reference does, next to what the reference's executor computes:
    pins the by-name evaluation of fds_b200/refcheckpoint.py operator by operator."""
    import dill
    from fds.checkpoint import Checkpoint as RefCheckpoint
    from fds.compute import run_compute
    rng = np.random.RandomState(21)
    N, m = 6, 3
    A = pascal.symbolic_array(field=rng.rand(m, N))
    A.value.shape = (m,)
    B = pascal.symbolic_array(field=rng.rand(2, m, N))
    B.value.shape = (2, m)
    x = pascal.symbolic_array(field=rng.rand(N))
    Bt = B.T                                        # transpose -> (m, 2)
    s1 = Bt.sum(1)                                  # sum over an axis -> (m,)
    d = pascal.dot(A, x)                            # ReduceSum -> (m,), not distributed
    V1 = A * 2.0 + s1 / (pascal.norm(x) + 1.0) - pascal.outer(d, x)      # broadcast, div, pow 0.5, reshape inside outer
    V2 = V1.copy()
    V2[1] = x ** 2.0                                # setitem, pow
    V3 = V2 + (-Bt).ravel()[1:4] * 0.25             # neg, reshape (ravel), getitem with a slice
    Q, R = pascal.qr_transpose(V3)
    b = pascal.dot(Q, x)
    vf = x - pascal.dot(b, Q)                       # (b * Q).sum(): sum over all axes of a distributed product
    run_compute([R, b])                             # what continue_shadowing evaluates at its last boundary
    lss = RefLss()
    lss.Rs.append(R.field)
    lss.bs.append(b.field)
    cp = RefCheckpoint(x, Q, vf, lss, [[rng.rand(2) for _ in range(m)]], [rng.rand(2)], [rng.rand(5, 2)],
                       [rng.rand(m)], [float(rng.rand())])
    with open(os.path.join(GOLD, 'refcp_ops_m3_segment1'), 'wb') as f:
        dill.dump(cp, f)
    run_compute([Q, vf])
    np.savez_compressed(os.path.join(GOLD, 'refcp_ops_expected.npz'), u0=x.field, V=Q.field, v=vf.field,
                        R=lss.Rs[0], b=lss.bs[0])
    print('reference graph-ops fixture: V', Q.field.shape, 'v', vf.field.shape)


def lorenz_vdp_dudt(u, s):
    """The reference's only pure-Python plugin (tests/test_lyapunov_vector.py:13-22): a Lorenz-63 system driving
    a Van der Pol-type oscillator."""
    x, y, z, x0, y0 = u
    return np.array([10 * (y - x), x * (s - z) - y, x * y - 8. / 3 * z, y0,
                     (z - x0 ** 2) * y0 - x0 * (2 * np.pi) ** 2])


def lorenz_vdp_run(u, s, nsteps, dt=0.05):
    """run(u, s, steps) of that system with scipy's odeint, objective [z, x0^2] (tests/test_lyapunov_vector.py:24-30)."""
    from scipy.integrate import odeint
    J = np.empty([nsteps, 2])
    for i in range(nsteps):
        u = odeint(lambda w, t: lorenz_vdp_dudt(w, s), u, [0, dt])[1]
        J[i] = u[2], u[3] ** 2
    return u, J


def lyapunov_plugin_fixture():
    """tests/test_lyapunov_vector.py:33-48 run by the reference itself: the five Lyapunov exponents (mean and
    standard error over the segments, fds/timeseries.py:3-7) of the Lorenz + Van der Pol system, time dilation off
    (run_ddt=0)."""
    m, steps, K, dt = 5, 50, 50, 0.05
    np.random.seed(11)
    cp = quiet(ref.shadowing, lorenz_vdp_run, np.ones(5), 28, m, K, steps, 100, run_ddt=0, return_checkpoint=True)
    L = cp.lss.lyapunov_exponents() / (steps * dt)
    Lm, Le = ref.timeseries.mean_std(L)
    np.savez_compressed(os.path.join(GOLD, 'lyapunov_lorenz_vdp_from_reference.npz'), L=Lm, L_err=Le, m=m, steps=steps,
                        K=K, dt=dt, s=28.0, runup=100, clv_shape=np.array(cp.lss.lyapunov_covariant_vectors().shape))
    print('lyapunov plugin fixture: L =', Lm, '+-', Le)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ['refops']:
        reference_graph_ops_fixture()
        return
    if sys.argv[1:] == ['lyap']:
        lyapunov_plugin_fixture()
        return
    if sys.argv[1:] == ['clv']:
        clv_fixture()
        return
    if sys.argv[1:] == ['refcp']:
        reference_checkpoint_fixture()
        return
    if sys.argv[1:] == ['vdp']:
        full_run('cfg2_vanderpol_m1_K20', plugins.run_vanderpol, np.array([1.0, 1.0]), 4.0,
                 1, 20, 50000, 2000, 1e-6, 0, pin_tol=1e-6)
        return
    kkt_known_answer()
    single_segment('seg_l96rk4_N40_m16', plugins.run_l96_rk4, 40, 16, 100, 1e-4, 0.0, 1)
    single_segment('seg_l96rk4_N37_m3', plugins.run_l96_rk4, 37, 3, 7, 1e-4, 0.1, 2)
    single_segment('seg_l96rk4_N1024_m8', plugins.run_l96_rk4, 1024, 8, 20, 1e-5, 0.0, 3)
    single_segment('seg_l96rk4_N2048_m32', plugins.run_l96_rk4, 2048, 32, 10, 4.5e-5, 0.0, 4)
    single_segment('seg_l96euler_N40_m16', plugins.run_l96_euler, 40, 16, 100, 1e-4, 0.1, 5)

    u40 = np.random.RandomState(0).rand(40)
    # BASELINE config 3: Lorenz 96 N=40 F=8 (s=0), m=16, 100 segments, RK4
    full_run('cfg3_l96rk4_N40_m16_K100', plugins.run_l96_rk4, u40, 0.0, 16, 100, 100,
             2000, 1e-4, 0, pin_tol=1e-7, workers=None)
    # the reference's own L96 test (tests/test_fun3d.py:18-25): Euler, s=0.1, K=20
    full_run('fun3d_l96euler_N40_m16_K20', plugins.run_l96_euler, u40, 0.1, 16, 20, 100,
             2000, 1e-4, 0, pin_tol=1e-7)
    # BASELINE config 1: Lorenz 63, m=2, 20 segments (tests/test_lorenz.py:48-67)
    u63 = np.array([1.0, 1.0, 28.0])
    full_run('cfg1_lorenz63_m2_K20', plugins.run_lorenz63, u63, 28.0, 2, 20, 1000,
             5000, 1e-6, 0, pin_tol=1e-6)
    # tightest analytic check of the reference suite (tests/test_circular.py:33-42)
    full_run('circular_m1_K5', plugins.run_circular, np.array([1.0, 0.5]), 1.0, 1, 5,
             5000, 10000, 1e-6, 0, pin_tol=1e-6)
    # BASELINE config 2: Van der Pol, m=1 (tests/test_vanderpol.py:33-36)
    full_run('cfg2_vanderpol_m1_K20', plugins.run_vanderpol, np.array([1.0, 1.0]), 4.0, 1,
             20, 50000, 2000, 1e-6, 0, pin_tol=1e-6)
    clv_fixture()
    reference_checkpoint_fixture()
    reference_graph_ops_fixture()
    lyapunov_plugin_fixture()


if __name__ == '__main__':
    main()
