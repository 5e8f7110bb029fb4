"""Checkpoints written by the REFERENCE (dill pickles of lazy pascal_lite graphs, fds/checkpoint.py:8-18) are
read without the reference, dill or pascal_lite importable: tests/golden/refcp_l96euler_m4_segment3 is such a
file, committed as the reference wrote it (oracle/gen_golden.py reference_checkpoint_fixture), next to what
the reference's own executor makes of it and the reference's continuation from it."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

FILE = os.path.join(ROOT, 'tests', 'golden', 'refcp_l96euler_m4_segment3')


def test_reference_checkpoint_file_reads_bit_for_bit_without_the_reference():
    # a clean interpreter: nothing of the reference (or dill) may be needed, and nothing of it gets imported
    code = ("import sys, numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "sys.modules['dill'] = None; sys.modules['fds'] = None; sys.modules['pascal_lite'] = None\n"
            "from fds_b200.checkpoint import load_checkpoint\n"
            "cp = load_checkpoint(%r)\n"
            "np.savez(sys.argv[1], u0=cp.u0, V=cp.V, v=cp.v, Rs=np.array(cp.lss.Rs), bs=np.array(cp.lss.bs),\n"
            "         G_lss=np.array(cp.G_lss), g_lss=np.array(cp.g_lss), J=np.array(cp.J),\n"
            "         G_dil=np.array(cp.G_dil), g_dil=np.array(cp.g_dil))\n" % (ROOT, FILE))
    out = os.path.join(ROOT, 'tests', '_refcp_out.npz')
    try:
        subprocess.run([sys.executable, '-c', code, out], check=True, capture_output=True, text=True)
        got, want = np.load(out), golden('refcp_l96euler_expected')
        for k in ('u0', 'V', 'v'):
            assert np.array_equal(got[k], want[k]), k                 # V, v: the stored lazy graphs, evaluated
        for k in ('Rs', 'bs', 'G_lss', 'g_lss', 'J', 'G_dil', 'g_dil'):
            assert np.array_equal(got[k], want['file_' + k]), k
    finally:
        if os.path.exists(out):
            os.remove(out)


def test_load_last_checkpoint_picks_up_a_reference_file(tmp_path):
    from fds_b200.checkpoint import load_last_checkpoint, verify_checkpoint
    shutil.copy(FILE, tmp_path / 'm4_segment3')
    (tmp_path / 'm4_segment_garbage').write_text('x')                  # tests/test_restart.py:46-50
    cp = load_last_checkpoint(str(tmp_path), 4)
    assert verify_checkpoint(cp) and cp.lss.K_segments() == 3 and cp.lss.m_modes() == 4
    assert cp.V.shape == (4, 40) and cp.u0.shape == (40,) and isinstance(cp.V, np.ndarray)
    assert load_last_checkpoint(str(tmp_path), 5) is None


def test_hostile_pickle_is_refused(tmp_path):
    import pickle
    from fds_b200.refcheckpoint import load_reference_checkpoint, ReferenceCheckpointError
    p = tmp_path / 'evil'
    p.write_bytes(pickle.dumps(os.system))                               # a global outside the whitelist
    with pytest.raises((ReferenceCheckpointError, pickle.UnpicklingError)):
        load_reference_checkpoint(str(p))


@pytest.mark.gpu
def test_continue_shadowing_from_a_reference_checkpoint():
    """The reference's restart protocol (tests/test_restart.py:52-64) across implementations: 3 segments
    run and written by the reference, 3 more run here on the GPU, against the reference's own continuation.
    Lorenz-96 forward Euler is the reference's own solver (bit-identical plugin), so the primal objective
    history of the new segments is identical; R, b and dJ/ds agree up to the QR sign convention / 1e-9."""
    import fds_b200
    from fds_b200.checkpoint import load_checkpoint
    want = golden('refcp_l96euler_expected')
    cp = load_checkpoint(FILE)
    run = fds_b200.L96(40, scheme='euler')
    cp2 = fds_b200.continue_shadowing(run, float(want['s']), cp, int(want['K1']), int(want['steps']),
                                      epsilon=float(want['eps']), return_checkpoint=True)
    run.close()
    assert cp2.lss.K_segments() == int(want['K1'])
    J = np.array(cp2.J)
    assert np.array_equal(J, want['cont_J'])                            # all six segments, bit for bit
    Rd, Rr = np.array(cp2.lss.Rs), want['cont_Rs']
    assert np.array_equal(Rd[:3], Rr[:3])                               # the reference's own, carried over
    # first new boundary: same V (from the file), so R agrees up to LAPACK's row signs (CholeskyQR: diag R > 0);
    # a flipped row is the finite difference in the opposite direction from then on: O(eps) differences
    eps = float(want['eps'])
    e3 = np.abs(np.abs(Rd[3]) - np.abs(Rr[3])).max() / np.abs(Rr[3]).max()
    e45 = np.abs(np.abs(Rd[4:]) - np.abs(Rr[4:])).max() / np.abs(Rr[4:]).max()
    G = fds_b200.lss_gradient(cp2)
    eG = np.abs(G - want['cont_G']).max() / np.abs(want['cont_G']).max()
    print('R3 %.2e  R4,5 %.2e  dJ/ds %.2e' % (e3, e45, eG))
    assert e3 < 1e-12 and e45 < eps and eG < 0.1 * eps       # measured: 6e-17, 1e-5, 1e-7


def _fixture_checkpoint():
    from fds_b200.checkpoint import Checkpoint
    from fds_b200.lsstan import LssTangent
    g = golden('refcp_l96euler_expected')
    return g, Checkpoint(g['u0'], g['V'], g['v'], LssTangent(list(g['file_Rs']), list(g['file_bs'])),
                         list(g['file_G_lss']), list(g['file_g_lss']), list(g['file_J']), list(g['file_G_dil']),
                         list(g['file_g_dil']))


def test_reference_format_round_trip(tmp_path):
    """save_reference_checkpoint writes the reference's classes by name; our own reader takes the file back."""
    from fds_b200.refcheckpoint import save_reference_checkpoint, is_reference_checkpoint
    from fds_b200.checkpoint import load_last_checkpoint, verify_checkpoint
    g, cp = _fixture_checkpoint()
    save_reference_checkpoint(str(tmp_path), cp)
    f = tmp_path / 'm4_segment3'
    assert f.exists() and is_reference_checkpoint(str(f))
    assert 'pascal_lite' not in sys.modules and 'fds' not in sys.modules
    back = load_last_checkpoint(str(tmp_path), 4)
    assert verify_checkpoint(back)
    assert np.array_equal(back.u0, cp.u0) and np.array_equal(back.V, cp.V) and np.array_equal(back.v, cp.v)
    assert np.array_equal(np.array(back.lss.Rs), g['file_Rs']) and np.array_equal(np.array(back.J), g['file_J'])
    assert np.array_equal(np.array(back.G_lss), g['file_G_lss']) and back.g_dil == list(g['file_g_dil'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/fds'), reason='needs the reference (build container only)')
def test_the_reference_continues_from_a_file_written_here(tmp_path):
    """The other direction of the restart protocol: a checkpoint written by save_reference_checkpoint is loaded by
    the REFERENCE (dill) and continued by the reference's continue_shadowing -- and gives what the reference gets
    from its own file (fixture), bit for bit: same evaluated V and v, same histories."""
    from fds_b200.refcheckpoint import save_reference_checkpoint
    g, cp = _fixture_checkpoint()
    save_reference_checkpoint(str(tmp_path), cp)
    code = ("import sys, io, contextlib, numpy as np\n"
            "sys.dont_write_bytecode = True\n"
            "sys.path.insert(0, '/root/reference'); sys.path.insert(0, %r)\n"
            "import fds\n"
            "from fds.checkpoint import load_last_checkpoint\n"
            "from oracle import plugins\n"
            "cp = load_last_checkpoint(%r, 4)\n"
            "assert type(cp).__module__ == 'fds.checkpoint' and type(cp.lss).__module__ == 'fds.lsstan'\n"
            "with contextlib.redirect_stdout(io.StringIO()):\n"
            "    cp2 = fds.continue_shadowing(plugins.run_l96_euler, %r, cp, %d, %d, epsilon=%r, return_checkpoint=True)\n"
            "np.savez(sys.argv[1], J=np.array(cp2.J), Rs=np.array(cp2.lss.Rs), G=fds.lss_gradient(cp2))\n"
            % (ROOT, str(tmp_path), float(g['s']), int(g['K1']), int(g['steps']), float(g['eps'])))
    out = str(tmp_path / 'cont.npz')
    r = subprocess.run([sys.executable, '-W', 'ignore', '-c', code, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    got = np.load(out)
    assert np.array_equal(got['J'], g['cont_J']) and np.array_equal(got['Rs'], g['cont_Rs'])
    assert np.array_equal(got['G'], g['cont_G'])


def test_every_operator_family_of_a_stored_graph():
    """tests/golden/refcp_ops_m3_segment1: a reference-format checkpoint whose lazy V and v run through neg,
    transpose, sum over an axis, pow, div, outer (reshape + broadcast), getitem with a slice, setitem, ravel, norm,
    dot and qr_transpose (oracle/gen_golden.py reference_graph_ops_fixture): the by-name evaluation gives what the
    reference's executor gives, bit for bit."""
    from fds_b200.refcheckpoint import load_reference_checkpoint
    want = golden('refcp_ops_expected')
    cp = load_reference_checkpoint(os.path.join(ROOT, 'tests', 'golden', 'refcp_ops_m3_segment1'))
    assert np.array_equal(cp.u0, want['u0']) and np.array_equal(cp.V, want['V']) and np.array_equal(cp.v, want['v'])
    assert np.array_equal(cp.lss.Rs[0], want['R']) and np.array_equal(cp.lss.bs[0], want['b'])
    assert np.abs(cp.V @ cp.V.T - np.eye(3)).max() < 1e-14
