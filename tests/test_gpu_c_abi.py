"""The boundary is a C ABI: a plain C program (tests/c_abi/segment.c, gcc, no Python, no torch) drives
the library and must produce, bit for bit, what the Python mirror produces through ctypes."""
import os
import subprocess

import numpy as np
import pytest

import fds_b200
from oracle import plugins
from conftest import ROOT

gpu = pytest.mark.gpu


def test_c_client_compiles_against_the_header(tmp_path):
    """CPU: the header is valid C99 and the client links against the built library."""
    exe = str(tmp_path / 'c_abi_segment')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c_abi', 'segment.c'), '-L', os.path.join(ROOT, 'fds_b200'),
                           '-lfds_b200', '-Wl,-rpath,' + os.path.join(ROOT, 'fds_b200'), '-o', exe])
    assert os.path.exists(exe)


@gpu
def test_c_client_equals_python_mirror(tmp_path):
    N, m, steps, eps, s = 512, 6, 11, 1e-4, 0.1
    rng = np.random.RandomState(0)
    u = plugins.run_l96_rk4(rng.rand(N), s, 100)[0]
    V = np.linalg.qr(rng.rand(m, N).T)[0].T
    v = rng.randn(N) / np.sqrt(N)
    exe, fin, fout = str(tmp_path / 'seg'), str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')
    subprocess.check_call(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c_abi', 'segment.c'), '-L', os.path.join(ROOT, 'fds_b200'),
                           '-lfds_b200', '-Wl,-rpath,' + os.path.join(ROOT, 'fds_b200'), '-o', exe])
    np.concatenate([u, V.ravel(), v]).tofile(fin)
    out = subprocess.run([exe, fin, fout, str(N), str(m), str(steps), repr(eps)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert 'sm_100a' in out.stdout
    raw = np.fromfile(fout)
    sizes = [N, m * N, N, steps * (m + 2) * 2, m * m, m, m, 1]
    cu, cV, cv, cJ, cR, cb, cGd, cgd = np.split(raw, np.cumsum(sizes)[:-1])
    # the same sequence through the Python mirror
    eng = fds_b200.L96(N).engine(m)
    eng.set_state(u)
    eng.set_tangents(V, v)
    eng.boundary(s, eps, from_ensemble=0, do_qr=0, build=1, keep_tangents=False)
    J = eng.ctx.segment(s, eps, steps)
    R, b, Gd, gd = eng.boundary(s, eps, from_ensemble=1, do_qr=1, build=0)
    pV, pv = eng.get_tangents()
    assert np.array_equal(cu, eng.get_state()) and np.array_equal(cV.reshape(m, N), pV) and np.array_equal(cv, pv)
    assert np.array_equal(cJ.reshape(steps, m + 2, 2), J)
    assert np.array_equal(cR.reshape(m, m), R) and np.array_equal(cb, b)
    assert np.array_equal(cGd, Gd) and cgd[0] == gd
