"""Install the UNMODIFIED reference into oracle/_ref (TEST INFRASTRUCTURE / CPU baseline only).

    python -m oracle.build_ref

The reference is pure Python, so "building" it is the one offline pip install the task allows: the tree
under /root/reference is copied to a scratch directory (the source tree is read-only and setuptools writes
build/ and *.egg-info next to setup.py) and installed with

    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <scratch copy>

oracle/_ref is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels
to the GPU box, where `bench.py --impl reference` and the `cpu_baseline` leg import `fds.shadowing` from it
(fds/fds.py:178-220) and time it with the reference's own multiprocessing pool (fds/segment.py:37).  Its
dependencies (numpy, scipy, dill) come from the image.  Nothing in fds_b200/ imports it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference'
TARGET = os.path.join(HERE, '_ref')


def build_ref(force=False):
    """Returns the install directory, or None when /root/reference is not there (the GPU box)."""
    if not os.path.isdir(os.path.join(REF_SRC, 'fds')):
        return TARGET if os.path.isdir(os.path.join(TARGET, 'fds')) else None
    if os.path.isdir(os.path.join(TARGET, 'fds')) and not force:
        return TARGET
    tmp = tempfile.mkdtemp(prefix='fds_ref_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns('.git', '__pycache__', 'docs', 'apps', 'tools', 'debug'))
        shutil.rmtree(TARGET, ignore_errors=True)
        subprocess.check_call([sys.executable, '-m', 'pip', 'install', '--quiet', '--no-index', '--no-build-isolation',
                               '--no-deps', '--target', TARGET, src])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == '__main__':
    print(build_ref(force='--force' in sys.argv))
