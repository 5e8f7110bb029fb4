"""A short randomised parity sweep (tools/fuzz_parity.py: one segment bit for bit and the boundary after a
device-resident segment against the oracle port, over random sizes, subspace dimensions, step counts, halo
capacities, schemes and time-dilation orders; sizes around the tile of the fused time-dilation kernel first)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_randomised_parity_sweep():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_parity.py'), '7', '20'],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0 and 'FUZZ' in out.stdout and ' 0 failures' in out.stdout, tail
