"""Run-up speed at the BASELINE size: one-step kernel vs the fused multi-step kernel (FDS_PRIMAL_FUSED)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fds_b200

N, steps = 1 << 24, 1000
eng = fds_b200.L96(N).engine(32)
eng.set_state(np.random.RandomState(0).rand(N))
for want in (False, True):
    eng.run(0.1, 50, want_J=want)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run(0.1, steps, want_J=want)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('N=2^24 %d steps, J %s: %.4f s = %.4f ms/step (fused=%s)' % (steps, 'recorded' if want else 'dropped', dt, dt / steps * 1e3,
                                                              os.environ.get('FDS_PRIMAL_FUSED', '1')))
eng.close()
