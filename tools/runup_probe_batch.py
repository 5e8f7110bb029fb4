"""Run-up of the batched sweep (4096 problems of N = 4096): fused multi-step march vs the one-step kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fds_b200.device import Context

B, N, steps = 4096, 4096, 512
bc = Context(N, 16, batch=B, max_halo_steps=16)
bc.set_batch_params(np.linspace(-0.5, 0.5, B))
bc.set_state(np.random.RandomState(0).rand(B, N))
bc.run(0.0, 32, want_J=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
bc.run(0.0, steps, want_J=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('%d x N=%d, %d steps: %.4f s = %.4f ms/step (fused=%s)' % (B, N, steps, dt, dt / steps * 1e3, os.environ.get('FDS_PRIMAL_FUSED', '1')))
bc.close()
