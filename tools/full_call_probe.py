"""One complete public-API call at the BASELINE size: fds_b200.shadowing(run, u0, s, m, K, steps, runup) from a host
u0 to (J, dJ/ds) -- run-up, random orthonormal tangents, K segments, LSS solve."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fds_b200

N, m, K, steps, runup = 1 << 24, 32, 20, 100, 1000
run = fds_b200.L96(N)
u0 = np.random.RandomState(0).rand(N)
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    J, G = fds_b200.shadowing(run, u0, 0.0, m, K, steps, runup, epsilon=1e-6 * np.sqrt(N), tangent_seed=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('shadowing(N=2^24, m=32, K=%d, steps=%d, runup=%d): %.3f s  -> %.3e state-updates/s incl. run-up and set-up; J=%s G=%s'
          % (K, steps, runup, dt, N * (m + 1) * steps * K / dt, J, G))
run.close()
