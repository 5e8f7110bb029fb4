"""Where a segment of the batched sweep (4096 x N=4096, m=16) spends its time: boundary vs steps vs sensitivities."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fds_b200

B, N, m, steps, eps = 4096, 4096, 16, 100, 1e-4
H = int(sys.argv[1]) if len(sys.argv) > 1 else 0
run = fds_b200.L96(N, distributed=False, max_halo_steps=H)
print('max_halo_steps', H or 'default (16)')
eng = run.batch_engine(m, B)
eng.ctx.set_batch_params(np.linspace(-2.0, 2.0, B))
eng.set_state(np.random.RandomState(0).rand(B, N))
eng.run(0.0, 200, want_J=False)
eng.orthonormal_random_tangents(seed=1)
eng.boundary(0.0, eps, 0, 0, 1)
eng.segment(0.0, eps, steps)
for it in range(3):
    eng.ctx.sync(); t0 = time.perf_counter()
    R = eng.boundary(0.0, eps, 1, 1, 1, keep_tangents=False)[0]
    eng.ctx.sync(); t1 = time.perf_counter()
    J = eng.segment_sens(0.0, eps, steps)[1]
    eng.ctx.sync(); t2 = time.perf_counter()
    print('boundary %.2f ms   segment+sens %.2f ms   launches so far %d' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, eng.ctx.launches))
run.close()
