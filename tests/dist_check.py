"""Multi-GPU consistency check, launched by tests/test_gpu_dist.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_check.py

Every rank owns a contiguous block of sites (N-sharding with halo exchange over NCCL
send/recv, one all-reduce per Gram).  Rank 0 also runs the same problem on its single GPU
and compares: primal and objective bit-identical, one segment's ensemble bit-identical,
R / b / dJ/ds to rounding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fds_b200  # noqa: E402
from fds_b200.engine import Engine  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    N, m, steps, K = 4096 * world, 6, 30, 4
    s, eps = 0.1, 1e-4
    u0 = np.random.RandomState(0).rand(N)

    run = fds_b200.L96(N, device=local, max_halo_steps=8)       # sharded: picks up the process group
    # --- plugin contract on sharded state: global arrays in / out
    u1, J = run(u0, s, 25)
    # --- full shadowing, sharded
    np.random.seed(5)
    cp = fds_b200.shadowing(run, u0, s, m, K, steps, 50, epsilon=eps, return_checkpoint=True, tangent_seed=7)
    G = fds_b200.lss_gradient(cp)
    ok = True
    if rank == 0:
        single = fds_b200.L96(N, device=local, max_halo_steps=8)
        single.kw['group'] = None
        eng1 = Engine(N, 1, device=local, distributed=False, max_halo_steps=8)
        eng1.set_state(u0)
        J1 = eng1.run(s, 25)
        ok &= bool(np.array_equal(u1, eng1.get_state()))
        ok &= bool(np.array_equal(J, J1))
        print('primal bit-identical over %d ranks: %s' % (world, ok))
        # single-GPU shadowing with the same device-generated tangents
        engm = Engine(N, m, device=local, distributed=False, max_halo_steps=8)

        class _Single:
            def engine(self, mm):
                return engm
        np.random.seed(5)
        cp1 = fds_b200.shadowing(_Single(), u0, s, m, K, steps, 50, epsilon=eps, return_checkpoint=True,
                                 tangent_seed=7)
        G1 = fds_b200.lss_gradient(cp1)
        jj = bool(np.array_equal(np.array(cp.J)[0], np.array(cp1.J)[0]))
        rr = float(np.abs(np.array(cp.lss.Rs)[0] - np.array(cp1.lss.Rs)[0]).max() / np.abs(np.array(cp1.lss.Rs)[0]).max())
        gg = float(np.abs(G - G1).max() / np.abs(G1).max())
        uu = bool(np.array_equal(cp.u0, cp1.u0))
        print('segment-0 objective bit-identical: %s; final primal bit-identical: %s; R0 rel err %.1e; dJ/ds rel err %.1e'
              % (jj, uu, rr, gg))
        ok &= jj and uu and rr < 1e-10 and gg < 1e-6
    # --- batched sweep split over the ranks (replicas only) vs the same sweep on one GPU
    B, Nb, mb = 3 * world, 128, 4
    params = np.linspace(-0.4, 0.5, B)
    rng = np.random.RandomState(9)
    Ub, Wb = rng.rand(B, Nb), rng.rand(B, mb, Nb)
    runb = fds_b200.L96(Nb, device=local)
    Jd, Gd, Hd = fds_b200.shadowing_batch(runb, Ub, params, mb, 3, 20, 40, epsilon=eps, tangents=Wb,
                                          return_histories=True)
    if rank == 0:
        from fds_b200 import fds as core
        Js, Gs, Hs = core._shadowing_batch_local(fds_b200.L96(Nb, device=local), Ub, params, mb, 3, 20, 40, eps,
                                                 None, None, True, Wb)
        bb = bool(np.array_equal(Jd, Js) and np.array_equal(Hd['Rs'], Hs['Rs']) and np.array_equal(Gd, Gs))
        print('batched sweep over %d ranks bit-identical to one GPU: %s' % (world, bb))
        ok &= bb and Jd.shape == (B, 2)
        print('DIST_CHECK_OK' if ok else 'DIST_CHECK_FAILED')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
