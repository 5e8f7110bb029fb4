"""Multi-GPU self-check: the N-sharded loop against the same problem on ONE GPU.

Called under ``torch.distributed`` (one process per GPU, NCCL) by ``bench.py --gpus N`` before its timed
region, by ``tests/dist_check.py`` and -- through a torchrun child -- by ``__graft_entry__.smoke()``.  Every
rank owns a contiguous block of sites, exactly as the reference shards its state over MPI ranks
(fds/compute.py:87-89; halo exchange of tests/solvers/mock_fun3d/fun3d:29-37).  Rank 0 also runs the same
small problem on its own GPU alone and compares:

* the plugin call ``run(u0, s, steps)`` on the sharded state: final state and objective history bit-identical;
* ``shadowing`` (several halo exchanges per segment): segment-0 objective history and final primal state
  bit-identical, ``R_0`` to 1e-10 (the Gram is summed in a different order), dJ/ds to 1e-6;
* both communication modes: the library's own NCCL communicator (one enqueue per boundary) and the
  host-driven ``torch.distributed`` ring;
* the tensor-core, device-gated boundary (m = 16) over the library's communicator, with the adaptive single pass
  and with CholeskyQR2 forced: objective histories and final state bit-identical, R and the tangents to 1e-10;
* the batched parameter sweep split over the ranks (replicas only): bit-identical to one GPU.

This is a consistency check of the product against itself on different device counts; parity with the
reference is the business of tests/ (oracle and golden fixtures).
"""
import numpy as np


def _relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def sharded_vs_single(local_device, sites_per_rank=4096, m=6, steps=30, K=4, with_sweep=True):
    """Returns the result dict on every rank (computed on rank 0, broadcast)."""
    import torch.distributed as dist
    import fds_b200
    from fds_b200.engine import Engine

    rank, world = dist.get_rank(), dist.get_world_size()
    N = int(sites_per_rank) * world
    s, eps = 0.1, 1e-4
    u0 = np.random.RandomState(0).rand(N)
    res = {'world': world, 'N': N, 'm': m, 'segments': K, 'steps': steps, 'halo_steps': 8}

    single = {}
    if rank == 0:
        eng1 = Engine(N, 1, device=local_device, distributed=False, max_halo_steps=8)
        eng1.set_state(u0)
        single['J'] = eng1.run(s, 25)
        single['u1'] = eng1.get_state()
        eng1.close()
        engm = Engine(N, m, device=local_device, distributed=False, max_halo_steps=8)

        class _Single:
            def engine(self, mm):
                return engm
        np.random.seed(5)
        cp1 = fds_b200.shadowing(_Single(), u0, s, m, K, steps, 50, epsilon=eps, return_checkpoint=True,
                                 tangent_seed=7)
        single['cp'], single['G'] = cp1, fds_b200.lss_gradient(cp1)
        engm.close()

    ok = True
    for comm in ('nccl', 'torch'):
        run = fds_b200.L96(N, device=local_device, max_halo_steps=8, comm=comm)   # sharded: picks up the process group
        u1, J = run(u0, s, 25)                                                    # plugin contract: global arrays in / out
        np.random.seed(5)
        cp = fds_b200.shadowing(run, u0, s, m, K, steps, 50, epsilon=eps, return_checkpoint=True, tangent_seed=7)
        G = fds_b200.lss_gradient(cp)
        coll = run.engine(m).ctx.collectives
        run.close()
        if rank == 0:
            cp1 = single['cp']
            r = {'primal_bitwise': bool(np.array_equal(u1, single['u1']) and np.array_equal(J, single['J'])),
                 'J0_bitwise': bool(np.array_equal(np.array(cp.J)[0], np.array(cp1.J)[0])),
                 'final_state_bitwise': bool(np.array_equal(cp.u0, cp1.u0)),
                 'R0_rel_err': _relerr(np.array(cp.lss.Rs)[0], np.array(cp1.lss.Rs)[0]),
                 'djds_rel_err': _relerr(G, single['G']),
                 'library_collectives': int(coll)}
            r['pass'] = bool(r['primal_bitwise'] and r['J0_bitwise'] and r['final_state_bitwise'] and
                             r['R0_rel_err'] < 1e-10 and r['djds_rel_err'] < 1e-6 and
                             (coll > 0) == (comm == 'nccl'))
            res[comm] = r
            ok = ok and r['pass']

    # the tensor-core, device-gated boundary (m = 16: what the headline workload runs with m = 32) over the
    # library's communicator: Gram all-reduce inside the enqueue-only boundary, including a forced CholeskyQR2
    for passes in (0, 2):
        mg, Kg, Sg = 16, 3, 10
        cpg1 = None
        if rank == 0:
            e1 = Engine(N, mg, device=local_device, distributed=False, max_halo_steps=8)
            e1.ctx.set_qr_passes(passes)

            class _S1:
                def engine(self, mm):
                    return e1
            cpg1 = fds_b200.shadowing(_S1(), u0, s, mg, Kg, Sg, 20, epsilon=eps, return_checkpoint=True, tangent_seed=3)
            e1.close()
        rung = fds_b200.L96(N, device=local_device, max_halo_steps=8, comm='nccl')
        rung.engine(mg).ctx.set_qr_passes(passes)
        cpg = fds_b200.shadowing(rung, u0, s, mg, Kg, Sg, 20, epsilon=eps, return_checkpoint=True, tangent_seed=3)
        rung.close()
        if rank == 0:
            r = {'J_bitwise': bool(np.array_equal(np.array(cpg.J), np.array(cpg1.J))),
                 'final_state_bitwise': bool(np.array_equal(cpg.u0, cpg1.u0)),
                 'R_rel_err': _relerr(np.array(cpg.lss.Rs), np.array(cpg1.lss.Rs)),
                 'V_rel_err': _relerr(cpg.V, cpg1.V)}
            r['pass'] = bool(r['J_bitwise'] and r['final_state_bitwise'] and r['R_rel_err'] < 1e-10 and r['V_rel_err'] < 1e-9)
            res['gated_m16_%s' % ('adaptive' if passes == 0 else 'choleskyqr2')] = r
            ok = ok and r['pass']

    if with_sweep:
        # batched sweep split over the ranks (replicas only) against the same sweep on one GPU
        B, Nb, mb = 3 * world, 128, 4
        params = np.linspace(-0.4, 0.5, B)
        rng = np.random.RandomState(9)
        Ub, Wb = rng.rand(B, Nb), rng.rand(B, mb, Nb)
        runb = fds_b200.L96(Nb, device=local_device)
        Jd, Gd, Hd = fds_b200.shadowing_batch(runb, Ub, params, mb, 3, 20, 40, epsilon=eps, tangents=Wb,
                                              return_histories=True)
        runb.close()
        if rank == 0:
            from fds_b200 import fds as core
            r1 = fds_b200.L96(Nb, device=local_device)
            Js, Gs, Hs = core._shadowing_batch_local(r1, Ub, params, mb, 3, 20, 40, eps, None, None, True, Wb)
            r1.close()
            res['sweep_bitwise'] = bool(np.array_equal(Jd, Js) and np.array_equal(Hd['Rs'], Hs['Rs']) and
                                        np.array_equal(Gd, Gs) and Jd.shape == (B, 2))
            ok = ok and res['sweep_bitwise']
    res['pass'] = bool(ok)
    box = [res]
    dist.broadcast_object_list(box, src=0)
    return box[0]
