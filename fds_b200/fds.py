"""Public API of the shadowing hot path: ``shadowing``, ``continue_shadowing``,
``lss_gradient`` -- same signatures, argument meaning and return values as the
reference ``fds/fds.py`` (:29-51, :92-176, :178-220).

``run`` must be a device plugin (``fds_b200.L96``, ``Lorenz63``, ...).  The whole segment loop
-- ensemble time stepping, finite-difference tangents, time dilation, projection,
QR, next perturbed ensemble -- runs in CUDA kernels on device-resident state;
per segment only KBs (R, b, G_dil, g_dil, the objective sums) come back to the
host.  There is no CPU fallback: an arbitrary Python ``run`` raises TypeError.  A solver
that cannot move to the GPU is wrapped explicitly, ``fds_b200.HostSolver(run)``: it advances
the trajectories on the host, the algebra of the loop still runs on the device.
"""
import sys
from copy import deepcopy

import numpy as np

from .checkpoint import Checkpoint, verify_checkpoint, save_checkpoint
from .lsstan import LssTangent
from .segment import trapez_mean, sensitivities
from . import timedilation
from .timedilation import stencil_weights
from .timeseries import windowed_mean

# host <-> device random-tangent switch: below this many elements the (m, N) block is
# drawn on the host with np.random.rand like the reference (fds/compute.py:38-42)
HOST_RNG_MAX_ELEMS = 1 << 22


def _engine_of(run, m):
    eng = getattr(run, 'engine', None)
    if eng is None:
        raise TypeError('fds_b200.shadowing needs a device plugin (e.g. fds_b200.L96); '
                        'arbitrary Python solvers are not run on the CPU behind your back -- wrap a '
                        'host solver explicitly with fds_b200.HostSolver(run)')
    return eng(m)


def _dilation_mode(run_ddt):
    """run_ddt argument of the reference -> Engine.set_time_dilation mode."""
    if run_ddt is None:
        return 1
    if callable(run_ddt) or isinstance(run_ddt, str):
        return run_ddt
    if run_ddt == 0:
        return 0
    raise TypeError('run_ddt: None, 0, "analytic" or a callable run_ddt(u0, parameter)')


def lss_gradient_series(checkpoint, segment_range=None):
    """Per-segment series (grad_lss, grad_dil) whose means add up to dJ/ds."""
    _, _, _, lss, G_lss, g_lss, J, G_dil, g_dil = checkpoint
    if segment_range is not None:
        sl = slice(segment_range) if isinstance(segment_range, int) else slice(*segment_range)
        lss = deepcopy(lss)
        lss.Rs, lss.bs = lss.Rs[sl], lss.bs[sl]
        G_lss, g_lss, J, G_dil, g_dil = G_lss[sl], g_lss[sl], J[sl], G_dil[sl], g_dil[sl]
    alpha = lss.solve()
    grad_lss = (alpha[:, :, np.newaxis] * np.array(G_lss)).sum(1) + np.array(g_lss)
    J = np.array(J)
    dJ = trapez_mean(J.mean(0), 0) - J[:, -1]
    dil = ((alpha * np.array(G_dil)).sum(1) + np.array(g_dil)) / J.shape[1]
    return grad_lss, dil[:, np.newaxis] * dJ


def lss_gradient(checkpoint, segment_range=None):
    """dJ/ds from the histories in a checkpoint (reference fds/fds.py:29-51)."""
    grad_lss, grad_dil = lss_gradient_series(checkpoint, segment_range)
    return windowed_mean(grad_lss) + windowed_mean(grad_dil)


def continue_shadowing(run, parameter, checkpoint, num_segments, steps_per_segment,
                       epsilon=1E-6, checkpoint_path=None, checkpoint_interval=1,
                       simultaneous_runs=None, run_ddt=None, return_checkpoint=False,
                       get_host_dir=None, spawn_compute_job=None, verbose=False):
    """Run segments K0 .. num_segments-1 (K0 = segments already in ``checkpoint``);
    ``num_segments`` is the TOTAL target (reference fds/fds.py:92-176).
    ``run_ddt``: None = finite-difference time dilation (the reference's default), 0 = off
    (fds/timedilation.py:59-61), ``'analytic'`` = the built-in system's ``dt * f(u)`` on the device, or a
    callable ``run_ddt(u0, parameter) -> du per step`` (the reference's TimeDilationExact, whose upstream
    implementation raises NameError at fds/timedilation.py:58).
    ``simultaneous_runs`` sets the worker count of a ``HostSolver`` (fds/segment.py:37) and is inert for
    device plugins; ``get_host_dir(run_id)`` says where a ``HostSolver`` in file mode puts the ``input.h5`` of a
    run (fds/segment.py:32-33,46,51); ``spawn_compute_job`` is inert (the algebra it would spawn MPI ranks for
    runs on the device)."""
    assert verify_checkpoint(checkpoint)
    u0, V, v, lss, G_lss, g_lss, J_hist, G_dil, g_dil = checkpoint
    V = np.asarray(V)
    eng = _engine_of(run, V.shape[0])
    eng.set_time_dilation(_dilation_mode(run_ddt))
    eng.set_dxdt_weights(stencil_weights(timedilation.ORDER_OF_ACCURACY))
    if hasattr(eng, 'set_host_dir'):
        eng.set_host_dir(get_host_dir)
    eng.set_state(u0)
    eng.set_tangents(V, v)
    if hasattr(eng, 'i_segment'):
        eng.i_segment = lss.K_segments()       # run_id numbering of a host solver continues
    if simultaneous_runs is not None and hasattr(eng, 'set_simultaneous_runs'):
        eng.set_simultaneous_runs(simultaneous_runs)
    return _segment_loop(eng, parameter, lss, G_lss, g_lss, J_hist, G_dil, g_dil, num_segments,
                         steps_per_segment, epsilon, checkpoint_path, checkpoint_interval,
                         return_checkpoint, verbose)


def _is_writer(eng):
    """Rank that writes checkpoints (every rank holds the same gathered arrays)."""
    ring = getattr(eng, 'ring', None)
    return ring is None or ring.rank == 0


def _save(eng, checkpoint_path, cp):
    """Checkpoint written once (rank 0), atomically, and visible to every rank when this returns."""
    if _is_writer(eng):
        save_checkpoint(checkpoint_path, cp)
    ring = getattr(eng, 'ring', None)
    if ring is not None:
        ring.dist.barrier(group=ring.group)


def _segment_loop(eng, parameter, lss, G_lss, g_lss, J_hist, G_dil, g_dil, num_segments,
                  steps, epsilon, checkpoint_path, checkpoint_interval, return_checkpoint, verbose):
    def record(J):
        J0, G, g = sensitivities(J, epsilon)
        J_hist.append(J0), G_lss.append(G), g_lss.append(g)

    def snapshot():
        V, v = eng.get_tangents()
        return Checkpoint(eng.get_state(), V, v, lss, G_lss, g_lss, J_hist, G_dil, g_dil)

    K0 = lss.K_segments()
    # first boundary of this call: time dilation + projection only (fds/fds.py:113-126)
    eng.boundary(parameter, epsilon, from_ensemble=0, do_qr=0, build=1)
    nq = num_segments - K0
    if nq >= 1 and not verbose and not checkpoint_path and getattr(eng, 'deferred_ok', lambda: False)():
        # Device-resident loop without host round trips: every boundary and segment of this call is put on
        # the stream back to back (adaptive CholeskyQR decided on the device, halo exchange / reductions
        # over the library's communicator), the KB-sized per-segment results land in page-locked memory and
        # are read once at the end.  Same arithmetic as the step-by-step loop below.
        eng.begin_deferred(nq, steps)
        eng.segment_deferred(parameter, epsilon, steps)
        for i in range(K0 + 1, num_segments + 1):
            last = i == num_segments
            eng.boundary_deferred(parameter, epsilon, from_ensemble=1, do_qr=1, build=not last,
                                  keep_tangents=last and return_checkpoint)
            if not last:
                eng.segment_deferred(parameter, epsilon, steps)
        bnd, Js = eng.end_deferred()
        for (R, b, Gd, gd), J in zip(bnd, Js):
            record(J)
            lss.append(R, b)
            G_dil.append(Gd), g_dil.append(gd)
    else:
        record(eng.segment(parameter, epsilon, steps))
        for i in range(K0 + 1, num_segments + 1):
            last = i == num_segments
            save = bool(checkpoint_path) and i % checkpoint_interval == 0
            R, b, Gd, gd = eng.boundary(parameter, epsilon, from_ensemble=1, do_qr=1, build=not last,
                                        keep_tangents=save or (last and return_checkpoint))
            lss.append(R, b)
            G_dil.append(Gd), g_dil.append(gd)
            if verbose:
                print(lss_gradient(Checkpoint(None, None, None, lss, G_lss, g_lss, J_hist, G_dil, g_dil)))
                sys.stdout.flush()
            if save:
                _save(eng, checkpoint_path, snapshot())
            if not last:
                record(eng.segment(parameter, epsilon, steps))
    if return_checkpoint:
        return snapshot()
    cp = Checkpoint(None, None, None, lss, G_lss, g_lss, J_hist, G_dil, g_dil)
    return np.array(J_hist).mean((0, 1)), lss_gradient(cp)


def shadowing(run, u0, parameter, subspace_dimension, num_segments, steps_per_segment, runup_steps,
              epsilon=1E-6, checkpoint_path=None, checkpoint_interval=1, simultaneous_runs=None,
              run_ddt=None, return_checkpoint=False, get_host_dir=None, spawn_compute_job=None,
              verbose=False, tangent_seed=None):
    """dJ/ds of the long-time average J of a chaotic system by finite-difference
    non-intrusive least-squares shadowing (reference fds/fds.py:178-220).

    run                 device plugin (``fds_b200.L96``)
    u0                  initial state, flat (N,)
    parameter           the scalar s
    subspace_dimension  m, number of homogeneous tangents (> number of positive Lyapunov exponents)
    num_segments, steps_per_segment, runup_steps, epsilon   as in the reference
    returns             (time-averaged J (n_qoi,), dJ/ds (n_qoi,)) or the Checkpoint
    """
    m = int(subspace_dimension)
    eng = _engine_of(run, m)
    eng.set_time_dilation(_dilation_mode(run_ddt))
    eng.set_dxdt_weights(stencil_weights(timedilation.ORDER_OF_ACCURACY))
    if hasattr(eng, 'set_host_dir'):
        eng.set_host_dir(get_host_dir)
    eng.set_state(u0)
    if hasattr(eng, 'i_segment'):
        eng.i_segment = 0                      # run_id numbering of a host solver starts over (fds/segment.py:38-49)
    if simultaneous_runs is not None and hasattr(eng, 'set_simultaneous_runs'):
        eng.set_simultaneous_runs(simultaneous_runs)
    if runup_steps > 0:
        eng.run(parameter, runup_steps, want_J=False)
    # random orthonormal tangents, v = 0 (fds/fds.py:21-27)
    if tangent_seed is None and m * eng.n_global <= HOST_RNG_MAX_ELEMS:
        eng.orthonormal_random_tangents(host_rng=np.random)
    else:
        seed = np.random.randint(2 ** 31 - 1) if tangent_seed is None else tangent_seed
        eng.orthonormal_random_tangents(seed=seed)
    return _segment_loop(eng, parameter, LssTangent(device_ordinal=eng.device), [], [], [], [], [],
                         num_segments, steps_per_segment, epsilon, checkpoint_path,
                         checkpoint_interval, return_checkpoint, verbose)


# --------------------------------------------------------------------------- #
# Batched parameter sweep (BASELINE configs[4]): B independent shadowing problems of the
# same system in ONE device context.  The reference can only loop over them
# (tests/test_lorenz.py:35-41, apps/lorenz_demo.py:86-89); the result of problem p equals
# shadowing(run, u0[p], parameters[p], ...) on its own.
# --------------------------------------------------------------------------- #
def lss_gradient_batch(Rs, bs, G_lss, g_lss, J, G_dil, g_dil, device=0):
    """Vectorised lss_gradient (reference fds/fds.py:29-51) over a leading problem axis.
    Rs (B,K,m,m), bs (B,K,m), G_lss (B,K,m,q), g_lss (B,K,q), J (B,K,steps,q), G_dil (B,K,m),
    g_dil (B,K) -> dJ/ds (B,q)."""
    from . import device as dev
    alpha = dev.lss_solve(Rs, bs, device)                                    # (B,K,m)
    grad_lss = (alpha[..., np.newaxis] * G_lss).sum(2) + g_lss                 # (B,K,q)
    dJ = trapez_mean(J.mean(1), 1)[:, np.newaxis] - J[:, :, -1]                # (B,K,q)
    dil = ((alpha * G_dil).sum(2) + g_dil) / J.shape[2]                        # (B,K)
    grad_dil = dil[..., np.newaxis] * dJ
    K = grad_lss.shape[1]
    return grad_lss.sum(1) / K + grad_dil.sum(1) / K


def _dist_world(group):
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    return None, 0, 1


def shadowing_batch(run, u0, parameters, subspace_dimension, num_segments, steps_per_segment,
                    runup_steps, epsilon=1E-6, run_ddt=None, tangent_seed=None, return_histories=False,
                    tangents=None, group=None):
    """``shadowing`` for a sweep of parameters: problem p starts from ``u0[p]`` (or the shared
    ``u0`` if it is one-dimensional) and runs at ``parameters[p]``.
    Returns (J (B, n_qoi), dJ/ds (B, n_qoi)); with ``return_histories`` also the dict of the
    per-segment histories (Rs, bs, G_lss, g_lss, J, G_dil, g_dil, each with a leading B axis).
    ``tangents`` (B, m, N): initial tangent blocks to orthonormalise instead of random ones.
    Under torch.distributed (one process per GPU) the sweep is split into contiguous slices, one
    per rank -- the problems are independent, so there is no data-path collective -- and every
    rank returns the gathered result."""
    dist, rank, world = _dist_world(group)
    if world > 1:
        from .ring import shard_range
        parameters = np.ascontiguousarray(parameters, dtype=np.float64).ravel()
        lo, hi = shard_range(parameters.size, rank, world)
        if hi - lo < 2:
            raise ValueError('shadowing_batch: fewer than two problems per rank')
        u0 = np.asarray(u0, dtype=np.float64)
        out = _shadowing_batch_local(run, u0 if u0.ndim == 1 else u0[lo:hi], parameters[lo:hi], subspace_dimension,
                                     num_segments, steps_per_segment, runup_steps, epsilon, run_ddt,
                                     None if tangent_seed is None else tangent_seed + rank, return_histories,
                                     None if tangents is None else np.asarray(tangents)[lo:hi])
        parts = [None] * world
        dist.all_gather_object(parts, out, group=group)
        res = [np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])]
        if return_histories:
            res.append({k: np.concatenate([p[2][k] for p in parts]) for k in parts[0][2]})
        return tuple(res)
    return _shadowing_batch_local(run, u0, parameters, subspace_dimension, num_segments, steps_per_segment,
                                  runup_steps, epsilon, run_ddt, tangent_seed, return_histories, tangents)


def _shadowing_batch_local(run, u0, parameters, subspace_dimension, num_segments, steps_per_segment,
                           runup_steps, epsilon, run_ddt, tangent_seed, return_histories, tangents):
    parameters = np.ascontiguousarray(parameters, dtype=np.float64).ravel()
    B, m = parameters.size, int(subspace_dimension)
    if callable(run_ddt):
        raise NotImplementedError('shadowing_batch: run_ddt may be None, 0 or "analytic"')
    if B < 2:
        raise ValueError('shadowing_batch needs at least two problems; use shadowing() for one')
    be = getattr(run, 'batch_engine', None)
    if be is None:
        raise TypeError('shadowing_batch needs a device plugin (e.g. fds_b200.L96)')
    eng = be(m, B)
    N = eng.n_global
    u0 = np.asarray(u0, dtype=np.float64)
    if u0.ndim == 1:
        u0 = np.broadcast_to(u0, (B, N))
    eng.set_time_dilation(_dilation_mode(run_ddt))
    eng.set_dxdt_weights(stencil_weights(timedilation.ORDER_OF_ACCURACY))
    eng.ctx.set_batch_params(parameters)          # problem p runs at 0.0 + parameters[p]
    eng.set_state(np.ascontiguousarray(u0))
    s = 0.0
    if runup_steps > 0:
        eng.run(s, runup_steps, want_J=False)
    if tangents is not None:
        eng.orthonormal_random_tangents(block=np.ascontiguousarray(tangents, dtype=np.float64))
    elif tangent_seed is None and B * m * N <= HOST_RNG_MAX_ELEMS:
        eng.orthonormal_random_tangents(host_rng=np.random)
    else:
        seed = np.random.randint(2 ** 31 - 1) if tangent_seed is None else tangent_seed
        eng.orthonormal_random_tangents(seed=seed)

    Rs, bs, G_lss, g_lss, J_hist, G_dil, g_dil = [], [], [], [], [], [], []

    def record_segment():
        J0, G, g = eng.segment_sens(s, epsilon, steps_per_segment)      # (B,steps,q), (B,m,q), (B,q)
        J_hist.append(J0), G_lss.append(G.copy()), g_lss.append(g.copy())

    eng.boundary(s, epsilon, from_ensemble=0, do_qr=0, build=1)
    record_segment()
    for i in range(1, num_segments + 1):
        last = i == num_segments
        R, b, Gd, gd = eng.boundary(s, epsilon, from_ensemble=1, do_qr=1, build=not last, keep_tangents=False)
        Rs.append(R), bs.append(b), G_dil.append(Gd), g_dil.append(gd)
        if not last:
            record_segment()
    H = dict(Rs=np.stack(Rs, 1), bs=np.stack(bs, 1), G_lss=np.stack(G_lss, 1), g_lss=np.stack(g_lss, 1),
             J=np.stack(J_hist, 1), G_dil=np.stack(G_dil, 1), g_dil=np.stack(g_dil, 1))
    G = lss_gradient_batch(H['Rs'], H['bs'], H['G_lss'], H['g_lss'], H['J'], H['G_dil'], H['g_dil'], eng.device)
    Jm = H['J'].mean((1, 2))
    return (Jm, G, H) if return_histories else (Jm, G)
