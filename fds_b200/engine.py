"""Driver of one shard's device context through a segment boundary / a segment.

The C library splits a boundary at its reduction points (include/fds_b200.h);
this module is the host side that fills those points: periodic wrap on one GPU,
``Ring`` halo exchange + all-reduce when the state is sharded over ranks.  The
sequence is the reference's per-segment algorithm (SURVEY.md appendix A;
fds/fds.py:128-170, fds/timedilation.py:63-82, fds/lsstan.py:20-34,
fds/segment.py:43-79).
"""
import numpy as np

from . import _lib
from .device import Context, TOY_SYSTEMS, comm_unique_id
from .ring import Ring, shard_range


class _DevView:
    """Expose a span of a library-owned device buffer through __cuda_array_interface__."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {'shape': (int(count),), 'typestr': '<f8',
                                         'data': (int(ptr), False), 'version': 2}


class Engine:
    def __init__(self, n_global, m, scheme='rk4', math='exact', dt=0.01, device=None,
                 group=None, distributed=None, max_halo_steps=0, simple_kernels=False, stream=None, batch=1,
                 comm='nccl'):
        """``comm`` (sharded state only): ``'nccl'`` -- the library holds its own NCCL communicator and a
        boundary / segment is ONE enqueue on the context's stream (halo send/recv, Gram all-reduce and the
        fixed-order objective sum inside, no host round trip); ``'torch'`` -- the host drives the phases and
        fills the reduction points with ``torch.distributed`` calls (``Ring``; also what the gloo CPU tests
        exercise)."""
        self.n_global = int(n_global)
        self.batch = int(batch)
        if self.batch > 1:
            distributed = False        # a batch lives on one GPU; ranks take disjoint slices of the sweep
        self.m = int(m)
        self.M = self.m + 2
        # Lorenz-96: the device returns SUMS over sites of [x, x^2] -> divide by N;
        # small ODE systems: objective values as they are, n_qoi of the 2 slots used
        self.j_div = 1.0 if scheme in TOY_SYSTEMS else float(self.n_global)
        self.n_qoi = TOY_SYSTEMS[scheme][1] if scheme in TOY_SYSTEMS else 2
        if scheme in TOY_SYSTEMS:
            distributed = False
        self.ring = None
        self.torch = None
        if distributed is None:
            try:
                import torch.distributed as dist
                distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
            except ImportError:
                distributed = False
        if distributed:
            import torch
            self.torch = torch
            self.ring = Ring(group)
            if device is None:
                device = torch.cuda.current_device()
            torch.cuda.set_device(device)
            if stream is None:
                stream = torch.cuda.current_stream().cuda_stream
        if device is None:
            device = 0
        self.device = int(device)
        rank, world = (self.ring.rank, self.ring.world) if self.ring else (0, 1)
        self.lo, self.hi = shard_range(self.n_global, rank, world)
        self.ctx = Context(self.hi - self.lo, m, scheme=scheme, math=math, dt=dt, device=self.device,
                           n_global=self.n_global, site_offset=self.lo, rank=rank, world=world,
                           max_halo_steps=max_halo_steps, simple_kernels=simple_kernels, stream=stream,
                           batch=self.batch)
        self.time_dilation = True
        self.dil_mode = 1
        self.run_ddt = None
        self._dq = None
        # whole boundaries / segments inside the library: one GPU, or ranks joined by the library's communicator
        self.lib_comm = self.ring is not None and comm != 'torch'
        if self.lib_comm:
            ids = [comm_unique_id() if rank == 0 else None]
            self.ring.dist.broadcast_object_list(ids, src=self.ring._global(0), group=self.ring.group)
            self.ctx.comm_init(ids[0])
        self.in_library = self.ring is None or self.lib_comm

    # -- sharding of host arrays --------------------------------------------- #
    def local(self, a):
        """Slice the trailing (site) axis of a global host array to this shard."""
        a = np.asarray(a, dtype=np.float64)
        if a.shape[-1] == self.n_global and self.ring is not None:
            return np.ascontiguousarray(a[..., self.lo:self.hi])
        return a

    def gather_sites(self, a):
        """Concatenate per-rank host arrays along the site axis (all ranks get the result)."""
        if self.ring is None:
            return a
        parts = [None] * self.ring.world
        self.ring.dist.all_gather_object(parts, a, group=self.ring.group)
        return np.concatenate(parts, axis=-1)

    # -- device views / exchanges --------------------------------------------- #
    def _view(self, which, off, cnt):
        base = self.ctx.buffer_ptr(which)
        return self.torch.as_tensor(_DevView(base + 8 * off, cnt), device='cuda:%d' % self.device)

    def _halo(self, which, hs):
        if self.ring is None:
            (self.ctx.halo_wrap_ensemble if which == _lib.FDS_BUF_ENSEMBLE else self.ctx.halo_wrap_primal)(hs)
            return
        off, cnt = self.ctx.halo_layout(which, hs)
        v = [self._view(which, o, c) for o, c in zip(off, cnt)]
        self.ring.exchange_halo(v[0], v[1], v[2], v[3])
        self.ctx.halo_mark_valid(which, hs)

    def _reduce_gram(self):
        if self.ring is not None:
            self.ring.allreduce_sum(self._view(_lib.FDS_BUF_GRAM, 0, (self.m + 2) ** 2))

    # -- the solver plugin on the primal --------------------------------------- #
    def set_dxdt_weights(self, w):
        self.ctx.set_dxdt_weights(w)

    def set_time_dilation(self, mode):
        """How the time-dilation vector is obtained: True / 1 finite differences of three primal steps
        (the reference's TimeDilation), False / 0 none (``run_ddt=0``), ``'analytic'`` / 2 the built-in
        system's ``dt * f(u)``, or a callable ``run_ddt(u0, parameter) -> du per step`` evaluated on the
        host (the reference's TimeDilationExact, fds/timedilation.py:54-61)."""
        self.run_ddt = mode if callable(mode) else None
        if callable(mode):
            m = _lib.FDS_DILATION_GIVEN
        elif isinstance(mode, str):
            m = {'analytic': _lib.FDS_DILATION_ANALYTIC, 'fd': _lib.FDS_DILATION_FD, 'off': _lib.FDS_DILATION_OFF}[mode]
        else:
            m = int(mode)
        self.dil_mode = m
        self.time_dilation = m != 0
        if getattr(self, 'ctx', None) is not None:
            self.ctx.set_time_dilation(m)

    def _dilation_input(self, s):
        """Before fds_boundary_dxdt: the primal halo (finite differences / analytic) or the host's d."""
        if self.run_ddt is not None:
            d = np.asarray(self.run_ddt(self.get_state(), s), dtype=np.float64)
            self.ctx.set_dxdt(self.local(d))
        elif self.time_dilation:
            self._halo(_lib.FDS_BUF_PRIMAL, self.ctx.dilation_order)

    def run(self, s, steps, want_J=True):
        """Advance the primal; returns J (steps, 2) = [mean x, mean x^2] or None."""
        steps = int(steps)
        if self.in_library:
            Js = self.ctx.run(s, steps, want_J)
        else:
            self.ctx.boundary_begin(0)
            t = 0
            while t < steps:
                hs = self.ctx.halo_steps(steps - t)
                self._halo(_lib.FDS_BUF_PRIMAL, hs)
                self.ctx.primal_advance(s, t, hs, want_J)
                t += hs
            Js = None
            if want_J and steps > 0:
                part = self._view(_lib.FDS_BUF_JSUM, 0, steps * 2).clone()
                Js = self.ring.ordered_sum(part).reshape(steps, 2)
        return None if Js is None else Js[..., :self.n_qoi] / self.j_div

    # -- one boundary ----------------------------------------------------------- #
    def boundary(self, s, eps, from_ensemble, do_qr, build, keep_tangents=True):
        """Time dilation + projection (+ QR) at the current state; optionally builds the
        next perturbed ensemble.  ``keep_tangents=False`` (the hot loop) skips storing the
        tangents explicitly when the ensemble is built.  Returns (R, b, G_dil, g_dil)."""
        c = self.ctx
        build = (_lib.FDS_BUILD_ENSEMBLE if build else 0) | (_lib.FDS_BUILD_KEEP_TANGENTS if keep_tangents else 0)
        if self.in_library:
            if self.run_ddt is not None:
                self._dilation_input(s)
            return c.boundary(s, eps, from_ensemble, do_qr, build, 0)
        c.boundary_begin(from_ensemble)
        self._dilation_input(s)
        c.boundary_dxdt(s, from_ensemble, eps)
        self._reduce_gram()
        c.boundary_factor(do_qr, from_ensemble, eps)
        if do_qr and not c.single_pass_pending:      # the certified single pass launched no second Gram
            self._reduce_gram()
        c.boundary_finish(do_qr, eps, build)
        return c.boundary_results(do_qr)

    # -- the loop without host round trips ---------------------------------------- #
    def deferred_ok(self):
        """Can boundaries and segments be enqueued back to back, their KB-sized results fetched at the end?"""
        return self.in_library and self.run_ddt is None

    def begin_deferred(self, nseg, steps):
        """Page-locked landing area for ``nseg`` boundaries and ``nseg`` segments of ``steps`` steps."""
        L = self.ctx.boundary_out_doubles * max(1, self.batch)
        self._dq = dict(out=_lib.pinned_empty((int(nseg), L)),
                        J=_lib.pinned_empty((int(nseg), int(steps)) + self.ctx.lead + (self.M, 2)),
                        nb=0, ns=0, steps=int(steps))

    def boundary_deferred(self, s, eps, from_ensemble, do_qr, build, keep_tangents=True):
        q = self._dq
        build = (_lib.FDS_BUILD_ENSEMBLE if build else 0) | (_lib.FDS_BUILD_KEEP_TANGENTS if keep_tangents else 0)
        assert do_qr
        self.ctx.boundary_enqueue(s, eps, from_ensemble, do_qr, build, 0, q['out'][q['nb']])
        q['nb'] += 1

    def segment_deferred(self, s, eps, steps):
        q = self._dq
        assert int(steps) == q['steps']
        self.ctx.segment_enqueue(s, eps, steps, q['J'][q['ns']])
        q['ns'] += 1

    def end_deferred(self):
        """Wait for the stream; returns ([(R, b, G_dil, g_dil)] per boundary, [J (steps, m+2, n_qoi)] per segment)."""
        q, self._dq = self._dq, None
        self.ctx.sync()
        bnd = [self.ctx.boundary_unpack(q['out'][i], 1)[:4] for i in range(q['nb'])]
        Js = [q['J'][i][..., :self.n_qoi] / self.j_div for i in range(q['ns'])]
        return bnd, Js

    # -- one segment ------------------------------------------------------------ #
    def segment(self, s, eps, steps):
        """Advance the m+2 trajectories; returns J (steps, m+2, 2)."""
        steps = int(steps)
        if self.in_library:
            return self.ctx.segment(s, eps, steps)[..., :self.n_qoi] / self.j_div
        t = 0
        while t < steps:
            hs = self.ctx.halo_steps(steps - t)
            self._halo(_lib.FDS_BUF_ENSEMBLE, hs)
            self.ctx.segment_advance(s, eps, t, hs)
            t += hs
        part = self._view(_lib.FDS_BUF_JSUM, 0, steps * self.M * 2).clone()
        Js = self.ring.ordered_sum(part).reshape(steps, self.M, 2)
        return Js[..., :self.n_qoi] / self.j_div

    def segment_sens(self, s, eps, steps):
        """Advance one segment and return (J0, G, g) reduced on the device (single rank / batch):
        the (steps, m+2, n_qoi) objective block never crosses PCIe."""
        assert self.ring is None
        steps = int(steps)
        self.ctx.segment_enqueue(s, eps, steps, None)
        return self.ctx.segment_sensitivities(steps, eps)

    # -- state in / out (global host arrays) ------------------------------------ #
    def set_state(self, u):
        self.ctx.set_state(self.local(u))

    def get_state(self):
        return self.gather_sites(self.ctx.get_state())

    def set_tangents(self, V, v):
        self.ctx.set_tangents(self.local(V), None if v is None else self.local(v))

    def get_tangents(self):
        V, v = self.ctx.get_tangents()
        return self.gather_sites(V), self.gather_sites(v)

    def orthonormal_random_tangents(self, seed=None, host_rng=None, block=None):
        """V = Q of the thin QR of a uniform [0,1) (m, N) block, v = 0
        (reference fds/fds.py:21-27).  ``host_rng`` draws the block on the host with
        numpy (small N; honours np.random.seed like the reference); otherwise a
        counter-based device generator keyed on the global (row, site) index."""
        if block is not None:
            self.ctx.set_tangents(self.local(block), None)
        elif host_rng is not None:
            W = host_rng.rand(*(self.ctx.lead + (self.m, self.n_global)))
            self.ctx.set_tangents(self.local(W), None)
        else:
            self.ctx.random_tangents(seed)
        td = getattr(self, 'dil_mode', 1)
        self.ctx.set_time_dilation(0)
        try:
            c = self.ctx
            if self.in_library:
                c.boundary(0.0, 1.0, 0, 1, 0, 0)
            else:
                c.boundary_begin(0)
                c.boundary_dxdt(0.0, 0, 1.0)
                self._reduce_gram()
                c.boundary_factor(1, 0, 1.0)
                if not c.single_pass_pending:
                    self._reduce_gram()
                c.boundary_finish(1, 1.0, 0)
        finally:
            self.ctx.set_time_dilation(td)

    def close(self):
        self.ctx.close()


class _FieldFile:
    """A state the solver returned as a file: its name and the ``/field`` read from it."""
    def __init__(self, path, data):
        self.path, self.data = path, data


class HostSolverEngine(Engine):
    """Engine for a solver that lives on the host (reference SURVEY 8f-4): the user's own
    ``run(u0, parameter, steps[, run_id][, interprocess])`` advances the m+2 trajectories, everything
    else of the segment loop -- time-dilation algebra, projection, Gram, CholeskyQR, ``R``, ``b``,
    ``G_dil``, the next perturbed ensemble, the LSS solve -- runs on the device.  Per segment the
    ensemble crosses PCIe once in each direction, as it would cross the reference's process pool.
    The run_id strings are the reference's (fds/segment.py:38-49, fds/fds.py:109,131).

    File mode (the reference's MPI mode, fds/compute.py:59-132): when the initial state is the NAME of an HDF5 file
    (or ``field_files=True``), states travel to and from the solver as files holding one dataset ``/field``
    (``h5field.py``): the baseline run gets the file the previous baseline run returned, every perturbed
    trajectory its own ``<get_host_dir(run_id)>/input.h5`` (fds/segment.py:46,51), and ``run`` answers with
    ``(name of the file it wrote, J)`` (tests/test_fun3d_mpi.py:82-122)."""

    def __init__(self, run, m, math='exact', device=None, simultaneous_runs=None, field_files=False):
        import threading
        self.user_run = run
        self.field_files = bool(field_files)
        self.get_host_dir = None
        self._state_path, self._field_shape = None, None
        self.m, self.M = int(m), int(m) + 2
        self.kw = dict(math=math, device=0 if device is None else int(device))
        self.device = self.kw['device']
        self.set_simultaneous_runs(simultaneous_runs)
        # what the reference hands to every run: (Manager().Lock(), Manager().dict()) (fds/fds.py:203-204,
        # used by apps/vida.py and apps/openfoam4_pitzdaily/pisofoam.py to allocate nodes).  The runs here are
        # threads of this process, so a thread lock and a plain dict give the same semantics.
        self.interprocess = (threading.Lock(), {})
        self.ctx = None
        self.ring = None
        self.lib_comm, self.in_library, self._dq = False, True, None
        self.batch = 1
        self.time_dilation = True
        self.dil_mode, self.run_ddt = 1, None
        self.j_div, self.n_qoi = 1.0, None
        self.i_segment = 0
        self._weights = None

    def set_simultaneous_runs(self, n):
        """``simultaneous_runs`` of the reference (fds/segment.py:37): None = as many workers as cores."""
        import os
        self.simultaneous_runs = max(1, int(n)) if n else (os.cpu_count() or 1)

    def deferred_ok(self):
        return False

    # -- the reference's RunWrapper (fds/fds.py:53-90): 3-, 4- or 5-argument plugins -------------- #
    def _call(self, u0, parameter, steps, run_id):
        run, last, ip = self.user_run, None, self.interprocess
        if self.field_files and not isinstance(u0, str):
            u0 = self._write_input(u0, run_id)
        for kw in (dict(run_id=run_id, interprocess=ip), dict(run_id=run_id), dict(interprocess=ip), {}):
            try:
                u1, J = run(u0 if isinstance(u0, str) else np.array(u0), parameter, steps, **kw)
            except TypeError as e:
                last = e
                continue
            J = np.array(J, dtype=np.float64).reshape([steps, -1])
            if isinstance(u1, str):
                return _FieldFile(u1, self._read_field(u1)), J
            return np.asarray(u1, dtype=np.float64), J
        raise last

    # -- file mode: /field datasets in place of arrays (fds/compute.py:91-106) ------------------------- #
    def set_host_dir(self, get_host_dir):
        """``get_host_dir(run_id)`` of the reference (fds/segment.py:32-33): where the input of a run is put."""
        self.get_host_dir = get_host_dir

    def _read_field(self, path):
        from .h5field import read_field
        a = read_field(path)
        if self._field_shape is None:
            self._field_shape = a.shape
        return np.asarray(a, dtype=np.float64).ravel()

    def _write_input(self, u, run_id):
        from .h5field import write_field
        import os
        d = self.get_host_dir(run_id) if self.get_host_dir is not None else run_id
        os.makedirs(d, exist_ok=True)                     # fds/compute.py:123-125
        path = os.path.join(d, 'input.h5')
        u = np.asarray(u, dtype=np.float64)
        write_field(path, u.reshape(self._field_shape) if self._field_shape is not None else u)
        return path

    def _baseline(self):
        """What the baseline run starts from: the file that holds the device's primal state when there is one
        (the reference passes ``u0.field``, the name the previous run returned), else the state itself."""
        return self._state_path if self._state_path is not None else self.ctx.get_state()

    def _ran_baseline(self, u1):
        self._state_path = u1.path if isinstance(u1, _FieldFile) else None
        return u1.data if isinstance(u1, _FieldFile) else u1

    def _map(self, jobs):
        if self.simultaneous_runs == 1 or len(jobs) == 1:
            return [self._call(*j) for j in jobs]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(self.simultaneous_runs) as pool:
            return list(pool.map(lambda j: self._call(*j), jobs))

    def _ensure(self, n):
        if self.ctx is None:
            self.n_global = int(n)
            self.lo, self.hi = 0, self.n_global
            self.ctx = Context(self.n_global, self.m, scheme='host', **self.kw)
            self.ctx.set_time_dilation(self.dil_mode)
            if self._weights is not None:
                self.ctx.set_dxdt_weights(self._weights)
        elif int(n) != self.n_global:
            raise ValueError('state dimension changed from %d to %d' % (self.n_global, n))

    def set_dxdt_weights(self, w):
        self._weights = np.asarray(w, dtype=np.float64)
        if self.ctx is not None:
            self.ctx.set_dxdt_weights(self._weights)

    def set_state(self, u):
        self._state_path = None
        if isinstance(u, str):                            # the reference's file mode: u0 is a file name
            self.field_files, path = True, u
            u = self._read_field(path)
            self._state_path = path
        u = np.asarray(u, dtype=np.float64).ravel()
        self._ensure(u.size)
        self.ctx.set_state(u)

    def run(self, s, steps, want_J=True):
        u1, J = self._call(self._baseline(), s, int(steps), 'runup')           # fds/fds.py:207
        self.ctx.set_state(self._ran_baseline(u1))
        return J if want_J else None

    def boundary(self, s, eps, from_ensemble, do_qr, build, keep_tangents=True):
        c = self.ctx
        flags = (_lib.FDS_BUILD_ENSEMBLE if build else 0) | (_lib.FDS_BUILD_KEEP_TANGENTS if keep_tangents else 0)
        c.boundary_begin(from_ensemble)
        if self.run_ddt is not None:
            c.set_dxdt(np.asarray(self.run_ddt(c.get_state(), s), dtype=np.float64))
        elif self.time_dilation:
            u0 = self._baseline()
            rid = 'time_dilation_{0:02d}'.format(self.i_segment)
            res = self._map([(u0, s, k, rid + '_{0}steps'.format(k)) for k in (1, 2, 3)])
            for k, (uk, _) in zip((1, 2, 3), res):
                c.set_primal_history(k, uk.data if isinstance(uk, _FieldFile) else uk)
        c.boundary_dxdt(s, from_ensemble, eps)
        c.boundary_factor(do_qr, from_ensemble, eps)
        c.boundary_finish(do_qr, eps, flags)
        return c.boundary_results(do_qr)

    def segment(self, s, eps, steps):
        """The m+2 solver runs of fds/segment.py:56-64 on the host; returns J (steps, m+2, n_qoi)."""
        c, m, i = self.ctx, self.m, self.i_segment
        E = c.get_ensemble()
        ids = ['segment{0:02d}_baseline'.format(i)] + \
              ['segment{0:02d}_init_perturb{1:03d}'.format(i, j) for j in range(m)] + \
              ['segment{0:02d}_param_perturb{1:03d}'.format(i, m)]
        jobs = [(E[k], s + eps if k == m + 1 else s, int(steps), ids[k]) for k in range(m + 2)]
        if self._state_path is not None:
            jobs[0] = (self._state_path,) + jobs[0][1:]
        res = self._map(jobs)
        for k, (uk, _) in enumerate(res):
            E[k] = uk.data if isinstance(uk, _FieldFile) else uk
        self._ran_baseline(res[0][0])
        c.set_ensemble(E, eps)
        self.i_segment += 1
        J = np.stack([r[1] for r in res], axis=1)
        self.n_qoi = J.shape[2]
        return J

    def orthonormal_random_tangents(self, seed=None, host_rng=None, block=None):
        if block is None and host_rng is None and seed is None:
            host_rng = np.random
        return Engine.orthonormal_random_tangents(self, seed=seed, host_rng=host_rng, block=block)
