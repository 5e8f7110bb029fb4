"""Device context: thin object wrapper over the C ABI (include/fds_b200.h).

One ``Context`` owns the device-resident state of one shard of the shadowing
segment loop (ensemble, tangents, dxdt, objective sums, small matrices).  All
arrays crossing this boundary are host numpy float64; nothing here computes on
the CPU -- every method is one or more calls into libfds_b200.so and raises
``FdsError`` if the library or a CUDA device is missing.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import FdsError, dptr

SYSTEMS = {'rk4': _lib.FDS_SYS_L96_RK4, 'euler': _lib.FDS_SYS_L96_EULER,
           'lorenz63': _lib.FDS_SYS_LORENZ63, 'vanderpol': _lib.FDS_SYS_VANDERPOL,
           'circular': _lib.FDS_SYS_CIRCULAR, 'host': _lib.FDS_SYS_HOST}
# small ODE systems: (state dimension, n_qoi); their objective is recorded directly
TOY_SYSTEMS = {'lorenz63': (3, 2), 'vanderpol': (2, 1), 'circular': (2, 1)}
MATH = {'exact': _lib.FDS_MATH_EXACT, 'fast': _lib.FDS_MATH_FAST}


class Context:
    """Device-resident Lorenz-96 shadowing state for ``n_local`` sites x (m+2) trajectories."""

    def __init__(self, n_local, m, scheme='rk4', math='exact', dt=0.01, device=0,
                 n_global=None, site_offset=0, rank=0, world=1, max_halo_steps=0,
                 simple_kernels=False, stream=None, batch=1):
        self.lib = _lib.load()
        self.batch = int(batch)
        # leading problem axis of every host array of a batched context
        self.lead = (self.batch,) if self.batch > 1 else ()
        self.n = int(n_local)
        self.m = int(m)
        self.M = self.m + 2
        self.n_global = int(n_global if n_global is not None else n_local)
        self.rank, self.world = int(rank), int(world)
        cfg = _lib.FdsConfig(device=device, system=SYSTEMS[scheme], math=MATH[math], m=self.m,
                             n_local=self.n, n_global=self.n_global, site_offset=int(site_offset),
                             rank=self.rank, world=self.world, max_halo_steps=int(max_halo_steps),
                             flags=(_lib.FDS_FLAG_SIMPLE_KERNELS if simple_kernels else 0) |
                                   (_lib.FDS_FLAG_EXTERNAL_STREAM if stream is not None else 0),
                             dt=float(dt), stream=stream, batch=self.batch if self.batch > 1 else 0)
        h = ctypes.c_void_p()
        rc = self.lib.fds_ctx_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise FdsError('fds_ctx_create failed (%d): %s'
                           % (rc, self.lib.fds_last_error(None).decode()))
        self.h = h

    # -- plumbing ------------------------------------------------------------ #
    def _ck(self, rc):
        if rc != 0:
            raise FdsError(self.lib.fds_last_error(self.h).decode())

    def close(self):
        if getattr(self, 'h', None):
            self.lib.fds_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self.lib.fds_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.fds_kernel_launches(self.h))

    def profile_enable(self, enabled=True):
        """Time every launch of the ensemble step kernel with CUDA events on the context's stream."""
        self._ck(self.lib.fds_profile_enable(self.h, int(bool(enabled))))

    @property
    def profile_steps(self):
        """Time steps covered by the profiled launches (2 per launch with temporal blocking)."""
        return int(self.lib.fds_profile_steps(self.h))

    def profile_read(self):
        """(summed device milliseconds, launches) of the step kernel since profile_enable()."""
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        self._ck(self.lib.fds_profile_read(self.h, ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, cnt.value

    def profile_cycles(self):
        """Mean SM cycles a CTA of the profiled step-kernel launches ran (clock64 inside the kernel)."""
        c = ctypes.c_double()
        self._ck(self.lib.fds_profile_cycles(self.h, ctypes.byref(c)))
        return c.value

    def buffer_ptr(self, which):
        return self.lib.fds_buffer_ptr(self.h, which)

    # -- state --------------------------------------------------------------- #
    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == self.lead + (self.n,)
        self._ck(self.lib.fds_set_state(self.h, dptr(u)))

    def get_state(self):
        u = _lib.host_empty(self.lead + (self.n,))
        self._ck(self.lib.fds_get_state(self.h, dptr(u)))
        return u

    def set_tangents(self, V, v=None):
        V = np.ascontiguousarray(V, dtype=np.float64)
        assert V.shape == self.lead + (self.m, self.n)
        if v is not None:
            v = np.ascontiguousarray(v, dtype=np.float64)
            assert v.shape == self.lead + (self.n,)
        self._ck(self.lib.fds_set_tangents(self.h, dptr(V), dptr(v)))

    def get_tangents(self):
        V, v = _lib.host_empty(self.lead + (self.m, self.n)), _lib.host_empty(self.lead + (self.n,))
        self._ck(self.lib.fds_get_tangents(self.h, dptr(V), dptr(v)))
        return V, v

    # -- host-solver mode ------------------------------------------------------ #
    def get_ensemble(self):
        """(m+2, n): row 0 the primal, rows 1..m homogeneous, row m+1 inhomogeneous perturbed states."""
        E = _lib.host_empty((self.M, self.n))
        self._ck(self.lib.fds_get_ensemble(self.h, dptr(E)))
        return E

    def set_ensemble(self, E, eps):
        E = np.ascontiguousarray(E, dtype=np.float64)
        assert E.shape == (self.M, self.n)
        self._ck(self.lib.fds_set_ensemble(self.h, dptr(E), float(eps)))

    def set_primal_history(self, j, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (self.n,)
        self._ck(self.lib.fds_set_primal_history(self.h, int(j), dptr(u)))

    def random_tangents(self, seed):
        self._ck(self.lib.fds_random_tangents(self.h, ctypes.c_uint64(int(seed) & (2 ** 64 - 1))))

    def set_dxdt_weights(self, c):
        """One-sided first-derivative weights on order + 1 nodes (order 3 unless len(c) says otherwise)."""
        c = np.ascontiguousarray(c, dtype=np.float64)
        assert c.ndim == 1 and c.size >= 2
        self._ck(self.lib.fds_set_dxdt_stencil(self.h, c.size - 1, dptr(c)))

    @property
    def dilation_order(self):
        return int(self.lib.fds_dilation_order(self.h))

    def set_time_dilation(self, mode):
        """False / 0: off; True / 1: finite differences; 2: analytic dt*f(u); 3: given (set_dxdt)."""
        self._ck(self.lib.fds_set_time_dilation(self.h, int(mode)))

    def set_dxdt(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        assert d.shape == self.lead + (self.n,)
        self._ck(self.lib.fds_set_dxdt(self.h, dptr(d)))

    def set_qr_passes(self, passes):
        """0: adaptive (one CholeskyQR pass when certified well conditioned), 2: always CholeskyQR2."""
        self._ck(self.lib.fds_set_qr_passes(self.h, int(passes)))

    @property
    def last_condition(self):
        return float(self.lib.fds_last_condition(self.h))

    @property
    def last_qr_passes(self):
        """CholeskyQR passes (1 or 2) the last boundary with QR took."""
        return int(self.lib.fds_last_qr_passes(self.h))

    @property
    def single_pass_pending(self):
        """Split-phase API: the certified single CholeskyQR pass was accepted by boundary_factor."""
        return bool(self.lib.fds_single_pass_pending(self.h))

    def set_batch_params(self, ds):
        """Problem p of a batched context runs at parameter s + ds[p]."""
        ds = np.ascontiguousarray(ds, dtype=np.float64)
        assert ds.shape == (self.batch,)
        self._ck(self.lib.fds_set_batch_params(self.h, dptr(ds)))

    # -- solver plugin ------------------------------------------------------- #
    def run(self, s, steps, want_J=True):
        """Advance the primal ``steps`` steps; returns objective SUMS [steps][2] (or None)."""
        J = np.empty((int(steps), 2)) if want_J and steps > 0 else None
        self._ck(self.lib.fds_run(self.h, float(s), int(steps), dptr(J)))
        return J

    # -- halos / phases (world > 1 drives these itself) ---------------------- #
    def halo_steps(self, steps):
        return int(self.lib.fds_halo_steps(self.h, int(steps)))

    def halo_layout(self, which, hs):
        off, cnt = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
        if self.lib.fds_halo_layout(self.h, which, int(hs), off, cnt) != 0:
            raise FdsError('fds_halo_layout: bad arguments')
        return list(off), list(cnt)

    def halo_wrap_ensemble(self, hs):
        self._ck(self.lib.fds_halo_wrap_ensemble(self.h, int(hs)))

    def halo_wrap_primal(self, hs):
        self._ck(self.lib.fds_halo_wrap_primal(self.h, int(hs)))

    def halo_mark_valid(self, which, hs):
        self._ck(self.lib.fds_halo_mark_valid(self.h, which, int(hs)))

    def primal_advance(self, s, step0, nsteps, record):
        self._ck(self.lib.fds_primal_advance(self.h, float(s), int(step0), int(nsteps), int(record)))

    def boundary_begin(self, from_ensemble):
        self._ck(self.lib.fds_boundary_begin(self.h, int(from_ensemble)))

    def boundary_dxdt(self, s, from_ensemble, eps):
        self._ck(self.lib.fds_boundary_dxdt(self.h, float(s), int(from_ensemble), float(eps)))

    def boundary_factor(self, do_qr, from_ensemble, eps):
        self._ck(self.lib.fds_boundary_factor(self.h, int(do_qr), int(from_ensemble), float(eps)))

    def boundary_finish(self, do_qr, eps, build):
        self._ck(self.lib.fds_boundary_finish(self.h, int(do_qr), float(eps), int(build)))

    def boundary_results(self, do_qr):
        m, L = self.m, self.lead
        R = np.empty(L + (m, m)) if do_qr else None
        b = np.empty(L + (m,)) if do_qr else None
        Gd, gd = np.empty(L + (m,)), np.empty(L + (1,))
        self._ck(self.lib.fds_boundary_results(self.h, dptr(R), dptr(b), dptr(Gd), dptr(gd)))
        return R, b, Gd, (gd[:, 0] if L else float(gd[0]))

    def segment_advance(self, s, eps, step0, nsteps):
        self._ck(self.lib.fds_segment_advance(self.h, float(s), float(eps), int(step0), int(nsteps)))

    def segment_jsums(self, steps):
        J = np.empty((int(steps),) + self.lead + (self.M, 2))
        self._ck(self.lib.fds_segment_jsums(self.h, int(steps), dptr(J)))
        return J

    def segment_sensitivities(self, steps, eps):
        """(J0 lead+(steps, 2), G lead+(m, 2), g lead+(2,)) of the last segment, reduced on the device."""
        J0 = np.empty(self.lead + (int(steps), 2))
        G = np.empty(self.lead + (self.m + 1, 2))
        self._ck(self.lib.fds_segment_sensitivities(self.h, int(steps), float(eps), dptr(J0), dptr(G)))
        return J0, G[..., :self.m, :], G[..., self.m, :]

    # -- whole boundary / segment in one call (world == 1, or a communicator attached) -------- #
    def boundary(self, s, eps, from_ensemble, do_qr, build, halo_steps=0):
        """One segment boundary; returns (R, b, G_dil, g_dil) (R, b None unless do_qr)."""
        m, L = self.m, self.lead
        R = np.empty(L + (m, m)) if do_qr else None
        b = np.empty(L + (m,)) if do_qr else None
        Gd, gd = np.empty(L + (m,)), np.empty(L + (1,))
        self._ck(self.lib.fds_boundary(self.h, float(s), float(eps), int(from_ensemble), int(do_qr),
                                       int(build), int(halo_steps), dptr(R), dptr(b), dptr(Gd), dptr(gd)))
        return R, b, Gd, (gd[:, 0] if L else float(gd[0]))

    # -- the same without waiting for the device ------------------------------------------------ #
    @property
    def boundary_out_doubles(self):
        return int(self.lib.fds_boundary_out_doubles(self.h))

    def boundary_enqueue(self, s, eps, from_ensemble, do_qr, build, halo_steps, out):
        """Put one boundary on the stream; ``out`` (page-locked, batch x boundary_out_doubles) is filled
        when the stream gets there.  Decode with ``boundary_unpack`` after ``sync()``."""
        assert out is None or (out.size == max(1, self.batch) * self.boundary_out_doubles and out.flags['C_CONTIGUOUS'])
        self._ck(self.lib.fds_boundary_enqueue(self.h, float(s), float(eps), int(from_ensemble), int(do_qr),
                                               int(build), int(halo_steps), dptr(out)))

    def boundary_unpack(self, out, do_qr):
        """(R, b, G_dil, g_dil, two_passes, cond) from a filled ``out`` block; raises on a failed Cholesky pivot."""
        self._ck(self.lib.fds_boundary_check(self.h, dptr(out), int(do_qr)))
        m, nb = self.m, max(1, self.batch)
        blk = out.reshape(nb, -1)
        R = blk[:, :m * m].reshape(self.lead + (m, m)).copy() if do_qr else None
        b = blk[:, m * m:m * m + m].reshape(self.lead + (m,)).copy() if do_qr else None
        Gd = blk[:, m * m + m:m * m + 2 * m].reshape(self.lead + (m,)).copy()
        gd = blk[:, m * m + 2 * m].copy()
        st = blk[:, m * m + 2 * m + 1].copy().view(np.int32).reshape(nb, 2)
        cond = blk[:, m * m + 2 * m + 2].copy()
        return R, b, Gd, (gd if self.lead else float(gd[0])), bool(st[:, 1].any()), float(cond.max())

    def segment_enqueue(self, s, eps, steps, Jout):
        """Put one segment on the stream; ``Jout`` (page-locked, (steps,) + lead + (m+2, 2)) gets the objective sums."""
        assert Jout is None or (Jout.shape == (int(steps),) + self.lead + (self.M, 2) and Jout.flags['C_CONTIGUOUS'])
        self._ck(self.lib.fds_segment_enqueue(self.h, float(s), float(eps), int(steps), dptr(Jout)))

    # -- in-library communicator ------------------------------------------------------------------ #
    def comm_init(self, unique_id):
        """Join the NCCL communicator identified by the 128 bytes rank 0 got from ``comm_unique_id()``."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.fds_comm_init(self.h, buf))

    @property
    def comm_attached(self):
        return bool(self.lib.fds_comm_attached(self.h))

    @property
    def collectives(self):
        return int(self.lib.fds_collectives(self.h))

    # -- introspection ---------------------------------------------------------------------------- #
    def ensemble_rows(self, site0, nsites):
        """Rows [site0, site0 + nsites) of the current ensemble, (nsites, m+2); ghost sites allowed."""
        out = np.empty((int(nsites), self.M))
        self._ck(self.lib.fds_get_ensemble_rows(self.h, int(site0), int(nsites), dptr(out)))
        return out

    def ensemble_column(self, k):
        out = _lib.host_empty((self.n,))
        self._ck(self.lib.fds_get_ensemble_column(self.h, int(k), dptr(out)))
        return out

    def strip_plan(self):
        out = (ctypes.c_int64 * 8)()
        if self.lib.fds_strip_plan(self.h, out) != 0:
            raise FdsError('fds_strip_plan: this context does not use the streaming step kernel')
        keys = ('a0', 'strip_sites', 'strips', 'blocks_per_strip', 'node_shift', 'nodes', 'lanes_per_cta', 'steps_per_pass')
        return dict(zip(keys, (int(x) for x in out)))

    def segment(self, s, eps, steps, want_J=True):
        """Advance all m+2 trajectories ``steps`` steps; objective SUMS [steps][m+2][2]."""
        J = np.empty((int(steps),) + self.lead + (self.M, 2)) if want_J else None
        self._ck(self.lib.fds_segment(self.h, float(s), float(eps), int(steps), dptr(J)))
        return J

    def run_segment_host(self, u, V, v, s, eps, steps):
        """Host-buffer form of the reference's run_segment (fds/segment.py:14-82): u, V, v are read
        where they lie (at PCIe rate if page-locked), the advanced u, V, v come back in pinned arrays.
        (The C entry point fds_run_segment_host does the same in place on caller-owned buffers.)"""
        self.set_state(u)
        self.set_tangents(V, v)
        self.boundary_finish(0, eps, _lib.FDS_BUILD_ENSEMBLE | _lib.FDS_BUILD_KEEP_TANGENTS)   # u + eps [V; v]
        J = self.segment(s, eps, steps)
        V1, v1 = self.get_tangents()
        return self.get_state(), V1, v1, J


def comm_unique_id():
    """128-byte NCCL id (call on ONE rank, distribute to the others, then ``Context.comm_init``)."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(128)
    if lib.fds_comm_unique_id(buf) != 0:
        raise FdsError(lib.fds_last_error(None).decode())
    return buf.raw


def fp64_issue_rate(device=0):
    """(FP64 warp-instructions per second of the whole GPU, seconds the probe ran, FP64 warp-instructions per SM
    and SM clock): the measured roof of the exact-mode step kernel (same CTA shape and instruction mix, no memory
    traffic); the third value comes from the probe's own cycle counters and does not depend on the clock."""
    lib = _lib.load()
    r, t, pc = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    if lib.fds_fp64_issue_rate(int(device), ctypes.byref(r), ctypes.byref(t), ctypes.byref(pc)) != 0:
        raise FdsError(lib.fds_last_error(None).decode())
    return r.value, t.value, pc.value


def lss_solve(R, b, device=0):
    """LssTangent.solve on the device (reference fds/lsstan.py:36-50).  R (K, m, m), b (K, m)
    -> alpha (K, m); with a leading batch axis on both, one CTA per problem."""
    lib = _lib.load()
    R = np.ascontiguousarray(R, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    K, m = b.shape[-2:]
    assert R.shape == b.shape[:-1] + (m, m) and b.ndim in (2, 3)
    alpha = np.empty(b.shape)
    nbatch = b.shape[0] if b.ndim == 3 else 1
    rc = lib.fds_lss_solve_batch(int(device), dptr(R), dptr(b), K, m, nbatch, dptr(alpha))
    if rc != 0:
        raise FdsError(lib.fds_last_error(None).decode())
    return alpha
