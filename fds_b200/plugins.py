"""Built-in device solver plugins obeying the reference's ``run`` contract.

``u1, J = run(u0, parameter, steps[, run_id][, interprocess])`` (reference
fds/fds.py:184-200): ``u0`` flat float64 (N,), ``J`` (steps, n_qoi), the state is
advanced one step and then the objective recorded.  An ``L96`` instance is such a
callable (host numpy in / out, computed on the GPU), so the REFERENCE driver can run
it unchanged; handed to ``fds_b200.shadowing`` it is recognised as device-resident
and the whole segment loop stays on the GPU.
"""
import numpy as np

from .engine import Engine, HostSolverEngine


def _current_device():
    """The rank's GPU when running under torch.distributed (one process per GPU), else 0."""
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and torch.cuda.is_available():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


class _DevicePlugin:
    """Common part of the device solver plugins: engines cached per subspace dimension, the
    reference ``run`` call on host arrays, picklability (the reference's pool, fds/segment.py:37)."""
    n_qoi = 2

    def _init(self, N, **kw):
        self.N = int(N)
        self.kw = kw
        self._engines = {}

    def engine(self, m):
        """The device engine holding an (m+2)-trajectory ensemble (cached per m)."""
        if m not in self._engines:
            self._engines[m] = Engine(self.N, m, **self.kw)
        return self._engines[m]

    def batch_engine(self, m, batch):
        """Engine holding ``batch`` independent problems of this system (parameter sweep)."""
        key = (m, int(batch))
        if key not in self._engines:
            kw = dict(self.kw)
            kw.pop('group', None), kw.pop('distributed', None), kw.pop('comm', None)
            if kw.get('device') is None:
                kw['device'] = _current_device()
            self._engines[key] = Engine(self.N, m, batch=int(batch), **kw)
        return self._engines[key]

    def __call__(self, u0, parameter, steps, run_id=None, interprocess=None):
        eng = self.engine(1)
        eng.set_state(u0)
        J = eng.run(parameter, steps, want_J=True)
        if J is None:
            J = np.empty((0, self.n_qoi))
        return eng.get_state(), J

    def __getstate__(self):
        return {'N': self.N, 'kw': self.kw, '_engines': {}}

    def close(self):
        for e in self._engines.values():
            e.close()
        self._engines = {}


class HostSolver:
    """Opt-in wrapper for a solver that cannot move to the GPU: ``HostSolver(run)`` hands the user's
    ``run(u0, parameter, steps[, run_id][, interprocess]) -> (u1, J)`` (reference fds/fds.py:184-200) to
    ``shadowing`` / ``continue_shadowing``; the trajectories are advanced by that solver on the host, the
    algebra of the segment loop runs on the device (``HostSolverEngine``).  A bare Python callable is
    still rejected: nothing in this package runs the hot path on the CPU behind the user's back.

    File mode of the reference (fds/compute.py:59-132, tests/test_fun3d_mpi.py): pass the NAME of an HDF5 file
    with one dataset ``/field`` as ``u0`` (or ``field_files=True`` when continuing from a checkpoint) and ``run``
    is called with file names and answers ``(name of the file it wrote, J)``; ``get_host_dir`` of ``shadowing``
    says where the ``input.h5`` of each run goes."""

    def __init__(self, run, math='exact', device=None, simultaneous_runs=None, field_files=False):
        self.user_run = run
        self.kw = dict(math=math, device=device, simultaneous_runs=simultaneous_runs, field_files=field_files)
        self._engines = {}

    def engine(self, m):
        if m not in self._engines:
            self._engines[m] = HostSolverEngine(self.user_run, m, **self.kw)
        return self._engines[m]

    def __call__(self, u0, parameter, steps, run_id=None, interprocess=None):
        return self.user_run(u0, parameter, steps)

    def close(self):
        for e in self._engines.values():
            if e.ctx is not None:
                e.close()
        self._engines = {}


class Lorenz63(_DevicePlugin):
    """Lorenz-63, forward Euler dt = float32(0.001), parameters (10, s, 8/3),
    ``J = [(z - 28)^2, 100]`` -- the reference's tests/solvers/lorenz/lorenz.f03:10-26 as
    driven by tests/test_lorenz.py:18-31 (BASELINE config 1)."""
    n_qoi = 2

    def __init__(self, math='exact', device=None):
        self._init(3, scheme='lorenz63', math=math, device=device)


class VanDerPol(_DevicePlugin):
    """Van der Pol oscillator, forward Euler dt = float32(0.01), ``J = x0^2`` -- the
    reference's tests/solvers/vanderpol/vanderpol.f03:9-24 (BASELINE config 2)."""
    n_qoi = 1

    def __init__(self, math='exact', device=None):
        self._init(2, scheme='vanderpol', math=math, device=device)


class Circular(_DevicePlugin):
    """Limit cycle of radius sqrt(s + 1), forward Euler dt = float32(0.001),
    ``J = x0^2 + x1^2`` -- the reference's tests/solvers/circular/circular.f03:9-28."""
    n_qoi = 1

    def __init__(self, math='exact', device=None):
        self._init(2, scheme='circular', math=math, device=device)


class L96(_DevicePlugin):
    """Lorenz-96 ring, ``dx_i/dt = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F``, ``F = s + 8``,
    ``J = [mean x, mean x^2]`` -- the reference's own Lorenz-96 test solver
    (tests/solvers/mock_fun3d/fun3d:39-57) with ``scheme='euler'``, and the classical
    RK4 integrator of the north-star configuration with ``scheme='rk4'``.

    math='exact': FMA-free, numpy operation order (bit-identical to oracle/plugins.py);
    math='fast' : FMA contraction allowed.
    """
    n_qoi = 2

    def __init__(self, N, scheme='rk4', dt=0.01, math='exact', device=None, group=None,
                 max_halo_steps=0, simple_kernels=False, stream=None, distributed=None, comm='nccl'):
        """Under an initialised ``torch.distributed`` (one process per GPU) the state is sharded along N over
        the ranks of ``group``; ``comm='nccl'`` (default) lets the library run halo exchange and reductions on
        its own NCCL communicator, ``comm='torch'`` drives them from the host (``fds_b200/ring.py``)."""
        self._init(N, scheme=scheme, math=math, dt=dt, device=device, group=group,
                   max_halo_steps=max_halo_steps, simple_kernels=simple_kernels, stream=stream,
                   distributed=distributed, comm=comm)
