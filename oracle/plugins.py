"""numpy solver plugins obeying the reference's ``run`` contract.  TEST INFRASTRUCTURE.

Contract (reference fds/fds.py:184-200): ``u1, J = run(u0, parameter, steps)``
with ``u0`` a flat float64 array, ``J`` of shape (steps, n_qoi); the state is
advanced one step and THEN the objective recorded (reference
tests/solvers/common/solver.f03:32-35).

The Lorenz-96 plugins define -- once -- the floating-point operation order that
the CUDA kernels in fds_b200/csrc mirror instruction for instruction (no FMA
contraction in exact mode), so the device state is bit-identical to these
functions.  All module-level functions are picklable, as the reference driver
requires (it calls ``run`` from a multiprocessing.Pool, fds/segment.py:56-64).
"""
import numpy as np

L96_DT = 0.01          # reference tests/solvers/mock_fun3d/fun3d:49
L96_F_OFFSET = 8.0     # F = xmach + 8, reference tests/solvers/mock_fun3d/fun3d:48


# --------------------------------------------------------------------------- #
# specified reduction tree for the objective sums
# --------------------------------------------------------------------------- #
def tree_sum(x):
    """Sum over the LAST axis in a specified order: a balanced binary tree over
    adjacent elements (x[0]+x[1], x[2]+x[3], ...), the length padded with zeros
    to the next power of two.  Adding 0.0 is exact, so padding never changes a
    partial.  Any aligned power-of-two block of indices is a sub-tree, which is
    what lets the CUDA kernels (and any N-sharding across GPUs) reproduce the
    result bit for bit."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    p = 1
    while p < n:
        p *= 2
    if p != n:
        pad = np.zeros(x.shape[:-1] + (p - n,))
        x = np.concatenate([x, pad], axis=-1)
    while x.shape[-1] > 1:
        x = x[..., 0::2] + x[..., 1::2]
    return x[..., 0]


def l96_objective(x):
    """J = [mean x, mean x^2] ("lift", "drag" of reference
    tests/solvers/mock_fun3d/fun3d:39-46), sums in tree_sum order."""
    n = x.shape[-1]
    return np.stack([tree_sum(x) / n, tree_sum(x * x) / n], axis=-1)


# --------------------------------------------------------------------------- #
# Lorenz-96 right-hand sides (periodic ring; works on (..., N) arrays)
# --------------------------------------------------------------------------- #
def l96_rhs_factored(x, F):
    """dx_i = (x_{i+1} - x_{i-2}) * x_{i-1} - x_i + F, in exactly that order
    (reference tests/solvers/Lorenz96/Lorenz96.f90:20, "(-X(i-2) + X(i+1))*X(i-1)
    - X(i) + M"; -a+b and b-a are the same IEEE operation)."""
    xm2 = np.roll(x, 2, axis=-1)
    xm1 = np.roll(x, 1, axis=-1)
    xp1 = np.roll(x, -1, axis=-1)
    return (xp1 - xm2) * xm1 - x + F


def l96_rhs_fun3d(x, F):
    """dx_i = -x_{i-2} x_{i-1} + x_{i-1} x_{i+1} - x_i + F, in the order of the
    reference's own mock solver (tests/solvers/mock_fun3d/fun3d:54)."""
    xm2 = np.roll(x, 2, axis=-1)
    xm1 = np.roll(x, 1, axis=-1)
    xp1 = np.roll(x, -1, axis=-1)
    return -xm2 * xm1 + xm1 * xp1 - x + F


def l96_rk4_step(x, F, dt=L96_DT):
    """Classical RK4 with the factored RHS.  Operation order (mirrored by
    fds_b200/csrc/l96_core.cuh):
        y2 = x + (0.5*dt)*k1 ; y3 = x + (0.5*dt)*k2 ; y4 = x + dt*k3
        x' = x + (dt/6)*(((k1 + 2*k2) + 2*k3) + k4)
    """
    h2 = 0.5 * dt
    h6 = dt / 6.0
    k1 = l96_rhs_factored(x, F)
    k2 = l96_rhs_factored(x + h2 * k1, F)
    k3 = l96_rhs_factored(x + h2 * k2, F)
    k4 = l96_rhs_factored(x + dt * k3, F)
    return x + h6 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def l96_euler_step(x, F, dt=L96_DT):
    """Forward Euler, reference tests/solvers/mock_fun3d/fun3d:53-55
    (``x += dt * dxdt``)."""
    return x + dt * l96_rhs_fun3d(x, F)


def _run_l96(step, u0, s, steps):
    x = np.array(u0, dtype=np.float64)
    F = s + L96_F_OFFSET
    J = np.empty((int(steps), 2))
    for t in range(int(steps)):
        x = step(x, F)
        J[t] = l96_objective(x)
    return x, J


def run_l96_rk4(u0, s, steps):
    """Lorenz-96, RK4, dt = 0.01, F = s + 8 (north-star plugin)."""
    return _run_l96(l96_rk4_step, u0, s, steps)


def run_l96_euler(u0, s, steps):
    """Lorenz-96, forward Euler: the reference's mock FUN3D solver
    (tests/solvers/mock_fun3d/fun3d:29-57) without MPI."""
    return _run_l96(l96_euler_step, u0, s, steps)


def run_l96_ensemble(step, U, F, steps):
    """Vectorised helper: advance an (M, N) ensemble with per-row forcing F (M,).
    Row-wise identical to M separate ``run`` calls (elementwise ops only)."""
    x = np.array(U, dtype=np.float64)
    Fc = np.asarray(F, dtype=np.float64).reshape(-1, 1)
    J = np.empty((int(steps), x.shape[0], 2))
    for t in range(int(steps)):
        x = step(x, Fc)
        J[t] = l96_objective(x)
    return x, J


# --------------------------------------------------------------------------- #
# toy ODEs of the reference test-suite (Fortran restated; forward Euler)
# NB the Fortran DT literals are default-REAL, i.e. float32 values widened.
# --------------------------------------------------------------------------- #
_DT_1E3 = float(np.float32(0.001))   # lorenz.f03:4, circular.f03:15
_DT_1E2 = float(np.float32(0.01))    # vanderpol.f03:4


def run_lorenz63(u0, s, steps):
    """reference tests/solvers/lorenz/lorenz.f03:10-26 driven as in
    tests/test_lorenz.py:18-31: params (10, s, 8/3), J = [(z-28)^2, 100]."""
    x = np.array(u0, dtype=np.float64)
    s1, s2, s3 = 10.0, float(s), 8. / 3
    J = np.empty((int(steps), 2))
    for t in range(int(steps)):
        dx0 = s1 * (x[1] - x[0])
        dx1 = x[0] * (s2 - x[2]) - x[1]
        dx2 = x[0] * x[1] - s3 * x[2]
        x = x + _DT_1E3 * np.array([dx0, dx1, dx2])
        J[t, 0] = (x[2] - 28) ** 2
        J[t, 1] = 100.0
    return x, J


def run_vanderpol(u0, s, steps):
    """reference tests/solvers/vanderpol/vanderpol.f03:9-24; J = x1^2."""
    x0, x1 = float(u0[0]), float(u0[1])
    s = float(s)
    J = np.empty((int(steps), 1))
    for t in range(int(steps)):
        dx0 = x1
        dx1 = (s / 4) * (1 - x0 ** 2) * x1 - x0
        x0, x1 = x0 + _DT_1E2 * dx0, x1 + _DT_1E2 * dx1
        J[t, 0] = x0 ** 2
    return np.array([x0, x1]), J


def run_circular(u0, s, steps):
    """reference tests/solvers/circular/circular.f03:9-28; J = x1^2 + x2^2."""
    x0, x1 = float(u0[0]), float(u0[1])
    s = float(s)
    J = np.empty((int(steps), 1))
    for t in range(int(steps)):
        r = x0 * x0 + x1 * x1 - s - 1
        dx0 = +x1 - r * x0
        dx1 = -x0 - r * x1
        x0, x1 = x0 + _DT_1E3 * dx0, x1 + _DT_1E3 * dx1
        J[t, 0] = x0 * x0 + x1 * x1
    return np.array([x0, x1]), J
