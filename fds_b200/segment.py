"""One time segment (reference ``fds/segment.py``).

``trapez_mean`` is the reference's time-mean (:9-12).  ``run_segment`` is the
host-buffer form of the reference call (:14-82) for a device plugin: the m+2
perturbed trajectories are formed, advanced and differenced on the GPU.
"""
import numpy as np


def trapez_mean(J, dim):
    """Trapezoid time mean with the extrapolated sample J[-1] = 2 J[0] - J[1] in front
    (reference fds/segment.py:9-12; same operation order, so identical bits)."""
    J = np.moveaxis(np.asarray(J), dim, 0)
    ghost = 2 * J[0] - J[1]
    return (J.sum(0) + J[:-1].sum(0) + ghost) / (2 * J.shape[0])


def sensitivities(J, epsilon):
    """J (steps, m+2, n_qoi) of one segment -> (J0, G (m, n_qoi), g (n_qoi,))
    (reference fds/segment.py:75,79)."""
    J0 = np.ascontiguousarray(J[:, 0])
    m = J.shape[1] - 2
    G = [trapez_mean(J[:, 1 + j] - J0, 0) / epsilon for j in range(m)]
    g = trapez_mean(J[:, m + 1] - J0, 0) / epsilon
    return J0, G, g


def run_segment(run, u0, V, v, parameter, i_segment, steps, epsilon,
                simultaneous_runs=None, interprocess=None, get_host_dir=None,
                compute_outputs=None, spawn_compute_job=None):
    """u0 (N,), V (m, N), v (N,) host arrays -> (u0', V', v', J0, G, g); the pool /
    host-dir / compute-job arguments of the reference are accepted and inert."""
    V = np.asarray(V)
    eng = run.engine(V.shape[0])
    u1, V1, v1, Js = eng.ctx.run_segment_host(u0, V, v, parameter, epsilon, steps)
    J0, G, g = sensitivities(Js[:, :, :eng.n_qoi] / eng.j_div, epsilon)
    return u1, V1, v1, J0, G, g
