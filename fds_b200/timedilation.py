"""Time-dilation finite-difference stencil (host side of reference ``fds/timedilation.py``).

On the device the whole of TimeDilation (three short primal runs, dxdt, contribution,
projection; reference :37-82) happens inside the boundary kernels.  What stays on the
host is the tiny stencil solve whose weights are handed to the device
(fds_set_dxdt_weights) so both sides use bit-identical coefficients, and
``compute_dxdt`` for host arrays (reference tests/test_timedilation.py:13-16).
"""
import sys

import numpy as np

ORDER_OF_ACCURACY = 3     # reference fds/timedilation.py:64


def set_order_of_accuracy(order_of_accuracy):
    """Order of the one-sided difference used for the time-dilation vector: that many look-ahead steps of the
    primal per boundary (reference fds/timedilation.py:8-12, incl. its clamp at 2).  Takes effect at the next
    ``shadowing`` / ``continue_shadowing`` call; Lorenz-96 device plugins support 2 .. 8, the small ODE systems
    and host solvers 3."""
    global ORDER_OF_ACCURACY
    if order_of_accuracy < 2:
        order_of_accuracy = 2
        sys.stderr.write('Order of accuracy too low, setting to 2 instead\n')
    ORDER_OF_ACCURACY = int(order_of_accuracy)


def stencil_weights(order):
    """One-sided first-derivative weights on nodes 0..order, from the same Vandermonde
    solve as the reference (fds/timedilation.py:16-19) so the doubles are identical."""
    assert order >= 1
    nodes = np.arange(order + 1)
    vander = np.array([nodes ** p for p in range(order + 1)])
    rhs = np.zeros(order + 1)
    rhs[1] = 1
    return np.linalg.solve(vander, rhs)


def compute_dxdt_of_order(u, order):
    c = stencil_weights(order)
    acc = 0
    for k in range(order + 1):          # left-to-right accumulation, as Python's sum()
        acc = acc + c[k] * u[k]
    return acc


def compute_dxdt(u):
    """Per-step derivative from consecutive states u[0..p] (step size cancels out)."""
    return compute_dxdt_of_order(u, len(u) - 1)
