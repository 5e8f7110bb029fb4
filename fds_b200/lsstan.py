"""Least-squares-shadowing tangent bookkeeping (reference ``fds/lsstan.py``).

``LssTangent`` stores the per-boundary ``R_i`` (m x m) and ``b_i`` (m) that the
device produces (CholeskyQR2, fds_b200/csrc/boundary.cu) and solves the
block-bidiagonal minimum-norm problem on the device (fds_b200/csrc/lss.cu).
R has a positive diagonal here (the reference's serial path returns LAPACK's
signs, its MPI path normalises to positive, pascal_lite/operators/plinalg.py:62-64);
dJ/ds and the Lyapunov exponents do not depend on that choice.
"""
import numpy as np

from . import device


def _solve_device():
    """GPU of the KKT solve: this process's current device under torch.distributed (one process per
    GPU), else device 0.  Resolved at call time, never stored: a checkpoint written by rank 5 of an
    8-GPU run must load and solve on a one-GPU box."""
    try:
        import torch
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


class LssTangent:
    def __init__(self, Rs=None, bs=None, device_ordinal=None):
        self.Rs = list(Rs) if Rs is not None else []
        self.bs = list(bs) if bs is not None else []
        self.device_ordinal = device_ordinal      # explicit override for this process only (not pickled)

    def __getstate__(self):
        return {'Rs': self.Rs, 'bs': self.bs}

    def __setstate__(self, st):
        self.Rs, self.bs = list(st['Rs']), list(st['bs'])
        self.device_ordinal = None

    def K_segments(self):
        assert len(self.Rs) == len(self.bs)
        return len(self.Rs)

    def m_modes(self):
        return self.Rs[0].shape[0]

    def append(self, R, b):
        """Record one boundary (what LssTangent.checkpoint appends, fds/lsstan.py:32-33)."""
        self.Rs.append(np.asarray(R))
        self.bs.append(np.asarray(b))

    def solve(self):
        """alpha (K, m): min sum |alpha_i|^2 s.t. alpha_{i+1} = R_i alpha_i + b_i
        (fds/lsstan.py:36-50), on the device."""
        R, b = np.array(self.Rs), np.array(self.bs)
        assert R.ndim == 3 and b.ndim == 2 and R.shape == (b.shape[0], b.shape[1], b.shape[1])
        dev = self.device_ordinal if self.device_ordinal is not None else _solve_device()
        return device.lss_solve(R, b, dev)

    def lyapunov_exponents(self, segment_range=None):
        """log |diag R_i| per segment (fds/lsstan.py:58-64)."""
        R = np.array(self.Rs)
        if segment_range is not None:
            R = R[slice(*segment_range)]
        return np.log(np.abs(np.diagonal(R, axis1=1, axis2=2)))

    def lyapunov_covariant_vectors(self):
        """Backward recursion for the covariant vectors in the basis of the stored
        Q's; shape (m, K+1, m) (fds/lsstan.py:66-75)."""
        growth = np.exp(self.lyapunov_exponents().mean(0))
        cur = np.eye(self.m_modes())
        seq = [cur]
        for R in self.Rs[::-1]:
            cur = np.linalg.solve(R, cur) * growth
            seq.append(cur)
        return np.moveaxis(np.array(seq[::-1]), 2, 0)

    def lyapunov_covariant_magnitude_and_sin_angle(self):
        """|v_i| and the sines of the pairwise angles (fds/lsstan.py:77-85)."""
        v = self.lyapunov_covariant_vectors()
        mag = np.sqrt((v ** 2).sum(2))
        gram = np.einsum('ikc,jkc->ijk', v, v)
        cosang = gram / mag[:, None, :] / mag[None, :, :]
        i = np.arange(cosang.shape[0])
        cosang[i, i, :] = 1
        return mag, np.sqrt(1 - cosang ** 2)
