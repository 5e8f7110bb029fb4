"""Averaging utilities of the shadowing post-processing (host side, KBs of data).

Mirrors the reference module ``fds/timeseries.py`` (mean_std :3-7, windowed_mean
:9-12, exp_cum_mean :14-21): same names, same results.
"""
import numpy as np


def mean_std(x):
    """Mean and its standard error from 8 batch means -- the reference's only
    error estimator (fds/timeseries.py:3-7)."""
    x = np.asarray(x)
    cuts = np.linspace(0, len(x), 9).astype(int)
    batch = np.array([x[lo:hi].mean(0) for lo, hi in zip(cuts[:-1], cuts[1:])])
    return x.mean(0), batch.std(0) / np.sqrt(len(batch))


def windowed_mean(a):
    """Window-weighted mean over the leading axis; the reference's window is all
    ones (fds/timeseries.py:9-12), kept in the same form so rounding agrees."""
    a = np.asarray(a)
    win = np.ones(a.shape[0])
    return (a * win.reshape((-1,) + (1,) * (a.ndim - 1))).sum(0) / win.sum()


def exp_cum_mean(x):
    """Running mean with weights 1 - exp(-k / sqrt(n)) (fds/timeseries.py:14-21)."""
    x = np.asarray(x)
    out = []
    for n in range(1, len(x) + 1):
        w = 1 - np.exp(-np.arange(1, n + 1) / np.sqrt(n))
        w = w.reshape((-1,) + (1,) * (x.ndim - 1))
        out.append((x[:n] * w).sum(0) / w.sum())
    return np.array(out)
