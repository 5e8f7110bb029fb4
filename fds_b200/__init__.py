"""fds_b200 -- B200-native finite-difference shadowing (drop-in for qiqi/fds's
segment loop on device plugins).  Same public names as the reference package
(fds/__init__.py:1-11): ``shadowing``, ``continue_shadowing``, ``lss_gradient`` and
the sub-modules ``checkpoint``, ``timeseries``, ``lsstan``, ``segment``, ``timedilation``.
"""
from . import fds as core
from . import lsstan
from . import timedilation
from . import segment
from . import checkpoint
from . import timeseries
from .fds import shadowing, continue_shadowing, lss_gradient, lss_gradient_series, shadowing_batch
from .plugins import L96, Lorenz63, VanDerPol, Circular, HostSolver
from ._lib import FdsError, pinned_empty

__all__ = ['shadowing', 'continue_shadowing', 'lss_gradient', 'lss_gradient_series', 'shadowing_batch', 'L96', 'Lorenz63', 'VanDerPol', 'Circular', 'HostSolver',
           'FdsError', 'pinned_empty', 'checkpoint', 'timeseries', 'lsstan', 'segment', 'timedilation', 'core']
