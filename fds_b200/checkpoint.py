"""Checkpoint carrier between ``shadowing`` and ``continue_shadowing``.

Same 9-field order, file naming (``m{m}_segment{K}``) and load-latest semantics as
the reference ``fds/checkpoint.py`` (:6-52).  Unlike the reference (which pickles
un-evaluated lazy graphs with dill), ``u0``, ``V``, ``v`` here are plain numpy
arrays fetched from the device, so the standard pickle is enough.  Files written BY the
reference are recognised and read too (``refcheckpoint.py``), so a run started with the
reference can be continued here.
"""
import os
import pickle
from collections import namedtuple

Checkpoint = namedtuple('Checkpoint', 'u0 V v lss G_lss g_lss J G_dil g_dil')


def _filename(cp):
    return 'm{0}_segment{1}'.format(cp.lss.m_modes(), cp.lss.K_segments())


def save_checkpoint(checkpoint_path, cp, reference_format=None):
    """Written to a temporary name and renamed, so that a concurrent load_last_checkpoint never sees a
    truncated file (and never picks the temporary up: its name does not parse as m{m}_segment{K}).
    ``reference_format=True`` (or ``FDS_CHECKPOINT_FORMAT=reference`` in the environment) writes a file the
    REFERENCE loads and continues from (refcheckpoint.dumps_reference_checkpoint); both kinds are read back here."""
    if reference_format is None:
        reference_format = os.environ.get('FDS_CHECKPOINT_FORMAT', '') == 'reference'
    if reference_format:
        from . import refcheckpoint
        return refcheckpoint.save_reference_checkpoint(checkpoint_path, cp)
    final = os.path.join(checkpoint_path, _filename(cp))
    tmp = os.path.join(checkpoint_path, '.tmp_{0}_{1}'.format(_filename(cp), os.getpid()))
    with open(tmp, 'wb') as f:
        pickle.dump(cp, f)
    os.replace(tmp, final)


def load_checkpoint(checkpoint_file):
    """A checkpoint written by this package (plain pickle of numpy arrays) or by the reference (dill pickle of
    lazy pascal_lite graphs, fds/checkpoint.py:8-18 -- read by fds_b200/refcheckpoint.py without the reference
    on the path)."""
    from . import refcheckpoint
    if refcheckpoint.is_reference_checkpoint(checkpoint_file):
        return refcheckpoint.load_reference_checkpoint(checkpoint_file)
    with open(checkpoint_file, 'rb') as f:
        return pickle.load(f)


def verify_checkpoint(cp):
    """All per-segment histories have one entry per finished segment
    (reference fds/checkpoint.py:20-26)."""
    k = cp.lss.K_segments()
    return all(len(h) == k for h in (cp.G_lss, cp.g_lss, cp.J, cp.G_dil, cp.g_dil))


def _parse(filename):
    """(m, K) of a file named m{m}_segment{K}, else None (garbage names are ignored,
    reference tests/test_restart.py:46-50)."""
    head, sep, tail = filename.partition('_segment')
    if not sep or not head.startswith('m'):
        return None
    try:
        return int(head[1:]), int(tail)
    except ValueError:
        return None


def load_last_checkpoint(checkpoint_path, m):
    """The checkpoint with the largest segment count among those written for
    subspace dimension ``m``; None if there is none (fds/checkpoint.py:28-52)."""
    best = None
    for name in os.listdir(checkpoint_path):
        mk = _parse(name)
        if mk is not None and mk[0] == m and (best is None or mk[1] > best[0]):
            best = (mk[1], name)
    if best is not None:
        return load_checkpoint(os.path.join(checkpoint_path, best[1]))
