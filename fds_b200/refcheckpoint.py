"""Reading checkpoints written by the REFERENCE implementation (qiqi/fds).

The reference pickles its ``Checkpoint`` with dill (fds/checkpoint.py:1-18), and what it pickles is not
plain data: ``u0`` is an evaluated ``pascal_lite.symbolic_array``, but ``V`` and ``v`` are UN-EVALUATED lazy
expressions -- the last boundary of ``continue_shadowing`` computes ``R``, ``b``, ``G_dil``, ``g_dil`` only
(fds/fds.py:153-158) and leaves ``V = Q`` and ``v - b Q`` as operator graphs (``pascal_lite/operators/*.py``)
whose leaves hold the arrays of the last evaluated state.  The file therefore carries pickled classes of
``fds`` / ``pascal_lite`` and, through dill, the code objects of the operators' closures.

``load_reference_checkpoint`` reads such a file WITHOUT the reference (and without dill) on the path:

* a restricted unpickler maps every ``fds.*`` / ``pascal_lite.*`` class to an inert stand-in that only
  keeps its attributes, rebuilds numpy arrays and the namedtuple, and turns dill's function / code / cell
  constructors into inert placeholders -- no code from the file is ever executed;
* the operator graphs are then evaluated once by operator NAME (``add sub mul div pow neg sum reshape
  transpose getitem setitem broadcast ReduceSum qr``; pascal_lite/operators/arithmetics.py:13-39,
  indexing.py:9-23, shapes.py:9-29, linalg.py:15-47) with numpy, exactly as the reference's executor would
  on its first ``run_compute`` after loading (pascal_lite/symbolic_value.py:138-173, fds/compute.py:52-57).

This is file-format conversion of a one-off, KB- to MB-sized artefact -- the reference defers this
arithmetic into the file, so reading the file means finishing it (including LAPACK's sign convention of the
stored ``R``, which a device CholeskyQR would not reproduce).  It is not a compute path: the result is a plain
``fds_b200.checkpoint.Checkpoint`` of numpy arrays that ``continue_shadowing`` uploads to the device.

Caveat inherited from the reference: checkpoints it writes in the MIDDLE of a run (``i < num_segments``,
fds/fds.py:160-170) hold ``u0, V, v`` AFTER the next segment while the histories stop before it; only the
final checkpoint of a run (the one ``load_last_checkpoint`` picks up, tests/test_restart.py:52-62) is
consistent.  Files are read as they are.
"""
import collections
import io
import operator
import pickle

import numpy as np

from .checkpoint import Checkpoint
from .lsstan import LssTangent


class ReferenceCheckpointError(Exception):
    pass


# --------------------------------------------------------------------------- #
# inert stand-ins
# --------------------------------------------------------------------------- #
class _Stub:
    """Instance of a class of the reference: attributes only."""
    _ref_module = ''

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2:          # (dict, slots)
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)
        elif isinstance(state, dict):
            self.__dict__.update(state)



class _Inert(_Stub):
    """dill's function / code / cell / type / module placeholders: constructing or calling one gives another
    placeholder -- nothing stored in the file is ever executed."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()


_stub_classes = {}


def _stub_class(module, name):
    key = (module, name)
    if key not in _stub_classes:
        _stub_classes[key] = type(str(name), (_Stub,), {'_ref_module': module})
    return _stub_classes[key]


def _safe_getattr(obj, name, *default):
    if isinstance(obj, _Stub):
        if name == '__dict__':
            return {}
        return obj.__dict__.get(name, _Inert())
    return _Inert()


def _safe_setattr(obj, name, value):
    if isinstance(obj, _Stub):
        obj.__dict__[name] = value


def _create_namedtuple(name, fieldnames, modulename, defaults=None):
    if name == 'Checkpoint' and tuple(fieldnames) == Checkpoint._fields:
        return Checkpoint
    return collections.namedtuple(name, fieldnames, defaults=defaults)


def _create_array(f, args, state, npdict=None):
    array = f(*args)
    array.__setstate__(state)
    return array


_SAFE_TYPES = {'SliceType': slice, 'NoneType': type(None), 'EllipsisType': type(Ellipsis), 'TupleType': tuple,
               'ListType': list, 'DictType': dict, 'IntType': int, 'FloatType': float, 'BooleanType': bool,
               'StringType': str, 'BytesType': bytes,
               # (the spelling depends on the dill version: 0.3.x writes 'slice')
               'slice': slice, 'ellipsis': type(Ellipsis), 'tuple': tuple, 'list': list, 'dict': dict, 'int': int,
               'float': float, 'bool': bool, 'str': str, 'bytes': bytes}


def _load_type(name):
    """dill writes builtin types through _load_type (a slice index of a getitem operator arrives this way,
    pascal_lite/operators/indexing.py:9-13); everything that is not plain data (code, functions, cells) is inert."""
    return _SAFE_TYPES.get(name, _Inert)


def _eval_repr(text):
    return {'Ellipsis': Ellipsis, 'NotImplemented': NotImplemented}.get(text, _Inert())


_DILL = {
    '_load_type': _load_type,
    '_eval_repr': _eval_repr,
    '_create_namedtuple': _create_namedtuple,
    '_create_array': _create_array,
    '_setattr': _safe_setattr,
    '_getattr': _safe_getattr,
}
_OPERATOR = {k: getattr(operator, k) for k in ('add', 'sub', 'mul', 'truediv', 'pow', 'neg', 'getitem')}
_NUMPY_OK = {'ndarray', 'dtype', '_reconstruct', 'scalar', '_frombuffer', 'float64', 'int64', 'bool_'}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        top = module.split('.')[0]
        if top == 'numpy':
            if name in _NUMPY_OK:
                return super().find_class(module, name)
            raise ReferenceCheckpointError('unexpected numpy global %s.%s' % (module, name))
        if top in ('fds', 'pascal_lite'):
            if (module, name) == ('fds.checkpoint', 'Checkpoint'):       # written by name (dumps_reference_checkpoint)
                return Checkpoint
            return _stub_class(module, name)
        if module in ('dill._dill', 'dill.dill', 'dill'):
            return _DILL.get(name, _Inert)
        if module in ('_operator', 'operator'):
            return _OPERATOR.get(name, _Inert)
        if module == 'builtins':
            if name == 'getattr':
                return _safe_getattr
            if name == 'setattr':
                return _safe_setattr
            if name in ('tuple', 'list', 'dict', 'set', 'frozenset', 'slice', 'int', 'float', 'complex', 'bool',
                        'bytes', 'str', 'object', 'range'):
                return super().find_class(module, name)
            return _Inert
        if (module, name) in (('collections', 'OrderedDict'), ('copyreg', '_reconstructor'), ('copy_reg', '_reconstructor'),
                              ('_codecs', 'encode')):           # (protocol <= 2 stores array bytes through _codecs.encode)
            return super().find_class(module, name)
        raise ReferenceCheckpointError('refusing global %s.%s in a checkpoint file' % (module, name))


# --------------------------------------------------------------------------- #
# the lazy operator graphs
# --------------------------------------------------------------------------- #
def _is_value(x):
    return isinstance(x, _Stub) and 'owner' in x.__dict__ and 'shape' in x.__dict__


def _perform(op, ins):
    """One operator by name; returns a tuple of outputs."""
    name = getattr(op, 'name', None) or type(op).__name__
    kind = type(op).__name__
    if name in ('add', 'sub', 'mul', 'pow'):
        return (getattr(operator, name)(*ins),)
    if name == 'div':
        return (operator.truediv(*ins),)
    if name == 'neg':
        return (-ins[0],)
    if name == 'sum':
        return (ins[0].sum(op.axis),)
    if name == 'reshape':
        shape = tuple(op.shape)
        x = ins[0]
        return (x.reshape(shape) if x.size == int(np.prod(shape)) else x.reshape(shape + (-1,)),)
    if name == 'transpose':
        axes, x = tuple(op.axes), ins[0]
        return (x.transpose(axes) if x.ndim == len(axes) else x.transpose(axes + (-1,)),)
    if kind == 'getitem' or name.startswith('getitem'):
        return (ins[0][op.ind],)
    if kind == 'setitem' or name.startswith('setitem'):
        x = np.array(ins[0], copy=True)
        x[op.ind] = ins[1]
        return (x,)
    if name == 'broadcast':
        return (ins[0].reshape(ins[0].shape + (1,)),)
    if name == 'ReduceSum':                                   # Dot over the distributed (last) axis
        return ((ins[0] * ins[1]).sum(-1),)
    if name == 'qr':                                          # QRT: thin QR of A^T, Q returned transposed
        Q, R = np.linalg.qr(ins[0].T)
        return (Q.T, R)
    raise ReferenceCheckpointError('operator %r (%s) of the stored expression is not known' % (name, kind))


def _leaf_size(values):
    """Length of the distributed (last) axis: taken from any evaluated leaf (fds/compute.py:46-47)."""
    seen, stack = set(), [v for v in values if _is_value(v)]
    while stack:
        v = stack.pop()
        if id(v) in seen:
            continue
        seen.add(id(v))
        f = v.__dict__.get('field')
        if isinstance(f, np.ndarray) and f.ndim >= 1 and v.__dict__.get('is_distributed', True):
            return f.shape[-1]
        op = v.__dict__.get('owner')
        if op is not None:
            stack.extend(x for x in op.inputs if _is_value(x))
    return None


def _evaluate(value, done, size):
    """Array of a (possibly lazy) symbolic_array_value -- what the reference's executor computes
    (pascal_lite/symbolic_value.py:138-173 with the inputs of fds/compute.py:35-48)."""
    key = id(value)
    if key in done:
        return done[key]
    field = value.__dict__.get('field')
    op = value.__dict__.get('owner')
    if isinstance(field, np.ndarray):
        out = field
    elif op is None:
        if isinstance(field, (int, np.integer)) and size is not None:
            # builtin inputs (pascal_lite/symbolic_variable.py:231-239): 0 = zeros, otherwise a block of random
            # numbers drawn at evaluation time (fds/segment.py:70 draws one and overwrites every row of it)
            shape = tuple(value.shape) + (int(size),)
            out = np.random.rand(*shape) if field else np.zeros(int(size))
        else:
            raise ReferenceCheckpointError('the checkpoint holds an independent value without data')
    else:
        # constants meet the distributed axis through a trailing axis of length one (symbolic_value.py:150-156)
        ins = [_evaluate(x, done, size) if _is_value(x)
               else (lambda a: a.reshape(a.shape + (1,)))(np.asarray(x, dtype=np.float64)) for x in op.inputs]
        outs = _perform(op, ins)
        outputs = getattr(op, 'outputs', None) or (getattr(op, 'output', value),)
        for o, arr in zip(outputs, outs):
            done[id(o)] = arr
        if key not in done:
            raise ReferenceCheckpointError('malformed expression graph')
        out = done[key]
    done[key] = out
    return out


def _array(x, done, size=None):
    """symbolic_array / symbolic_array_value / ndarray / number -> ndarray."""
    if isinstance(x, _Stub):
        if _is_value(x):
            return np.array(_evaluate(x, done, size), dtype=np.float64)
        if 'value' in x.__dict__:
            return _array(x.value, done, size)
        raise ReferenceCheckpointError('unexpected object %s in a checkpoint' % type(x).__name__)
    return np.asarray(x, dtype=np.float64)


def load_reference_checkpoint(checkpoint_file):
    """A checkpoint file of the reference (``m{m}_segment{K}``) as a plain ``Checkpoint`` of numpy arrays that
    ``fds_b200.continue_shadowing`` and ``lss_gradient`` accept (9 fields in the reference's order)."""
    import sys
    with open(checkpoint_file, 'rb') as f:
        data = f.read()
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 20000))                    # setitem chains are as deep as the subspace dimension
    try:
        cp = _Unpickler(io.BytesIO(data)).load()
        if not (isinstance(cp, tuple) and len(cp) == 9):
            raise ReferenceCheckpointError('not a Checkpoint (9 fields expected)')
        u0, V, v, lss, G_lss, g_lss, J, G_dil, g_dil = cp
        done = {}
        size = _leaf_size([getattr(x, 'value', x) for x in (u0, V, v)])
        u0, V, v = _array(u0, done, size), _array(V, done, size), _array(v, done, size)
    finally:
        sys.setrecursionlimit(old)
    if not (isinstance(lss, _Stub) and 'Rs' in lss.__dict__ and 'bs' in lss.__dict__):
        raise ReferenceCheckpointError('the lss field is not an LssTangent')
    done = {}
    lss = LssTangent([_array(R, done) for R in lss.Rs], [_array(b, done) for b in lss.bs])
    G_lss = [[_array(g, done) for g in G] for G in G_lss]
    g_lss = [_array(g, done) for g in g_lss]
    J = [_array(j, done) for j in J]
    G_dil = [_array(g, done) for g in G_dil]
    g_dil = [float(_array(g, done)) for g in g_dil]
    if V.ndim != 2 or u0.ndim != 1 or v.shape != u0.shape or V.shape[1] != u0.shape[0]:
        raise ReferenceCheckpointError('unexpected shapes: u0 %s, V %s, v %s' % (u0.shape, V.shape, v.shape))
    return Checkpoint(u0, V, v, lss, G_lss, g_lss, J, G_dil, g_dil)


# --------------------------------------------------------------------------- #
# the other direction: a file the reference can load and continue from
# --------------------------------------------------------------------------- #
_REF_CLASSES = (('fds.checkpoint', 'Checkpoint'), ('fds.lsstan', 'LssTangent'),
                ('pascal_lite.symbolic_variable', 'symbolic_array'),
                ('pascal_lite.symbolic_value', 'symbolic_array_value'))


def dumps_reference_checkpoint(cp):
    """Bytes of a pickle that the REFERENCE's ``load_checkpoint`` (dill.load, fds/checkpoint.py:17-18) turns into ITS
    ``Checkpoint``: ``u0, V, v`` as evaluated ``pascal_lite.symbolic_array``s (``value.field`` = the array, site axis
    last: pascal_lite/symbolic_variable.py:47-60, symbolic_value.py:29-34), ``lss`` as its ``LssTangent`` with the
    ``Rs, bs`` lists (fds/lsstan.py:9-12), the histories as lists of arrays.  The classes are written BY NAME (the
    reference resolves them to its own on load); nothing of the reference is needed to write the file."""
    import sys
    import types
    stubs = {}
    for mod, name in _REF_CLASSES:
        if name == 'Checkpoint':
            cls = collections.namedtuple('Checkpoint', Checkpoint._fields)
        else:
            cls = type(name, (object,), {})
        cls.__module__, cls.__qualname__ = mod, name
        stubs[(mod, name)] = cls

    def sym(field, shape):
        val = stubs[('pascal_lite.symbolic_value', 'symbolic_array_value')]()
        val.__dict__.update(shape=tuple(shape), owner=None, field=np.array(field, dtype=np.float64), is_distributed=True)
        arr = stubs[('pascal_lite.symbolic_variable', 'symbolic_array')]()
        arr.__dict__.update(value=val)
        return arr

    V = np.asarray(cp.V, dtype=np.float64)
    lss = stubs[('fds.lsstan', 'LssTangent')]()
    lss.__dict__.update(Rs=[np.array(R, dtype=np.float64) for R in cp.lss.Rs],
                        bs=[np.array(b, dtype=np.float64) for b in cp.lss.bs])
    out = stubs[('fds.checkpoint', 'Checkpoint')](
        sym(cp.u0, ()), sym(V, (V.shape[0],)), sym(cp.v, ()), lss,
        [[np.array(g, dtype=np.float64) for g in G] for G in cp.G_lss],
        [np.array(g, dtype=np.float64) for g in cp.g_lss],
        [np.array(J, dtype=np.float64) for J in cp.J],
        [np.array(g, dtype=np.float64) for g in cp.G_dil],
        [np.float64(g) for g in cp.g_dil])
    # pickle writes classes as (module, name) after checking that the name resolves to the class: stand-in modules
    # for the duration of the dump (whatever was there -- e.g. the real reference -- is put back)
    names = sorted({m for m, _ in _REF_CLASSES} | {m.rsplit('.', 1)[0] for m, _ in _REF_CLASSES})
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in names:
            sys.modules[k] = types.ModuleType(k)
        for (mod, name), cls in stubs.items():
            setattr(sys.modules[mod], name, cls)
        return pickle.dumps(out, protocol=3)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def save_reference_checkpoint(checkpoint_path, cp):
    """Write ``cp`` under the reference's file name ``m{m}_segment{K}`` (fds/checkpoint.py:8-15) in a form the
    reference loads: a run continued here can go back."""
    import os
    name = 'm{0}_segment{1}'.format(cp.lss.m_modes(), cp.lss.K_segments())
    tmp = os.path.join(checkpoint_path, '.tmp_{0}_{1}'.format(name, os.getpid()))
    with open(tmp, 'wb') as f:
        f.write(dumps_reference_checkpoint(cp))
    os.replace(tmp, os.path.join(checkpoint_path, name))


def is_reference_checkpoint(checkpoint_file):
    """Does the file look like a dill pickle of the reference (it names ``pascal_lite``)?"""
    with open(checkpoint_file, 'rb') as f:
        head = f.read(1 << 16)
    return b'pascal_lite' in head or b'dill._dill' in head
