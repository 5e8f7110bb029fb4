"""The ``/field`` wire format of the reference's file mode, without an HDF5 library.

In the reference's MPI mode a state never enters the driver process: ``u0`` is the NAME of an HDF5 file whose one
dataset ``/field`` holds the flat state (fds/compute.py:91-106, ``mpi_read_field`` / ``mpi_write_field``;
tests/test_fun3d_mpi.py:43-50, ``save_hdf5`` / ``load_hdf5``), the perturbed initial conditions are written as
``<get_host_dir(run_id)>/input.h5`` (fds/segment.py:46,51) and the solver answers with the name of the file it
wrote.  Neither h5py nor libhdf5 is in this image, so the two functions below speak the file format directly
("HDF5 File Format Specification Version 3.0", The HDF Group):

``read_field(path)``   finds ``/field`` in a file written by libhdf5 -- superblock versions 0-3, old-style groups
                       (symbol table: B-tree v1 + local heap + symbol-table nodes) and new-style compact groups
                       (link messages), object headers v1 and v2, dataspace v1/v2, fixed-point and IEEE types
                       of either byte order, layout v1-v4 contiguous or compact, v1-v3 chunked (B-tree v1 of
                       chunks; deflate, shuffle, fletcher32 filters) -- and returns a numpy array.  Version-4
                       chunk indices, other filters and externally stored data are refused by name.
``write_field(path, a)`` writes what ``h5py.File(path, 'w').create_dataset('field', data=a)`` writes with the
                       library's default (earliest) format bounds: superblock v0, the root group as a symbol
                       table, one v1 object header (dataspace v1, datatype v1, fill value v2, layout v3,
                       contiguous), raw data at a 2048-byte boundary.

What pins it: the reader is checked against the one libhdf5-written file this image holds (scipy's MATLAB v7.3
test file, superblock v0 behind a 512-byte user block) and the writer byte for byte against structures of that
file and through the reader (tests/test_h5field.py).  Acceptance of written files BY h5py is unverified here
(DESIGN.md section 7 says so).
"""
import os
import struct

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


# --------------------------------------------------------------------------- #
# reader
# --------------------------------------------------------------------------- #
class _File:
    def __init__(self, path):
        self.path = path
        self.f = open(path, 'rb')
        self.size = os.fstat(self.f.fileno()).st_size
        self.base = 0
        self._superblock()

    def close(self):
        self.f.close()

    def at(self, addr, n):
        """n bytes at file address ``addr`` (relative to the base address, spec II.A)."""
        pos = addr + self.base
        if addr == _UNDEF or pos < 0 or pos + n > self.size:
            raise H5FormatError('%s: address %#x (+%d) outside the file' % (self.path, addr, n))
        self.f.seek(pos)
        b = self.f.read(n)
        if len(b) != n:
            raise H5FormatError('%s: short read at %#x' % (self.path, addr))
        return b

    def _superblock(self):
        # the signature sits at 0 or behind a user block of 512, 1024, ... bytes (spec II.A)
        off = 0
        while True:
            if off + 8 > self.size:
                raise H5FormatError('%s: not an HDF5 file' % self.path)
            self.f.seek(off)
            if self.f.read(8) == _SIG:
                break
            off = 512 if off == 0 else off * 2
        self.f.seek(off)
        head = self.f.read(128)
        ver = head[8]
        if ver in (0, 1):
            so, sl = head[13], head[14]
            p = 24 + (4 if ver == 1 else 0)
        elif ver in (2, 3):
            so, sl = head[9], head[10]
            p = 12
        else:
            raise H5FormatError('%s: superblock version %d' % (self.path, ver))
        if so != 8 or sl != 8:
            raise H5FormatError('%s: %d-byte offsets / %d-byte lengths are not supported' % (self.path, so, sl))
        if ver in (0, 1):
            base, _free, _eof, _drv = struct.unpack_from('<4Q', head, p)
            p += 32
            # root group symbol table entry: name offset, header address, cache type, reserved, scratch
            _name, self.root, cache = struct.unpack_from('<QQI', head, p)
        else:
            base, _ext, _eof, self.root = struct.unpack_from('<4Q', head, p)
        # a file copied behind a user block without h5repack keeps base 0 and absolute addresses
        self.base = base if base != _UNDEF else off

    # -- object headers (spec IV.A) ------------------------------------------------------------------- #
    def messages(self, addr):
        """[(type, flags, bytes)] of the object header at ``addr``, continuation blocks followed."""
        first = self.at(addr, 16)
        out = []
        if first[:4] == b'OHDR':
            self._messages_v2(addr, out)
        elif first[0] == 1:
            _v, _r, nmsg, _ref, size = struct.unpack_from('<BBHII', first, 0)
            blocks = [(addr + 16, size)]
            while blocks:
                a, n = blocks.pop(0)
                buf = self.at(a, n)
                p = 0
                while p + 8 <= n:
                    t, sz, fl = struct.unpack_from('<HHB', buf, p)
                    body = buf[p + 8:p + 8 + sz]
                    if len(body) != sz:
                        raise H5FormatError('%s: message overruns its header block' % self.path)
                    if t == 0x10:
                        blocks.append(struct.unpack_from('<QQ', body, 0))
                    else:
                        out.append((t, fl, body))
                    p += 8 + sz
        else:
            raise H5FormatError('%s: no object header at %#x' % (self.path, addr))
        return out

    def _messages_v2(self, addr, out):
        head = self.at(addr, 6)
        flags = head[5]
        p = 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        w = 1 << (flags & 3)
        size = int.from_bytes(self.at(addr + p, w), 'little')
        p += w
        track = bool(flags & 0x04)
        blocks = [(addr + p, size)]
        first = True
        while blocks:
            a, n = blocks.pop(0)
            buf = self.at(a, n)
            q = 0
            if not first:
                if buf[:4] != b'OCHK':
                    raise H5FormatError('%s: bad continuation block' % self.path)
                q, n = 4, n - 4                      # signature in front, checksum behind
            first = False
            hdr = 4 + (2 if track else 0)
            while q + hdr <= n:
                t, sz, fl = buf[q], int.from_bytes(buf[q + 1:q + 3], 'little'), buf[q + 3]
                body = buf[q + hdr:q + hdr + sz]
                if t == 0x10:
                    blocks.append(struct.unpack_from('<QQ', body, 0))
                else:
                    out.append((t, fl, body))
                q += hdr + sz

    # -- groups (spec III.A-D, IV.A.2.g/r) --------------------------------------------------------------- #
    def lookup(self, group_addr, name):
        want = name.encode()
        for t, _fl, body in self.messages(group_addr):
            if t == 0x11:                               # symbol table: B-tree v1 + local heap
                btree, heap = struct.unpack_from('<QQ', body, 0)
                return self._stab_lookup(btree, heap, want)
            if t == 0x06:                               # link message (compact new-style group)
                got = self._link(body)
                if got is not None and got[0] == want:
                    return got[1]
            if t == 0x02:                               # link info: dense storage lives in a fractal heap
                fheap = struct.unpack_from('<Q', body, 2 + (8 if body[1] & 1 else 0))[0]
                if fheap != _UNDEF:
                    raise H5FormatError('%s: densely stored groups are not supported' % self.path)
        return None

    def _link(self, body):
        flags = body[1]
        p = 2
        kind = 0
        if flags & 0x08:
            kind = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        w = 1 << (flags & 3)
        n = int.from_bytes(body[p:p + w], 'little')
        p += w
        nm = body[p:p + n]
        p += n
        if kind != 0:
            return None                                 # soft / external links are not followed
        return nm, struct.unpack_from('<Q', body, p)[0]

    def _heap_data(self, heap):
        h = self.at(heap, 32)
        if h[:4] != b'HEAP':
            raise H5FormatError('%s: no local heap at %#x' % (self.path, heap))
        size, _free, data = struct.unpack_from('<QQQ', h, 8)
        return self.at(data, size)

    def _stab_lookup(self, btree, heap, want):
        names = self._heap_data(heap)

        def name_at(off):
            end = names.index(b'\0', off)
            return names[off:end]

        def walk(node):
            h = self.at(node, 8)
            if h[:4] == b'SNOD':
                n = struct.unpack_from('<H', h, 6)[0]
                ents = self.at(node + 8, 40 * n)
                for i in range(n):
                    off, hdr = struct.unpack_from('<QQ', ents, 40 * i)
                    if name_at(off) == want:
                        return hdr
                return None
            if h[:4] != b'TREE' or h[4] != 0:
                raise H5FormatError('%s: no group B-tree node at %#x' % (self.path, node))
            used = struct.unpack_from('<H', h, 6)[0]
            body = self.at(node + 24, 16 * used + 8)
            for i in range(used):                       # key, child, key, child, ..., key
                child = struct.unpack_from('<Q', body, 16 * i + 8)[0]
                got = walk(child)
                if got is not None:
                    return got
            return None

        return walk(btree)

    # -- datasets ------------------------------------------------------------------------------------------ #
    def dataset(self, addr):
        shape = dtype = None
        layout = None
        filters = []
        for t, fl, body in self.messages(addr):
            if fl & 0x02 and t in (0x01, 0x03, 0x05, 0x08):
                raise H5FormatError('%s: shared (committed) header messages are not supported' % self.path)
            if t == 0x01:
                shape = _dataspace(body, self.path)
            elif t == 0x03:
                dtype = _datatype(body, self.path)
            elif t == 0x08:
                layout = body
            elif t == 0x0B:
                filters = _filters(body, self.path)
            elif t == 0x07:
                raise H5FormatError('%s: externally stored data is not supported' % self.path)
        if shape is None or dtype is None or layout is None:
            raise H5FormatError('%s: object at %#x is not a dataset' % (self.path, addr))
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = count * dtype.itemsize
        ver = layout[0]
        if ver in (1, 2):
            rank, cls = layout[1], layout[2]
            if cls == 1:
                data = struct.unpack_from('<Q', layout, 8)[0]
                raw = self._raw(data, nbytes)
            elif cls == 0:
                p = 8 + 4 * rank
                n = struct.unpack_from('<I', layout, p)[0]
                raw = layout[p + 4:p + 4 + n]
            else:
                dims = struct.unpack_from('<%dI' % rank, layout, 16)
                return self._chunked(struct.unpack_from('<Q', layout, 8)[0], shape, dtype, dims[:-1], filters)
        elif ver in (3, 4):
            cls = layout[1]
            if cls == 1:
                data, n = struct.unpack_from('<QQ', layout, 2)
                if n < nbytes:
                    raise H5FormatError('%s: %d bytes stored, %d expected' % (self.path, n, nbytes))
                raw = self._raw(data, nbytes)
            elif cls == 0:
                n = struct.unpack_from('<H', layout, 2)[0]
                raw = layout[4:4 + n]
            elif cls == 2 and ver == 3:
                rank = layout[2]                           # dataset rank + 1: the last "dimension" is the element size
                dims = struct.unpack_from('<%dI' % rank, layout, 11)
                return self._chunked(struct.unpack_from('<Q', layout, 3)[0], shape, dtype, dims[:-1], filters)
            else:
                raise H5FormatError('%s: version-4 chunk indices and virtual datasets are not supported' % self.path)
        else:
            raise H5FormatError('%s: data layout version %d' % (self.path, ver))
        if len(raw) < nbytes:
            raise H5FormatError('%s: %d bytes stored, %d expected' % (self.path, len(raw), nbytes))
        return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)

    # -- chunked storage: B-tree v1 of raw-data chunks (spec III.A.1, node type 1) ------------------------ #
    def _chunked(self, btree, shape, dtype, chunk, filters):
        if len(chunk) != len(shape) or not shape:
            raise H5FormatError('%s: chunk rank %d for a dataset of rank %d' % (self.path, len(chunk), len(shape)))
        out = np.zeros(shape, dtype=dtype)                 # chunks that were never written hold the fill value
        csize = int(np.prod(chunk, dtype=np.int64)) * dtype.itemsize
        rank = len(shape)

        def walk(node):
            h = self.at(node, 24)
            if h[:4] != b'TREE' or h[4] != 1:
                raise H5FormatError('%s: no chunk B-tree node at %#x' % (self.path, node))
            level, used = h[5], struct.unpack_from('<H', h, 6)[0]
            ksz = 8 + 8 * (rank + 1)
            body = self.at(node + 24, used * (ksz + 8) + ksz)
            for i in range(used):
                q = i * (ksz + 8)
                nbytes, mask = struct.unpack_from('<II', body, q)
                offs = struct.unpack_from('<%dQ' % rank, body, q + 8)
                child = struct.unpack_from('<Q', body, q + ksz)[0]
                if level:
                    walk(child)
                    continue
                raw = bytes(self.at(child, nbytes))
                for k in range(len(filters) - 1, -1, -1):  # undo the pipeline back to front, skipping masked filters
                    if not mask >> k & 1:
                        raw = filters[k](raw)
                if len(raw) != csize:
                    raise H5FormatError('%s: chunk of %d bytes, %d expected' % (self.path, len(raw), csize))
                block = np.frombuffer(raw, dtype=dtype).reshape(chunk)
                sel = tuple(slice(o, min(o + c, n)) for o, c, n in zip(offs, chunk, shape))
                out[sel] = block[tuple(slice(0, t.stop - t.start) for t in sel)]   # edge chunks are stored whole

        if btree != _UNDEF:
            walk(btree)
        return out

    def _raw(self, data, nbytes):
        if nbytes == 0:
            return b''
        if data == _UNDEF:                              # never written: the library hands out the fill value (default 0)
            return bytes(nbytes)
        pos = data + self.base
        if pos + nbytes > self.size:
            raise H5FormatError('%s: truncated (%d bytes of data missing)' % (self.path, pos + nbytes - self.size))
        return np.fromfile(self.path, dtype=np.uint8, count=nbytes, offset=pos)


def _filters(body, path):
    """The filter pipeline message (spec IV.A.2.l) as a list of inverse transforms, in pipeline order."""
    import zlib
    ver, n = body[0], body[1]
    p = 8 if ver == 1 else 2
    out = []
    for _ in range(n):
        fid = struct.unpack_from('<H', body, p)[0]
        if ver == 1 or fid >= 256:
            nlen = struct.unpack_from('<H', body, p + 2)[0]
            p += 4
        else:
            nlen = 0
            p += 2
        _flags, ncd = struct.unpack_from('<HH', body, p)
        p += 4 + (nlen + (-nlen % 8 if ver == 1 else 0))
        cd = struct.unpack_from('<%dI' % ncd, body, p)
        p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
        if fid == 1:
            out.append(zlib.decompress)
        elif fid == 2:
            size = cd[0]
            out.append(lambda raw, size=size: np.frombuffer(raw, np.uint8).reshape(size, -1).T.tobytes()
                       if size > 1 and len(raw) % size == 0 else raw)
        elif fid == 3:
            out.append(lambda raw: raw[:-4])               # Fletcher-32 checksum behind the data (not verified)
        else:
            raise H5FormatError('%s: filter %d is not supported (deflate, shuffle and fletcher32 are)' % (path, fid))
    return out


def _dataspace(body, path):
    ver, rank, flags = body[0], body[1], body[2]
    if ver == 1:
        p = 8
    elif ver == 2:
        if body[3] == 2:
            raise H5FormatError('%s: null dataspace' % path)
        p = 4
    else:
        raise H5FormatError('%s: dataspace version %d' % (path, ver))
    return tuple(struct.unpack_from('<%dQ' % rank, body, p)) if rank else ()


def _datatype(body, path):
    cls, ver = body[0] & 0x0F, body[0] >> 4
    bits0 = body[1]
    size = struct.unpack_from('<I', body, 4)[0]
    order = '>' if bits0 & 1 else '<'
    if cls == 0:                                        # fixed point
        kind = 'i' if bits0 & 0x08 else 'u'
        _off, prec = struct.unpack_from('<HH', body, 8)
        if prec != 8 * size or size not in (1, 2, 4, 8):
            raise H5FormatError('%s: %d-bit integers in %d bytes' % (path, prec, size))
        return np.dtype(order + kind + str(size))
    if cls == 1:                                        # floating point: accept exactly IEEE binary32 / binary64
        _off, prec, epos, esize, mpos, msize, bias = struct.unpack_from('<HHBBBBI', body, 8)
        ieee = {4: (32, 23, 8, 0, 23, 127), 8: (64, 52, 11, 0, 52, 1023)}
        if ieee.get(size) != (prec, epos, esize, mpos, msize, bias) or bits0 & 0x40:
            raise H5FormatError('%s: floating-point type is not IEEE binary32/64' % path)
        return np.dtype(order + 'f' + str(size))
    raise H5FormatError('%s: datatype class %d (version %d) is not a plain number' % (path, cls, ver))


def read_dataset(path, name):
    """The dataset ``name`` (a path from the root group, e.g. ``'/field'``) as a numpy array."""
    f = _File(path)
    try:
        addr = f.root
        for part in [p for p in name.split('/') if p]:
            addr = f.lookup(addr, part)
            if addr is None:
                raise KeyError('%s: no object %r' % (path, name))
        return f.dataset(addr)
    finally:
        f.close()


def read_field(path):
    """``handle['/field'][:]`` of fds/compute.py:91-97 and tests/test_fun3d_mpi.py:47-50, native byte order."""
    a = read_dataset(path, '/field')
    return np.ascontiguousarray(a, dtype=a.dtype.newbyteorder('='))


# --------------------------------------------------------------------------- #
# writer
# --------------------------------------------------------------------------- #
_GROUP_LEAF_K, _GROUP_INTERNAL_K = 4, 16               # library defaults, stored in the superblock
_ROOT_HDR, _BTREE, _HEAP, _HEAP_DATA, _DSET_HDR = 96, 136, 680, 712, 800
_HEAP_DATA_SIZE = 88
_DATA_ALIGN = 2048                                     # the library's metadata aggregation block


def _msg(kind, body, flags=0):
    body = body + b'\0' * (-len(body) % 8)              # v1 headers: every message padded to 8 bytes
    return struct.pack('<HHB3x', kind, len(body), flags) + body


def _float_type(dt):
    size = dt.itemsize
    prec, epos, esize, msize, bias = {4: (32, 23, 8, 23, 127), 8: (64, 52, 11, 52, 1023)}[size]
    bits0 = 0x20 | (1 if dt.byteorder == '>' else 0)    # mantissa normalisation: implied leading 1
    return struct.pack('<BBBBI', 0x11, bits0, prec - 1, 0, size) + \
        struct.pack('<HHBBBBI', 0, prec, epos, esize, 0, msize, bias)


def _int_type(dt):
    bits0 = (0x08 if dt.kind == 'i' else 0) | (1 if dt.byteorder == '>' else 0)
    return struct.pack('<BBBBI', 0x10, bits0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)


def field_file_header(shape, dtype, name='field'):
    """(bytes in front of the raw data, address of the end of the file) for one contiguous dataset ``name`` in
    the root group."""
    dt = np.dtype(dtype)
    if dt.byteorder == '=':
        dt = dt.newbyteorder('<' if np.little_endian else '>')
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        tmsg = _float_type(dt)
    elif dt.kind in 'iu' and dt.itemsize in (1, 2, 4, 8):
        tmsg = _int_type(dt)
    else:
        raise TypeError('cannot store dtype %s' % dt)
    nm = name.encode() + b'\0'
    nm += b'\0' * (-len(nm) % 8)
    if 8 + len(nm) + 16 > _HEAP_DATA_SIZE:
        raise ValueError('dataset name too long')
    shape = tuple(int(s) for s in shape)
    rank = len(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize if rank else dt.itemsize

    # dataset object header (v1): dataspace, datatype, fill value, layout
    space = struct.pack('<BBB5x', 1, rank, 1) + struct.pack('<%dQ' % (2 * rank), *(shape + shape))
    fill = struct.pack('<BBBBI', 2, 2, 2, 1, 0)         # v2, allocate late, write if set, default value (size 0)
    snod = _DSET_HDR + 16 + len(_msg(1, space)) + len(_msg(3, tmsg)) + len(_msg(5, fill)) + 32
    snod += -snod % 8
    meta_end = snod + 8 + 2 * _GROUP_LEAF_K * 40
    data = meta_end + (-meta_end % _DATA_ALIGN)
    eof = data + nbytes
    layout = struct.pack('<BBQQ', 3, 1, data if nbytes else _UNDEF, nbytes)
    msgs = _msg(1, space) + _msg(3, tmsg, 1) + _msg(5, fill, 1) + _msg(8, layout)
    dset = struct.pack('<BBHII4x', 1, 0, 4, 1, len(msgs)) + msgs

    out = bytearray(data)
    # superblock v0 (spec II.A): versions, sizes, group K values, flags, four addresses, root symbol table entry
    sb = _SIG + struct.pack('<8B', 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack('<HHI', _GROUP_LEAF_K, _GROUP_INTERNAL_K, 0)
    sb += struct.pack('<4Q', 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack('<QQII', 0, _ROOT_HDR, 1, 0) + struct.pack('<QQ', _BTREE, _HEAP)
    assert len(sb) == _ROOT_HDR
    out[0:len(sb)] = sb
    # root group: a v1 object header with one symbol-table message
    root = struct.pack('<BBHII4x', 1, 0, 1, 1, 24) + _msg(0x11, struct.pack('<QQ', _BTREE, _HEAP))
    assert _ROOT_HDR + len(root) == _BTREE
    out[_ROOT_HDR:_BTREE] = root
    # group B-tree (spec III.A.1): one leaf node, one child; keys are heap offsets of names ('' and the largest
    # name in the child)
    bt = b'TREE' + struct.pack('<BBH', 0, 0, 1) + struct.pack('<QQ', _UNDEF, _UNDEF) + struct.pack('<QQQ', 0, snod, 8)
    bt += b'\0' * (24 + (2 * _GROUP_INTERNAL_K + 1) * 8 + 2 * _GROUP_INTERNAL_K * 8 - len(bt))
    assert _BTREE + len(bt) == _HEAP
    out[_BTREE:_HEAP] = bt
    # local heap (spec III.D): '' at 0, the name at 8, the rest one free block (next = 1 means none)
    free = 8 + len(nm)
    heap = b'HEAP' + struct.pack('<B3x', 0) + struct.pack('<QQQ', _HEAP_DATA_SIZE, free, _HEAP_DATA)
    seg = b'\0' * 8 + nm + struct.pack('<QQ', 1, _HEAP_DATA_SIZE - free)
    seg += b'\0' * (_HEAP_DATA_SIZE - len(seg))
    assert _HEAP + len(heap) == _HEAP_DATA and _HEAP_DATA + len(seg) == _DSET_HDR
    out[_HEAP:_DSET_HDR] = heap + seg
    out[_DSET_HDR:_DSET_HDR + len(dset)] = dset
    # symbol table node (spec III.C): one entry (name offset, header address, nothing cached)
    sn = b'SNOD' + struct.pack('<BBH', 1, 0, 1) + struct.pack('<QQII16x', 8, _DSET_HDR, 0, 0)
    out[snod:snod + len(sn)] = sn
    return bytes(out), eof


def write_field(path, a, name='field'):
    """``h5py.File(path, 'w').create_dataset('field', data=a)`` (tests/test_fun3d_mpi.py:43-46;
    fds/compute.py:99-106 writes the same dataset collectively).  Written under a temporary name and renamed, so a
    solver polling for the file never sees half of it."""
    a = np.ascontiguousarray(a)
    head, eof = field_file_header(a.shape, a.dtype, name)
    tmp = '%s.tmp%d' % (path, os.getpid())
    with open(tmp, 'wb') as f:
        f.write(head)
        a.tofile(f)
        assert f.tell() == eof
    os.replace(tmp, path)
