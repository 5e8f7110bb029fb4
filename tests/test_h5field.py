"""The /field wire format of the reference's file mode (fds/compute.py:91-106) read and written without an HDF5
library (fds_b200/h5field.py).  The reader is pinned against the one libhdf5-written file this image holds
(scipy's MATLAB v7.3 test file); the writer against that file's structures and through the reader."""
import os
import struct

import numpy as np
import pytest

from fds_b200 import h5field


def _libhdf5_file():
    try:
        import scipy.io.matlab
    except ImportError:
        return None
    p = os.path.join(os.path.dirname(scipy.io.matlab.__file__), 'tests', 'data', 'testhdf5_7.4_GLNX86.mat')
    return p if os.path.exists(p) else None


needs_libhdf5_file = pytest.mark.skipif(_libhdf5_file() is None, reason='scipy test data not installed')


@needs_libhdf5_file
def test_reader_on_a_file_written_by_libhdf5():
    """Superblock v0 behind a 512-byte user block with a non-zero base address, symbol-table root group (B-tree,
    local heap, symbol-table node), a v1 object header with seven messages, dataspace v1, layout v2."""
    a = h5field.read_dataset(_libhdf5_file(), '/testdouble')
    assert a.dtype == np.dtype('<f8') and a.shape == (9, 1)
    assert np.array_equal(a.ravel(), np.arange(9) * (np.pi / 4))
    with pytest.raises(KeyError):
        h5field.read_dataset(_libhdf5_file(), '/field')


@needs_libhdf5_file
def test_writer_structures_equal_the_librarys():
    """The bytes libhdf5 wrote for the IEEE binary64 datatype message, the superblock prefix and the sizes of the
    group B-tree node (K = 16) and symbol-table node (K = 4) against what write_field produces."""
    raw = open(_libhdf5_file(), 'rb').read()[512:]
    head, eof = h5field.field_file_header((40,), np.float64)
    assert head[:8] == raw[:8] and head[8:20] == raw[8:20]            # signature, versions, sizes, K values
    lib_type = raw[raw.index(b'\x03\x00\x18\x00\x01\x00\x00\x00') + 8:][:20]
    mine = head[head.index(b'\x03\x00\x18\x00\x01\x00\x00\x00') + 8:][:20]
    assert mine == lib_type
    tree, root_hdr = raw.index(b'TREE'), struct.unpack_from('<Q', raw, 56 + 8)[0]
    assert root_hdr - tree == head.index(b'HEAP') - head.index(b'TREE') == 544
    # same message framing as the library's root group header: symbol-table message, 16 bytes, B-tree and heap
    assert raw[root_hdr + 16:root_hdr + 20] == head[96 + 16:96 + 20] == b'\x11\x00\x10\x00'
    assert struct.unpack_from('<QQ', head, 96 + 24) == (head.index(b'TREE'), head.index(b'HEAP'))
    assert eof == len(head) + 40 * 8 and len(head) % 2048 == 0


@pytest.mark.parametrize('a', [np.arange(40.) * 0.1, np.random.RandomState(0).rand(1000, 3), np.zeros(0),
                               np.arange(7, dtype='>f8'), np.arange(5, dtype=np.int32),
                               np.float32(np.random.RandomState(1).rand(3, 4, 5)), np.uint8([[1, 2], [3, 4]])])
def test_round_trip(tmp_path, a):
    p = str(tmp_path / 'x.h5')
    h5field.write_field(p, a)
    b = h5field.read_field(p)
    assert b.shape == a.shape and b.dtype == a.dtype.newbyteorder('=') and np.array_equal(a, b)
    assert os.path.getsize(p) == struct.unpack_from('<Q', open(p, 'rb').read(48), 40)[0]   # end-of-file address
    assert not [f for f in os.listdir(str(tmp_path)) if 'tmp' in f]


def test_written_file_layout_is_the_default_h5py_layout(tmp_path):
    """Offsets every file of h5py.File(p, 'w').create_dataset('field', data=a) has with the default format
    bounds: root header 96, B-tree 136, heap 680, heap data 712 with the name at offset 8, dataset header 800,
    raw data at a multiple of 2048."""
    p = str(tmp_path / 'u0.h5')
    h5field.write_field(p, np.arange(12.))
    raw = open(p, 'rb').read()
    assert raw.index(b'TREE') == 136 and raw.index(b'HEAP') == 680 and raw.index(b'SNOD') % 8 == 0
    assert raw[712 + 8:712 + 14] == b'field\0'
    assert struct.unpack_from('<BBHII', raw, 800) == (1, 0, 4, 1, 112)
    assert np.array_equal(np.frombuffer(raw[2048:], dtype='<f8'), np.arange(12.))


def test_refusals(tmp_path):
    p = str(tmp_path / 'bad.h5')
    open(p, 'wb').write(b'not an hdf5 file' * 100)
    with pytest.raises(h5field.H5FormatError, match='not an HDF5 file'):
        h5field.read_field(p)
    good = str(tmp_path / 'good.h5')
    h5field.write_field(good, np.arange(100.))
    raw = open(good, 'rb').read()
    open(p, 'wb').write(raw[:-80])
    with pytest.raises(h5field.H5FormatError, match='truncated'):
        h5field.read_field(p)
    with pytest.raises(TypeError):
        h5field.write_field(p, np.array(['a', 'b']))


def test_unallocated_storage_reads_as_the_default_fill_value(tmp_path):
    """create_dataset('field', shape, dtype) that was never written (fds/compute.py:104 before the collective write):
    late allocation leaves the address undefined and the library returns zeros."""
    head, eof = h5field.field_file_header((5,), np.float64)
    raw = bytearray(head)
    at = raw.index(struct.pack('<BBQ', 3, 1, len(head)))          # the layout message: version 3, contiguous, address
    raw[at + 2:at + 10] = struct.pack('<Q', h5field._UNDEF)
    p = str(tmp_path / 'empty.h5')
    open(p, 'wb').write(bytes(raw) + b'\0' * 40)
    assert np.array_equal(h5field.read_field(p), np.zeros(5))


def test_new_style_file(tmp_path):
    """libver='latest' files: superblock v2, v2 object headers (OHDR), a compact group of link messages,
    dataspace v2, layout v4 -- assembled here by hand from the format specification."""
    data = np.arange(6.)
    name = b'field'

    def ohdr(msgs):
        body = b''.join(struct.pack('<BHB', t, len(m), 0) + m for t, m in msgs)
        h = b'OHDR' + struct.pack('<BB', 2, 0) + struct.pack('<B', len(body)) + body
        return h + b'\0\0\0\0'                                  # checksum (not verified by the reader)
    dset_addr, data_addr = 200, 400
    link = struct.pack('<BB', 1, 0) + struct.pack('<B', len(name)) + name + struct.pack('<Q', dset_addr)
    root = ohdr([(0x02, struct.pack('<BBQQ', 0, 0, h5field._UNDEF, h5field._UNDEF)), (0x06, link)])
    space = struct.pack('<BBBB', 2, 1, 0, 1) + struct.pack('<Q', 6)
    dtype = h5field._float_type(np.dtype('<f8'))
    layout = struct.pack('<BBQQ', 4, 1, data_addr, 48)
    dset = ohdr([(0x01, space), (0x03, dtype), (0x08, layout)])
    sb = h5field._SIG + struct.pack('<BBBB', 2, 8, 8, 0) + struct.pack('<4Q', 0, h5field._UNDEF, 448, 48) + b'\0' * 4
    raw = bytearray(448)
    raw[:len(sb)] = sb
    raw[48:48 + len(root)] = root
    raw[dset_addr:dset_addr + len(dset)] = dset
    raw[data_addr:] = data.tobytes()
    p = str(tmp_path / 'latest.h5')
    open(p, 'wb').write(bytes(raw))
    assert np.array_equal(h5field.read_field(p), data)


def test_chunked_compressed_dataset(tmp_path):
    """A 2-D /field stored in chunks behind a two-level chunk B-tree, shuffle + deflate (what
    create_dataset(..., chunks=(4, 3), shuffle=True, compression='gzip') stores): edge chunks stored whole, one
    chunk with the deflate filter masked out, one chunk never written (fill value) -- assembled by hand from the
    format specification."""
    import zlib
    U = h5field._UNDEF
    data = np.arange(10 * 7, dtype='<f8').reshape(10, 7) * 0.5 - 3
    chunk = (4, 3)
    want = data.copy()
    want[8:10, 6:7] = 0.0                                           # the chunk at (8, 6) is left out below

    def ohdr(msgs):
        body = b''.join(struct.pack('<BHB', t, len(m), 0) + m for t, m in msgs)
        return b'OHDR' + struct.pack('<BB', 2, 1) + struct.pack('<H', len(body)) + body + b'\0\0\0\0'

    blob = bytearray(4096)
    pos = [1024]

    def put(b):
        a = pos[0]
        blob[a:a + len(b)] = b
        pos[0] = a + len(b) + (-len(b) % 8)
        return a

    def key(nbytes, mask, offs):
        return struct.pack('<II', nbytes, mask) + struct.pack('<3Q', offs[0], offs[1], 0)

    def shuffle(raw):
        return np.frombuffer(raw, np.uint8).reshape(-1, 8).T.tobytes()

    entries = []
    for i in range(0, 10, 4):
        for j in range(0, 7, 3):
            if (i, j) == (8, 6):
                continue
            full = np.zeros(chunk)
            part = data[i:i + 4, j:j + 3]
            full[:part.shape[0], :part.shape[1]] = part
            raw = shuffle(full.astype('<f8').tobytes())
            mask = 0
            if (i, j) == (4, 3):
                mask = 2                                            # second filter (deflate) skipped for this chunk
            else:
                raw = zlib.compress(raw)
            entries.append((key(len(raw), mask, (i, j)), put(raw)))

    def node(level, ents, last):
        b = b'TREE' + struct.pack('<BBH', 1, level, len(ents)) + struct.pack('<QQ', U, U)
        for k, child in ents:
            b += k + struct.pack('<Q', child)
        return put(b + last)
    end = key(0, 0, (12, 9))
    leaf0 = node(0, entries[:5], entries[5][0])
    leaf1 = node(0, entries[5:], end)
    root_node = node(1, [(entries[0][0], leaf0), (entries[5][0], leaf1)], end)

    space = struct.pack('<BBBB', 2, 2, 0, 1) + struct.pack('<QQ', 10, 7)
    layout = struct.pack('<BBB', 3, 2, 3) + struct.pack('<Q', root_node) + struct.pack('<III', 4, 3, 8)
    pipeline = struct.pack('<BB', 2, 2) + struct.pack('<HHHI', 2, 0, 1, 8) + struct.pack('<HHHI', 1, 0, 1, 4)
    dset = put(ohdr([(0x01, space), (0x03, h5field._float_type(np.dtype('<f8'))), (0x0B, pipeline), (0x08, layout)]))
    link = struct.pack('<BB', 1, 0) + struct.pack('<B', 5) + b'field' + struct.pack('<Q', dset)
    root = ohdr([(0x02, struct.pack('<BBQQ', 0, 0, U, U)), (0x06, link)])
    blob[48:48 + len(root)] = root
    sb = h5field._SIG + struct.pack('<BBBB', 2, 8, 8, 0) + struct.pack('<4Q', 0, U, len(blob), 48) + b'\0' * 4
    blob[:len(sb)] = sb
    p = str(tmp_path / 'chunked.h5')
    open(p, 'wb').write(bytes(blob))
    got = h5field.read_field(p)
    assert got.shape == (10, 7) and np.array_equal(got, want)
