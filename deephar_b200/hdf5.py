"""Minimal pure-Python HDF5 reader / writer -- just enough for Keras weight files.

Why: every reference entry script calls `model.load_weights(<released .h5>)`
(exp/mpii/eval_mpii_singleperson.py:54, exp/h36m/eval_h36m.py:53,
exp/pennaction/eval_penn_multitask.py:76 with by_name=True) and h5py is not in this image.
What Keras 2.1.4 + h5py write (`keras/engine/topology.py::save_weights_to_hdf5_group`) is the
plain "earliest libver" subset of the format, which is what this module understands:

  reader : superblock v0/v1/v2/v3 (optionally behind a user block), object headers v1 and v2,
           old-style groups (symbol-table message -> v1 B-tree -> SNOD -> local heap) and new-style
           compact groups (Link messages), contiguous and compact datasets of fixed-point / IEEE
           float types, attributes v1/v2/v3 holding numeric data, fixed-length strings or
           variable-length strings (global heap).  Chunked / filtered datasets, dense link or
           attribute storage and external links raise `Hdf5Error` (Keras never produces them).
  writer : superblock v0, old-style groups, contiguous little-endian datasets, numeric and
           fixed-length-string attributes.  Used by the tests (a committed tiny .h5 fixture) and by
           `Model.save_weights('*.h5')`.

Format reference: "HDF5 File Format Specification Version 3.0" (restated from the published
spec; no code from libhdf5 / h5py / pyfive).
"""
import mmap
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_DATASPACE, MSG_LINKINFO, MSG_DATATYPE, MSG_LINK, MSG_LAYOUT = 0x01, 0x02, 0x03, 0x06, 0x08
MSG_FILTER, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMTAB, MSG_ATTRINFO = 0x0B, 0x0C, 0x10, 0x11, 0x15


# what parsing a damaged file image raises at the lowest level (bad offsets, sizes, strings, type codes)
CORRUPT = (IndexError, ValueError, TypeError, KeyError, AttributeError, OverflowError, UnicodeDecodeError, MemoryError)


class Hdf5Error(IOError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# =============================================================================
# reader
# =============================================================================
class _Buf(object):
    """Little-endian cursor over the file image."""

    def __init__(self, data, so=8, sl=8):
        self.d = data
        self.so, self.sl = so, sl

    def u(self, pos, n):
        return int.from_bytes(self.d[pos:pos + n], 'little')

    def off(self, pos):
        v = self.u(pos, self.so)
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v

    def ln(self, pos):
        return self.u(pos, self.sl)


class Dataset(object):
    def __init__(self, f, name, msgs):
        self._f = f
        self.name = name
        self._msgs = msgs
        self.attrs = f._attributes(msgs)
        self.shape = f._dataspace(self._one(MSG_DATASPACE))
        self.dtype, self._vlen = f._datatype(self._one(MSG_DATATYPE))

    def _one(self, t):
        for (mt, _, data) in self._msgs:
            if mt == t:
                return data
        raise Hdf5Error('%s: object header has no message 0x%02x' % (self.name, t))

    def __getitem__(self, key):
        a = self.read()
        return a if key == () or key is Ellipsis else a[key]

    def read(self):
        f, b = self._f, self._f._b
        for (mt, _, _) in self._msgs:
            if mt == MSG_FILTER:
                raise Hdf5Error('%s: filtered (compressed) datasets are not supported' % self.name)
        lay = self._one(MSG_LAYOUT)
        count = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        nbytes = count * self.dtype.itemsize
        ver = lay[0]
        if ver == 3:
            cls = lay[1]
            if cls == 1:
                addr = int.from_bytes(lay[2:2 + b.so], 'little')
                raw = bytes(nbytes) if addr == (1 << (8 * b.so)) - 1 else f._slice(f._base + addr, nbytes)
            elif cls == 0:
                size = int.from_bytes(lay[2:4], 'little')
                raw = bytes(lay[4:4 + size])
            else:
                raise Hdf5Error('%s: chunked datasets are not supported' % self.name)
        elif ver in (1, 2):
            rank, cls = lay[1], lay[2]
            p = 8
            if cls == 1:
                addr = int.from_bytes(lay[p:p + b.so], 'little')
                raw = f._slice(f._base + addr, nbytes)
            elif cls == 0:
                p += 4 * rank
                size = int.from_bytes(lay[p:p + 4], 'little')
                raw = bytes(lay[p + 4:p + 4 + size])
            else:
                raise Hdf5Error('%s: chunked datasets are not supported' % self.name)
        else:
            raise Hdf5Error('%s: data layout message version %d is not supported' % (self.name, ver))
        if self._vlen:
            raise Hdf5Error('%s: variable-length datasets are not supported' % self.name)
        return np.frombuffer(raw, dtype=self.dtype, count=count).reshape(self.shape).copy()


class Group(object):
    def __init__(self, f, name, msgs):
        self._f = f
        self.name = name
        self._msgs = msgs
        self.attrs = f._attributes(msgs)
        self._kids = None

    def _children(self):
        if self._kids is None:
            self._kids = self._f._links(self._msgs, self.name)
        return self._kids

    def keys(self):
        return list(self._children().keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            kids = node._children()
            if part not in kids:
                raise KeyError('%s (no member %r in %s)' % (path, part, node.name))
            node = node._f._open(kids[part], (node.name.rstrip('/') + '/' + part))
        return node

    def visit_datasets(self):
        """Yield (path relative to this group, Dataset) depth-first, members in file (name) order."""
        for k in self.keys():
            child = self[k]
            if isinstance(child, Group):
                for sub, d in child.visit_datasets():
                    yield k + '/' + sub, d
            else:
                yield k, child


class File(Group):
    def __init__(self, path):
        self._fh = open(path, 'rb')
        try:
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise Hdf5Error('%s is empty' % path)
        self._data = memoryview(self._mm)
        self.path = path
        self._objects = {}
        try:
            self._read_superblock(path)
        except Hdf5Error:
            self.close()
            raise
        except CORRUPT as e:            # a damaged image walks the parser off the data: report it as what it is
            self.close()
            raise Hdf5Error('%s: corrupt or truncated HDF5 file (%s: %s)' % (path, type(e).__name__, e))

    def _read_superblock(self, path):
        sb = 0
        while True:
            if sb + 8 > len(self._data):
                raise Hdf5Error('%s: not an HDF5 file (no superblock signature)' % path)
            if bytes(self._data[sb:sb + 8]) == SIGNATURE:
                break
            sb = 512 if sb == 0 else sb * 2
        d = self._data
        ver = d[sb + 8]
        if ver in (0, 1):
            so, sl = d[sb + 13], d[sb + 14]
            self._b = _Buf(d, so, sl)
            p = sb + 24 + (4 if ver == 1 else 0)
            base = self._b.off(p)
            p += 4 * so
            # root group symbol-table entry: link name offset, object header address, cache type, ...
            root_addr = self._b.off(p + so)
        elif ver in (2, 3):
            so, sl = d[sb + 9], d[sb + 10]
            self._b = _Buf(d, so, sl)
            base = self._b.off(sb + 12)
            root_addr = self._b.off(sb + 12 + 3 * so)
        else:
            raise Hdf5Error('%s: superblock version %d is not supported' % (path, ver))
        if so not in (4, 8) or sl not in (4, 8):
            raise Hdf5Error('%s: unsupported offset/length sizes %d/%d' % (path, so, sl))
        self._base = 0 if base == UNDEF else base
        Group.__init__(self, self, '/', self._object_header(self._base + root_addr))

    def close(self):
        try:
            self._data.release()
            self._mm.close()
        except BufferError:             # slices still referenced (e.g. by a traceback in flight): the GC unmaps later
            pass
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- low level --------------------------------------------------------------
    def _slice(self, pos, n):
        if pos < 0 or pos + n > len(self._data):
            raise Hdf5Error('%s: truncated file (need bytes %d..%d of %d)' % (self.path, pos, pos + n, len(self._data)))
        return self._data[pos:pos + n]

    def _open(self, addr, name):
        msgs = self._object_header(addr)
        types = set(m[0] for m in msgs)
        if MSG_LAYOUT in types or (MSG_DATATYPE in types and MSG_DATASPACE in types):
            return Dataset(self, name, msgs)
        return Group(self, name, msgs)

    def _object_header(self, addr):
        """-> [(type, flags, bytes)] with continuation blocks followed."""
        if addr in self._objects:
            return self._objects[addr]
        b, d = self._b, self._data
        msgs = []
        if bytes(d[addr:addr + 4]) == b'OHDR':
            flags = d[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nsz = 1 << (flags & 3)
            chunk = b.u(p, nsz)
            p += nsz
            blocks = [(p, chunk)]
            hdr = 4 + (2 if flags & 0x04 else 0)
            while blocks:
                q, size = blocks.pop(0)
                end = q + size
                while q + hdr <= end:
                    mt, ms, mf = d[q], b.u(q + 1, 2), d[q + 3]
                    q += hdr
                    body = bytes(d[q:q + ms])
                    q += ms
                    if mt == MSG_CONTINUATION:
                        ca, cl = int.from_bytes(body[:b.so], 'little'), int.from_bytes(body[b.so:b.so + b.sl], 'little')
                        if bytes(d[self._base + ca:self._base + ca + 4]) != b'OCHK':
                            raise Hdf5Error('bad object header continuation block')
                        blocks.append((self._base + ca + 4, cl - 8))
                    elif mt != 0:
                        msgs.append((mt, mf, body))
        else:
            if d[addr] != 1:
                raise Hdf5Error('%s: object header version %d at %d is not supported' % (self.path, d[addr], addr))
            nmsg = b.u(addr + 2, 2)
            size = b.u(addr + 8, 4)
            blocks = [(addr + 16, size)]
            while blocks and len(msgs) < nmsg + 64:
                q, size = blocks.pop(0)
                end = q + size
                while q + 8 <= end:
                    mt, ms, mf = b.u(q, 2), b.u(q + 2, 2), d[q + 4]
                    q += 8
                    body = bytes(d[q:q + ms])
                    q += ms
                    if mt == MSG_CONTINUATION:
                        ca, cl = int.from_bytes(body[:b.so], 'little'), int.from_bytes(body[b.so:b.so + b.sl], 'little')
                        blocks.append((self._base + ca, cl))
                    elif mt != 0:
                        msgs.append((mt, mf, body))
        self._objects[addr] = msgs
        return msgs

    # ---- groups -------------------------------------------------------------------
    def _links(self, msgs, gname):
        b = self._b
        out = {}
        for (mt, _, body) in msgs:
            if mt == MSG_SYMTAB:
                btree = self._base + int.from_bytes(body[:b.so], 'little')
                heap = self._base + int.from_bytes(body[b.so:2 * b.so], 'little')
                if bytes(self._data[heap:heap + 4]) != b'HEAP':
                    raise Hdf5Error('%s: bad local heap in group %s' % (self.path, gname))
                heap_data = self._base + b.off(heap + 8 + 2 * b.sl)
                self._walk_btree(btree, heap_data, out)
            elif mt == MSG_LINK:
                name, addr = self._link_message(body)
                out[name] = addr
            elif mt == MSG_LINKINFO:
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                fheap = int.from_bytes(body[p:p + b.so], 'little')
                if fheap != (1 << (8 * b.so)) - 1:
                    raise Hdf5Error('%s: group %s uses dense link storage (not supported)' % (self.path, gname))
        return out

    def _link_message(self, body):
        b = self._b
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nsz = 1 << (flags & 3)
        nlen = int.from_bytes(body[p:p + nsz], 'little')
        p += nsz
        name = body[p:p + nlen].decode('utf-8')
        p += nlen
        if ltype != 0:
            raise Hdf5Error('soft / external links are not supported (%s)' % name)
        return name, self._base + int.from_bytes(body[p:p + b.so], 'little')

    def _cstr(self, pos):
        end = pos
        d = self._data
        while d[end] != 0:
            end += 1
        return bytes(d[pos:end]).decode('utf-8')

    def _walk_btree(self, addr, heap_data, out):
        b, d = self._b, self._data
        if bytes(d[addr:addr + 4]) != b'TREE':
            raise Hdf5Error('%s: bad B-tree node at %d' % (self.path, addr))
        if d[addr + 4] != 0:
            raise Hdf5Error('%s: unexpected B-tree node type %d' % (self.path, d[addr + 4]))
        level, used = d[addr + 5], b.u(addr + 6, 2)
        p = addr + 8 + 2 * b.so
        for i in range(used):
            child = self._base + b.off(p + b.sl)
            p += b.sl + b.so
            if level > 0:
                self._walk_btree(child, heap_data, out)
            else:
                if bytes(d[child:child + 4]) != b'SNOD':
                    raise Hdf5Error('%s: bad symbol table node at %d' % (self.path, child))
                n = b.u(child + 6, 2)
                q = child + 8
                for _ in range(n):
                    name = self._cstr(heap_data + b.off(q))
                    out[name] = self._base + b.off(q + b.so)
                    q += 2 * b.so + 24

    # ---- messages -----------------------------------------------------------------
    def _dataspace(self, body):
        b = self._b
        ver, rank = body[0], body[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:          # null dataspace
                return (0,)
            p = 4
        else:
            raise Hdf5Error('dataspace message version %d is not supported' % ver)
        return tuple(int.from_bytes(body[p + i * b.sl:p + (i + 1) * b.sl], 'little') for i in range(rank))

    def _datatype(self, body):
        """-> (numpy dtype, vlen kind or None)."""
        cls, bits0 = body[0] & 0x0F, body[1]
        size = int.from_bytes(body[4:8], 'little')
        order = '>' if (bits0 & 1) else '<'
        if cls == 0:
            return np.dtype('%s%s%d' % (order, 'i' if bits0 & 0x08 else 'u', size)), None
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error('unsupported floating-point size %d' % size)
            return np.dtype('%sf%d' % (order, size)), None
        if cls == 3:
            return np.dtype('S%d' % size), None
        if cls == 9:
            kind = 'str' if (bits0 & 0x0F) == 1 else 'seq'
            return np.dtype('V%d' % size), kind
        raise Hdf5Error('unsupported datatype class %d' % cls)

    def _attributes(self, msgs):
        b = self._b
        out = {}
        for (mt, _, body) in msgs:
            if mt == MSG_ATTRINFO:
                flags = body[1]
                p = 2 + (2 if flags & 1 else 0)
                fheap = int.from_bytes(body[p:p + b.so], 'little')
                if fheap != (1 << (8 * b.so)) - 1:
                    raise Hdf5Error('%s: dense attribute storage is not supported' % self.path)
            if mt != MSG_ATTRIBUTE:
                continue
            ver = body[0]
            nsz, tsz, ssz = (int.from_bytes(body[2 + 2 * i:4 + 2 * i], 'little') for i in range(3))
            p = 8 + (1 if ver == 3 else 0)
            rnd = _pad8 if ver == 1 else (lambda n: n)
            name = body[p:p + nsz].split(b'\0', 1)[0].decode('utf-8')
            p += rnd(nsz)
            tbody = body[p:p + tsz]
            p += rnd(tsz)
            sbody = body[p:p + ssz]
            p += rnd(ssz)
            try:
                dtype, vlen = self._datatype(tbody)
                shape = self._dataspace(sbody)
            except Hdf5Error:
                out[name] = None
                continue
            count = int(np.prod(shape, dtype=np.int64)) if shape else 1
            raw = body[p:p + count * dtype.itemsize]
            if vlen == 'str':
                vals = []
                for i in range(count):
                    rec = raw[i * dtype.itemsize:(i + 1) * dtype.itemsize]
                    vals.append(self._global_heap_object(int.from_bytes(rec[4:4 + b.so], 'little'),
                                                         int.from_bytes(rec[4 + b.so:8 + b.so], 'little'),
                                                         int.from_bytes(rec[:4], 'little')))
                out[name] = vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
            elif vlen:
                out[name] = None
            else:
                a = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape).copy()
                out[name] = a[()] if shape == () else a
        return out

    def _global_heap_object(self, addr, index, length):
        b, d = self._b, self._data
        pos = self._base + addr
        if bytes(d[pos:pos + 4]) != b'GCOL':
            raise Hdf5Error('%s: bad global heap collection' % self.path)
        size = b.ln(pos + 8)
        q, end = pos + 8 + b.sl, pos + size
        while q + 8 + b.sl <= end:
            idx, osz = b.u(q, 2), b.ln(q + 8)
            if idx == index:
                return bytes(d[q + 8 + b.sl:q + 8 + b.sl + min(osz, length)])
            if idx == 0:
                break
            q += 8 + b.sl + _pad8(osz)
        raise Hdf5Error('%s: global heap object %d not found' % (self.path, index))


# =============================================================================
# writer
# =============================================================================
class _WNode(object):
    def __init__(self, data=None):
        self.children = {}       # groups only
        self.attrs = []          # [(name, value)]
        self.data = data         # ndarray for datasets


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        size = dt.itemsize
        props = {2: (0, 16, 10, 5, 0, 10, 15), 4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[size]
        off, prec, epos, esz, mpos, msz, bias = props
        sign = prec - 1
        body = struct.pack('<BBBBI', 0x11, 0x20, sign, 0, size)
        body += struct.pack('<HHBBBBI', off, prec, epos, esz, mpos, msz, bias)
        return body
    if dt.kind in 'iu':
        size = dt.itemsize
        body = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0, 0, 0, size)
        body += struct.pack('<HH', 0, 8 * size)
        return body
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, dt.itemsize)       # null-padded ASCII
    raise Hdf5Error('writer: unsupported dtype %s' % dt)


def _dataspace_message(shape):
    rank = len(shape)
    body = struct.pack('<BBBB4x', 1, rank, 0, 0)
    for s in shape:
        body += struct.pack('<Q', s)
    return body


def _message(mtype, body, flags=0):
    pad = _pad8(len(body)) - len(body)
    return struct.pack('<HHB3x', mtype, len(body) + pad, flags) + body + b'\0' * pad


def _attr_message(name, value):
    if isinstance(value, (bytes, str)):
        value = np.array(value.encode('utf-8') if isinstance(value, str) else value)
    a = np.asarray(value)
    if a.dtype.kind == 'U':
        a = np.char.encode(a, 'utf-8')
    if a.dtype.kind == 'S' and a.dtype.itemsize == 0:
        a = a.astype('S1')
    if a.dtype.byteorder == '>':
        a = a.astype(a.dtype.newbyteorder('<'))
    nm = name.encode('utf-8') + b'\0'
    tb, sb = _dtype_message(a.dtype), _dataspace_message(a.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(tb), len(sb))
    for part in (nm, tb, sb):
        body += part + b'\0' * (_pad8(len(part)) - len(part))
    body += np.ascontiguousarray(a).tobytes()
    if len(body) > 65000:
        raise Hdf5Error('attribute %s is too large for an object header message (%d bytes)' % (name, len(body)))
    return _message(MSG_ATTRIBUTE, body)


class Writer(object):
    """with Writer(path) as w: w.create_group('a'); w.create_dataset('a/b/x', arr); w.set_attr('a', 'k', v)"""
    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self, path):
        self.path = path
        self.root = _WNode()

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.close()

    def _node(self, path, create=True):
        node = self.root
        for part in [p for p in path.split('/') if p]:
            if part not in node.children:
                if not create:
                    raise KeyError(path)
                node.children[part] = _WNode()
            node = node.children[part]
        return node

    def create_group(self, path):
        return self._node(path)

    def create_dataset(self, path, data):
        parts = [p for p in path.split('/') if p]
        parent = self._node('/'.join(parts[:-1]))
        a = np.ascontiguousarray(data)
        if a.dtype.byteorder == '>':
            a = a.astype(a.dtype.newbyteorder('<'))
        parent.children[parts[-1]] = _WNode(a)

    def set_attr(self, path, name, value):
        self._node(path, create=False).attrs.append((name, value))

    # ---- serialisation ----------------------------------------------------------
    def close(self):
        self.buf = bytearray(96)              # superblock v0 (56 B) + root symbol-table entry (40 B)
        root_hdr, root_btree, root_heap = self._write_group(self.root)
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQII', 0, root_hdr, 1, 0) + struct.pack('<QQ', root_btree, root_heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        with open(self.path, 'wb') as f:
            f.write(self.buf)

    def _alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _write_header(self, messages):
        body = b''.join(messages)
        hdr = struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body))
        return self._alloc(hdr + body)

    def _write_dataset(self, node):
        a = node.data
        addr = self._alloc(a.tobytes()) if a.size else UNDEF
        msgs = [_message(MSG_DATASPACE, _dataspace_message(a.shape)),
                _message(MSG_DATATYPE, _dtype_message(a.dtype), flags=1),
                _message(MSG_LAYOUT, struct.pack('<BBQQ', 3, 1, addr, a.nbytes))]
        msgs += [_attr_message(n, v) for n, v in node.attrs]
        return self._write_header(msgs)

    def _write_group(self, node):
        # children first (their object header addresses go into the symbol table)
        entries = []
        for name in sorted(node.children, key=lambda s: s.encode('utf-8')):
            child = node.children[name]
            if child.data is not None:
                entries.append((name, self._write_dataset(child), 0, b'\0' * 16))
            else:
                h, bt, hp = self._write_group(child)
                entries.append((name, h, 1, struct.pack('<QQ', bt, hp)))
        # local heap: offset 0 holds the empty string
        heap = bytearray(b'\0' * 8)
        offs = {}
        for name, _, _, _ in entries:
            offs[name] = len(heap)
            nb = name.encode('utf-8') + b'\0'
            heap += nb + b'\0' * (_pad8(len(nb)) - len(nb))
        free_off = len(heap)
        heap += struct.pack('<QQ', 1, 16)            # one free block at the end: next = 1 (none), size 16
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_off, heap_data))
        # symbol-table nodes of <= 2K entries, then B-tree levels of <= 2K' children
        cap = 2 * self.LEAF_K
        nodes = []            # (address, key offset of the largest name below)
        for i in range(0, max(len(entries), 1), cap):
            part = entries[i:i + cap]
            body = b'SNOD' + struct.pack('<BBH', 1, 0, len(part))
            for name, hdr, ctype, scratch in part:
                body += struct.pack('<QQII', offs[name], hdr, ctype, 0) + scratch
            body += b'\0' * (40 * (cap - len(part)))
            nodes.append((self._alloc(body), offs[part[-1][0]] if part else 0))
        level = 0
        icap = 2 * self.INTERNAL_K
        while True:
            parents = []
            for i in range(0, len(nodes), icap):
                part = nodes[i:i + icap]
                first_key = 0 if i == 0 else nodes[i - 1][1]
                body = b'TREE' + struct.pack('<BBH', 0, level, len(part))
                body += struct.pack('<QQ', UNDEF, UNDEF)        # siblings (filled below for completeness)
                body += struct.pack('<Q', first_key)
                for addr, key in part:
                    body += struct.pack('<QQ', addr, key)
                body += b'\0' * (16 * (icap - len(part)))
                parents.append((self._alloc(body), part[-1][1]))
            # sibling pointers
            for j, (addr, _) in enumerate(parents):
                left = parents[j - 1][0] if j > 0 else UNDEF
                right = parents[j + 1][0] if j + 1 < len(parents) else UNDEF
                self.buf[addr + 8:addr + 24] = struct.pack('<QQ', left, right)
            nodes = parents
            level += 1
            if len(nodes) == 1:
                break
        btree = nodes[0][0]
        msgs = [_message(MSG_SYMTAB, struct.pack('<QQ', btree, heap_addr))]
        msgs += [_attr_message(n, v) for n, v in node.attrs]
        return self._write_header(msgs), btree, heap_addr
