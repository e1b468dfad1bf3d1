"""Minimal HDF5 writer / reader for the matrix file (raw_peak_bc_matrix.h5).

The build image has no h5py / libhdf5 / pytables, so the container is written by hand against the HDF5 File
Format Specification: version-0 superblock, version-1 object headers, "old style" groups (symbol table = v1
B-tree + local heap + one SNOD per group), contiguous little-endian datasets of int32 / int64 / fixed-length
ASCII strings, fixed-length string and int64 scalar attributes.  This is the oldest, most widely readable
subset of the format.  Layout written = what `cellranger.matrix.CountMatrix.load` reads
(lib/cellranger/lib/python/cellranger/matrix.py:534-556,836-857, feature_ref.py:120-152):

    /            attrs: filetype='matrix', version=2, software_version=<str>
    /matrix/{barcodes S<n>[B], data int32[nnz], indices int64[nnz], indptr int64[B+1], shape int32[2]}
    /matrix/features/{_all_tag_keys, derivation, feature_type, genome, id, name}  (S<n> each)

The four matrix arrays (data, indices, indptr, shape) are stored as the reference stores them (matrix.py:441-447):
chunked (80000 elements), shuffle + gzip filters, unlimited maximum dimension -- a version-3 chunked layout message, a
version-1 filter pipeline message and a version-1 B-tree (node type 1) over the chunks; string datasets are contiguous
(cellranger/io.py:323-358 passes no storage options).  `compression=None` writes contiguous arrays instead.
The reader below understands this subset -- contiguous and chunked layouts, shuffle and deflate filters, object-header
continuation blocks -- so it also reads the arrays of a matrix written by the reference's h5py, and fails with a clear
message on anything else; compatibility with libhdf5 itself could not be exercised in the build image (stated in
DESIGN.md).
"""
import struct
import zlib

import numpy as np

CHUNK_ELEMS = 80000          # cellranger/matrix.py:29 HDF5_CHUNK_SIZE
GZIP_LEVEL = 4               # h5py's default for compression='gzip' (matrix.py:27,446)
ISTORE_K = 32                # chunk B-tree: up to 2 * 32 entries per node (the library default; a version-0 superblock cannot say otherwise)

UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K = 4
GROUP_INTERNAL_K = 16
SIGNATURE = b"\x89HDF\r\n\x1a\n"


def _pad8(b):
    return b + b"\0" * ((-len(b)) % 8)


# --------------------------------------------------------------------------------------------- messages
def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "i":
        # class 0 (fixed point), version 1; flags: little endian, signed
        return struct.pack("<BBBBI", 0x10, 0x08, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        # class 3 (string), version 1; flags: null padded (1), ASCII
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)
    raise TypeError("unsupported dtype %r" % dt)


def _dataspace_msg(shape, unlimited=False):
    # version 1: version, rank, flags (bit 0: maximum dimensions present), reserved(1), reserved(4), dims[, max dims]
    body = struct.pack("<BBBBI", 1, len(shape), 1 if unlimited else 0, 0, 0)
    for d in shape:
        body += struct.pack("<Q", int(d))
    if unlimited:
        for _ in shape:
            body += struct.pack("<Q", UNDEF)                                  # H5S_UNLIMITED
    return body


def _shuffle(raw, itemsize):
    """HDF5 shuffle filter: byte j of every element, for j = 0 .. itemsize - 1."""
    return np.frombuffer(raw, dtype=np.uint8).reshape(-1, itemsize).T.tobytes()


def _unshuffle(raw, itemsize):
    return np.frombuffer(raw, dtype=np.uint8).reshape(itemsize, -1).T.tobytes()


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _object_header(messages):
    blob = b"".join(messages)
    # version 1 header: version, reserved, n messages, ref count, header size, pad to 8
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(blob)) + b"\0" * 4 + blob


def _attribute_msg(name, value):
    if isinstance(value, (bytes, str)):
        raw = value.encode("ascii") if isinstance(value, str) else value
        if len(raw) == 0:
            raw = b"\0"
        dt = _dtype_msg("S%d" % len(raw))
        data = raw
    else:
        dt = _dtype_msg("<i8")
        data = struct.pack("<q", int(value))
    ds = _dataspace_msg(())
    nm = name.encode("ascii") + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
    return _message(0x000C, body)


# --------------------------------------------------------------------------------------------- writer
class _Writer(object):
    """Appends to the file as it goes (the 1.5 GB index / data arrays of a full-size matrix are written straight from the
    numpy buffers, never copied into a staging bytearray); the superblock is patched in at the end."""

    def __init__(self, f):
        self.f = f
        self.pos = 0

    def tell(self):
        return self.pos

    def alloc(self, blob, align=8):
        pad = (-self.pos) % align
        if pad:
            self.f.write(b"\0" * pad)
        addr = self.pos + pad
        self.f.write(blob)
        self.pos = addr + (blob.nbytes if isinstance(blob, memoryview) else len(blob))
        return addr

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        data_addr = self.alloc(memoryview(arr.reshape(-1).view(np.uint8))) if arr.nbytes else UNDEF
        msgs = [
            _message(0x0001, _dataspace_msg(arr.shape)),
            _message(0x0003, _dtype_msg(arr.dtype), flags=1),                  # constant message
            _message(0x0005, struct.pack("<BBBB", 2, 1, 2, 0)),                # fill value v2: early alloc, write if set, undefined
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)),   # layout v3, contiguous
        ]
        return self.alloc(_object_header(msgs))

    def dataset_chunked(self, arr, chunk=CHUNK_ELEMS, level=GZIP_LEVEL, threads=None):
        """1-D array as the reference stores its matrix arrays: chunks of `chunk` elements (the last one padded with
        zeros to full size, as the library does), each shuffled and deflated; a version-1 B-tree indexes the chunks."""
        arr = np.ascontiguousarray(arr).reshape(-1)
        item = arr.dtype.itemsize
        n_chunks = (arr.size + chunk - 1) // chunk

        # shuffle + deflate of all chunks on the host cores (libcratac: cratac_h5_filter_chunks), written in order
        from . import _ffi
        entries = []                                                          # (filtered size, element offset, address)
        if n_chunks:
            buf, stride, sizes = _ffi.h5_filter_chunks(arr, chunk, level, threads or 0)
            mv = memoryview(buf)
            for k in range(n_chunks):
                entries.append((int(sizes[k]), k * chunk, self.alloc(mv[k * stride:k * stride + int(sizes[k])])))
        key = lambda size, off: struct.pack("<IIQQ", size, 0, off, 0)         # chunk size, filter mask, offsets (element dim + the datatype dim)
        end_off = n_chunks * chunk
        node_bytes = 24 + 2 * ISTORE_K * 8 + (2 * ISTORE_K + 1) * 24          # nodes have their full size on disk
        level_nodes = entries
        node_level = 0
        btree_addr = UNDEF
        while level_nodes:
            groups = [level_nodes[g:g + 2 * ISTORE_K] for g in range(0, len(level_nodes), 2 * ISTORE_K)]
            pad = (-self.pos) % 8
            base = self.pos + pad                                             # the nodes of one level sit side by side
            blob = b""
            parents = []
            for gi, grp in enumerate(groups):
                nxt = groups[gi + 1][0][1] if gi + 1 < len(groups) else end_off
                left = base + (gi - 1) * node_bytes if gi > 0 else UNDEF
                right = base + (gi + 1) * node_bytes if gi + 1 < len(groups) else UNDEF
                node = b"TREE" + struct.pack("<BBHQQ", 1, node_level, len(grp), left, right)
                for size, off, addr in grp:
                    node += key(size, off) + struct.pack("<Q", addr)
                node += key(0, nxt)
                blob += node + b"\0" * (node_bytes - len(node))
                parents.append((grp[0][0], grp[0][1], base + gi * node_bytes))
            assert self.alloc(blob) == base
            if len(parents) == 1:
                btree_addr = parents[0][2]
                break
            level_nodes = parents
            node_level += 1
        filters = struct.pack("<BBHI", 1, 2, 0, 0)                            # pipeline v1, two filters
        filters += struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", item) + b"\0" * 4      # shuffle (optional), client data: element size
        filters += struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", level) + b"\0" * 4     # deflate (optional), client data: level
        msgs = [
            _message(0x0001, _dataspace_msg(arr.shape, unlimited=True)),
            _message(0x0003, _dtype_msg(arr.dtype), flags=1),
            _message(0x0005, struct.pack("<BBBB", 2, 3, 2, 0)),                # fill value v2: incremental allocation, write if set, undefined
            _message(0x000B, filters),
            _message(0x0008, struct.pack("<BBBQII", 3, 2, 2, btree_addr, chunk, item)),   # layout v3, chunked: rank + 1, B-tree, chunk dims, element size
        ]
        return self.alloc(_object_header(msgs))

    def group(self, children, attrs=None):
        """children: dict name -> object header address.  Returns (header addr, btree addr, heap addr)."""
        names = sorted(children)                                              # SNOD entries are ordered by name
        assert len(names) <= 2 * GROUP_LEAF_K, "one symbol-table node per group in this writer"
        # local heap data segment: "" at offset 0, then the names, then one free block
        seg = bytearray(b"\0" * 8)
        offs = {}
        for nm in names:
            offs[nm] = len(seg)
            seg += _pad8(nm.encode("ascii") + b"\0")
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 32) + b"\0" * 16                         # free block: next = H5HL_FREE_NULL (1), size 32
        seg_addr = self.alloc(bytes(seg))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(seg), free_off, seg_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for nm in names:
            snod += struct.pack("<QQII", offs[nm], children[nm], 0, 0) + b"\0" * 16
        snod += b"\0" * (8 + 2 * GROUP_LEAF_K * 40 - len(snod))
        snod_addr = self.alloc(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        if names:
            tree += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]])
        tree += b"\0" * (24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8 - len(tree))
        tree_addr = self.alloc(tree)
        msgs = [_message(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in (attrs or {}).items():
            msgs.append(_attribute_msg(k, v))
        return self.alloc(_object_header(msgs)), tree_addr, heap_addr


def _string_array(values):
    """cellranger/io.py:323-358 (create_hdf5_string_dataset): fixed-length ASCII, xmlcharrefreplace, min length 1."""
    enc = [("" if v is None else v).encode("ascii", "xmlcharrefreplace") if not isinstance(v, bytes) else v for v in values]
    n = max([len(x) for x in enc] + [1])
    return np.array(enc, dtype="S%d" % n) if enc else np.zeros(0, dtype="S1")


def write_matrix_h5(path, indptr, indices, data, n_features, barcodes, feature_ids, genome, sw_version="", compression="gzip"):
    """CountMatrix.save_h5_file (cellranger/matrix.py:415-447) + FeatureReference.to_hdf5 (feature_ref.py:103-118)
    + atac from_peaks_bed (cellranger/atac/feature_ref.py:45-71: id = name = 'chr:start-end', type 'Peaks',
    tags genome / derivation='').  compression="gzip" (default): the four matrix arrays chunked / shuffled / deflated as
    save_h5_group does; None: contiguous."""
    f = open(path, "wb")
    w = _Writer(f)
    arr = w.dataset_chunked if compression == "gzip" else w.dataset
    w.alloc(b"\0" * 96)                                                      # superblock placeholder
    n_bc = len(barcodes)
    feats = {
        "_all_tag_keys": w.dataset(_string_array(["genome", "derivation"])),
        "derivation": w.dataset(_string_array([""] * n_features)),
        "feature_type": w.dataset(_string_array(["Peaks"] * n_features)),
        "genome": w.dataset(_string_array(list(genome) if isinstance(genome, (list, tuple)) else [genome] * n_features)),
        "id": w.dataset(_string_array(feature_ids)),
        "name": w.dataset(_string_array(feature_ids)),
    }
    feat_hdr, _, _ = w.group(feats)
    matrix = {
        "barcodes": w.dataset(_string_array(barcodes)),
        "data": arr(np.asarray(data, dtype="<i4")),
        "indices": arr(np.asarray(indices, dtype="<i8")),
        "indptr": arr(np.asarray(indptr, dtype="<i8")),
        "shape": arr(np.array([n_features, n_bc], dtype="<i4")),
        "features": feat_hdr,
    }
    mat_hdr, _, _ = w.group(matrix)
    root_hdr, root_tree, root_heap = w.group({"matrix": mat_hdr},
                                             attrs={"filetype": "matrix", "version": 2, "software_version": sw_version or ""})
    pad = (-w.pos) % 8
    if pad:
        f.write(b"\0" * pad)
    eof = w.pos + pad
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_tree, root_heap)   # root symbol table entry
    assert len(sb) == 96
    f.seek(0)
    f.write(sb)
    f.close()

# --------------------------------------------------------------------------------------------- reader (same subset)
class _Reader(object):
    def __init__(self, blob):
        self.b = blob
        assert blob[:8] == SIGNATURE and blob[8] == 0, "not a version-0 HDF5 superblock"
        self.root_hdr = struct.unpack_from("<Q", blob, 24 + 32 + 8)[0]

    def messages(self, addr):
        ver, _, n, _, size = struct.unpack_from("<BBHII", self.b, addr)
        if ver != 1:
            raise ValueError("object header version %d at %d: only version-1 headers (libver='earliest') are read" % (ver, addr))
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n:
            p, left = blocks.pop(0)
            while left >= 8 and len(out) < n:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.b, p)
                body = self.b[p + 8:p + 8 + msize]
                if mtype == 0x0010:                                            # continuation: more messages elsewhere
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                out.append((mtype, body))
                p += 8 + msize
                left -= 8 + msize
        return out

    def children(self, hdr):
        for mtype, body in self.messages(hdr):
            if mtype == 0x0011:
                tree, heap = struct.unpack_from("<QQ", body, 0)
                seg_addr = struct.unpack_from("<Q", self.b, heap + 24)[0]
                assert self.b[tree:tree + 4] == b"TREE"
                n_ent = struct.unpack_from("<H", self.b, tree + 6)[0]
                out = {}
                for e in range(n_ent):
                    snod = struct.unpack_from("<Q", self.b, tree + 24 + 8 + 16 * e)[0]
                    assert self.b[snod:snod + 4] == b"SNOD"
                    n_sym = struct.unpack_from("<H", self.b, snod + 6)[0]
                    for k in range(n_sym):
                        name_off, obj = struct.unpack_from("<QQ", self.b, snod + 8 + 40 * k)
                        end = self.b.index(b"\0", seg_addr + name_off)
                        out[self.b[seg_addr + name_off:end].decode("ascii")] = obj
                return out
        return None

    def attrs(self, hdr):
        out = {}
        for mtype, body in self.messages(hdr):
            if mtype == 0x000C:
                _, _, nlen, dlen, slen = struct.unpack_from("<BBHHH", body, 0)
                p = 8
                name = body[p:p + nlen].rstrip(b"\0").decode("ascii")
                p += (nlen + 7) // 8 * 8
                dt = body[p:p + dlen]
                p += (dlen + 7) // 8 * 8 + (slen + 7) // 8 * 8
                cls = dt[0] & 0x0F
                size = struct.unpack_from("<I", dt, 4)[0]
                raw = body[p:p + size]
                out[name] = struct.unpack("<q", raw)[0] if cls == 0 else raw.rstrip(b"\0").decode("ascii")
        return out

    def _chunks(self, node, out):
        """(element offset, address, filtered size, filter mask) of every chunk below B-tree node `node` (1-D datasets)."""
        if self.b[node:node + 4] != b"TREE" or self.b[node + 4] != 1:
            raise ValueError("no chunk B-tree node at %d" % node)
        level, used = struct.unpack_from("<BH", self.b, node + 5)
        p = node + 24
        for _ in range(used):
            size, mask, off, _zero, child = struct.unpack_from("<IIQQQ", self.b, p)
            if level == 0:
                out.append((off, child, size, mask))
            else:
                self._chunks(child, out)
            p += 32

    def dataset(self, hdr):
        shape, dtype, layout, filters = (), None, None, []
        for mtype, body in self.messages(hdr):
            if mtype == 0x0001:
                if body[0] != 1:
                    raise ValueError("dataspace message version %d is not read" % body[0])
                rank = body[1]
                shape = struct.unpack_from("<%dQ" % rank, body, 8) if rank else ()
            elif mtype == 0x0003:
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                if cls == 0:
                    if body[1] & 1:
                        raise ValueError("big-endian integers are not read")
                    dtype = np.dtype("<%s%d" % ("i" if body[1] & 8 else "u", size))
                elif cls == 3:
                    dtype = np.dtype("S%d" % size)
                else:
                    raise ValueError("datatype class %d is not read (fixed-point and fixed-length strings only)" % cls)
            elif mtype == 0x000B:
                if body[0] != 1:
                    raise ValueError("filter pipeline version %d is not read" % body[0])
                p = 8
                for _ in range(body[1]):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", body, p)
                    p += 8 + (nlen + 7) // 8 * 8
                    cd = struct.unpack_from("<%dI" % ncd, body, p)
                    p += 4 * ncd + (4 if ncd % 2 else 0)
                    filters.append((fid, cd))
            elif mtype == 0x0008:
                if body[0] != 3:
                    raise ValueError("data layout message version %d is not read" % body[0])
                if body[1] == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", body, 2)
                elif body[1] == 2:
                    ndim = body[2]
                    if ndim != 2:
                        raise ValueError("chunked datasets of rank %d are not read (1-D only)" % (ndim - 1))
                    layout = ("chunked", struct.unpack_from("<Q", body, 3)[0]) + struct.unpack_from("<II", body, 11)
                elif body[1] == 0:
                    n = struct.unpack_from("<H", body, 2)[0]
                    layout = ("compact", bytes(body[4:4 + n]))
                else:
                    raise ValueError("data layout class %d is not read" % body[1])
        if layout is None or dtype is None:
            raise ValueError("dataset header at %d lacks a datatype or layout message" % hdr)
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            if layout[2] == 0 or layout[1] == UNDEF:
                return np.zeros(shape, dtype=dtype)
            return np.frombuffer(self.b, dtype=dtype, count=count, offset=layout[1]).reshape(shape).copy()
        if layout[0] == "compact":
            return np.frombuffer(layout[1], dtype=dtype, count=count).reshape(shape).copy()
        _, tree, chunk, item = layout
        out = np.zeros(count, dtype=dtype)
        if tree == UNDEF or count == 0:
            return out.reshape(shape)
        chunks = []
        self._chunks(tree, chunks)
        for off, addr, size, mask in chunks:
            raw = bytes(self.b[addr:addr + size])
            for k, (fid, cd) in reversed(list(enumerate(filters))):           # undo the pipeline back to front
                if mask & (1 << k):
                    continue                                                   # this filter was skipped for this chunk
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    raw = _unshuffle(raw, cd[0] if cd else item)
                else:
                    raise ValueError("filter %d is not read (deflate and shuffle only)" % fid)
            vals = np.frombuffer(raw, dtype=dtype)
            take = min(vals.size, count - off)
            if take > 0:
                out[off:off + take] = vals[:take]
        return out.reshape(shape)


def _read_with_h5py(path):
    import h5py
    with h5py.File(path, "r") as f:
        attrs = {}
        for k, v in f.attrs.items():
            attrs[k] = v.decode() if isinstance(v, bytes) else (int(v) if np.issubdtype(type(v), np.integer) else v)
        m = f["matrix"]
        out = dict(attrs=attrs, features={k: m["features"][k][()] for k in m["features"]})
        for k in ("indptr", "indices", "data", "shape", "barcodes"):
            out[k] = m[k][()]
    return out


def read_matrix_h5(path):
    """-> dict(attrs, indptr, indices, data, shape, barcodes, features={name: array}).  Where h5py is installed (any
    cellranger installation) it does the reading; otherwise the reader above, which raises ValueError with the reason on
    a file outside its subset."""
    try:
        return _read_with_h5py(path)
    except ImportError:
        pass
    with open(path, "rb") as f:
        r = _Reader(f.read())
    root = r.children(r.root_hdr)
    attrs = r.attrs(r.root_hdr)
    m = r.children(root["matrix"])
    feats = r.children(m["features"])
    out = dict(attrs=attrs, features={k: r.dataset(v) for k, v in feats.items()})
    for k in ("indptr", "indices", "data", "shape", "barcodes"):
        out[k] = r.dataset(m[k])
    return out
