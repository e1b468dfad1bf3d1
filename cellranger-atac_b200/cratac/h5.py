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

Chunking/gzip/shuffle (matrix.py:441-447) are storage details no reader depends on; datasets are contiguous.
The reader below understands exactly this subset and is what the tests use for round trips; compatibility
with libhdf5 itself could not be exercised in the build image (stated in DESIGN.md).
"""
import struct

import numpy as np

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


def _dataspace_msg(shape):
    # version 1: version, rank, flags, reserved(1), reserved(4), dims
    body = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0)
    for d in shape:
        body += struct.pack("<Q", int(d))
    return body


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


def write_matrix_h5(path, indptr, indices, data, n_features, barcodes, feature_ids, genome, sw_version=""):
    """CountMatrix.save_h5_file (cellranger/matrix.py:415-447) + FeatureReference.to_hdf5 (feature_ref.py:103-118)
    + atac from_peaks_bed (cellranger/atac/feature_ref.py:45-71: id = name = 'chr:start-end', type 'Peaks',
    tags genome / derivation='')."""
    f = open(path, "wb")
    w = _Writer(f)
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
        "data": w.dataset(np.asarray(data, dtype="<i4")),
        "indices": w.dataset(np.asarray(indices, dtype="<i8")),
        "indptr": w.dataset(np.asarray(indptr, dtype="<i8")),
        "shape": w.dataset(np.array([n_features, n_bc], dtype="<i4")),
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
        assert ver == 1
        p = addr + 16
        out = []
        for _ in range(n):
            mtype, msize, _flags = struct.unpack_from("<HHB", self.b, p)
            out.append((mtype, self.b[p + 8:p + 8 + msize]))
            p += 8 + msize
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

    def dataset(self, hdr):
        shape, dtype, addr, nbytes = (), None, UNDEF, 0
        for mtype, body in self.messages(hdr):
            if mtype == 0x0001:
                rank = body[1]
                shape = struct.unpack_from("<%dQ" % rank, body, 8) if rank else ()
            elif mtype == 0x0003:
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                dtype = np.dtype("<i%d" % size) if cls == 0 else np.dtype("S%d" % size)
            elif mtype == 0x0008:
                addr, nbytes = struct.unpack_from("<QQ", body, 2)
        if nbytes == 0:
            return np.zeros(shape, dtype=dtype)
        return np.frombuffer(self.b, dtype=dtype, count=int(np.prod(shape)), offset=addr).reshape(shape).copy()


def read_matrix_h5(path):
    """-> dict(attrs, indptr, indices, data, shape, barcodes, features={name: array})."""
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
