"""Tabix index (.tbi) of fragments.tsv.gz: where each contig's records sit in the BGZF file.

The reference reads one contig at a time through ``pysam.TabixFile(filename).fetch(contig)``
(lib/python/tools/io.py:82-93); a rank of the multi-GPU step owns a contiguous range of contigs, so it needs the
BGZF virtual-offset span of that range and inflates only those blocks (``cratac_inflate_range``).

File layout (tabix format, Li 2011 / htslib ``tbx.c``; every integer little endian; the file itself is BGZF):

    magic "TBI\\1" | n_ref | format | col_seq | col_beg | col_end | meta | skip | l_nm | names (NUL terminated)
    per reference:  n_bin, then per bin: bin (u32), n_chunk, n_chunk x (cnk_beg u64, cnk_end u64)
                    n_intv, n_intv x ioff (u64)          -- linear index, one entry per 16 kb
    optional n_no_coor (u64)

A virtual offset is ``compressed offset of the BGZF block << 16 | offset inside the inflated block``.  Bin 37450 is
htslib's pseudo-bin (record counts), not a set of chunks.  pysam / htslib are not in this image, so the reader is
checked against indexes written by ``build_index`` below (same binning scheme: min_shift 14, depth 5) -- not against
an htslib-written file.
"""
import struct

from . import bgzf

MAGIC = b"TBI\x01"
PSEUDO_BIN = 37450
TBX_UCSC = 0x10000          # 0-based half-open coordinates (the "bed" preset: -s1 -b2 -e3)


def _read_bgzf(path):
    with open(path, "rb") as f:
        return bgzf.decompress(f.read())


def read_index(path):
    """-> dict(names=[...], format, col_seq, col_beg, col_end, meta, skip,
                refs=[dict(bins={bin: [(beg, end), ...]}, linear=[ioff, ...]), ...])."""
    raw = _read_bgzf(path)
    if raw[:4] != MAGIC:
        raise ValueError("%s is not a tabix index" % path)
    n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<8i", raw, 4)
    pos = 36
    names = raw[pos:pos + l_nm].split(b"\x00")[:n_ref]
    pos += l_nm
    refs = []
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", raw, pos)
        pos += 4
        bins = {}
        for _b in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", raw, pos)
            pos += 8
            chunks = [struct.unpack_from("<QQ", raw, pos + 16 * k) for k in range(n_chunk)]
            pos += 16 * n_chunk
            bins[b] = chunks
        (n_intv,) = struct.unpack_from("<i", raw, pos)
        pos += 4
        linear = list(struct.unpack_from("<%dQ" % n_intv, raw, pos))
        pos += 8 * n_intv
        refs.append(dict(bins=bins, linear=linear))
    return dict(names=[n.decode() for n in names], format=fmt, col_seq=col_seq, col_beg=col_beg, col_end=col_end,
                meta=meta, skip=skip, refs=refs)


def contig_spans(index):
    """-> {contig: (voff_begin, voff_end)}: the span that holds every record of the contig (contigs without
    records are absent).  fragments.tsv.gz is sorted by contig, so the spans of different contigs do not overlap."""
    spans = {}
    for name, ref in zip(index["names"], index["refs"]):
        chunks = [c for b, cs in ref["bins"].items() if b != PSEUDO_BIN for c in cs]
        if chunks:
            spans[name] = (min(c[0] for c in chunks), max(c[1] for c in chunks))
    return spans


def range_span(index, contigs):
    """Virtual-offset span of a set of contigs that are consecutive in the file -> (begin, end), or None when none
    of them has a record."""
    spans = contig_spans(index)
    have = [spans[c] for c in contigs if c in spans]
    if not have:
        return None
    return min(s[0] for s in have), max(s[1] for s in have)


# ------------------------------------------------------------------------------------------ writer (synthetic inputs)
def reg2bin(beg, end, min_shift=14, depth=5):
    """UCSC binning scheme of tabix / BAM: smallest bin that contains [beg, end)."""
    end -= 1
    s, t = min_shift, ((1 << depth * 3) - 1) // 7
    for level in range(depth, 0, -1):
        if beg >> s == end >> s:
            return t + (beg >> s)
        s += 3
        t -= 1 << (level - 1) * 3
    return 0


def _block_offsets(blob):
    """Compressed offsets and inflated sizes of the blocks of a BGZF file."""
    offs, sizes = [], []
    p = 0
    while p < len(blob):
        xlen = struct.unpack_from("<H", blob, p + 10)[0]
        q, bsize = p + 12, None
        while q < p + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", blob, q)
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", blob, q + 4)[0] + 1
            q += 4 + slen
        if bsize is None:
            raise ValueError("not a BGZF file")
        offs.append(p)
        sizes.append(struct.unpack_from("<I", blob, p + bsize - 4)[0])
        p += bsize
    return offs, sizes


def build_index(bgzf_path, index_path=None):
    """Write the .tbi of a sorted BGZF fragments file ("bed" preset).  Python loops over every line: for tests and
    synthetic inputs, not for production-size files (those come with the index cellranger-atac wrote)."""
    import bisect
    with open(bgzf_path, "rb") as f:
        blob = f.read()
    offs, sizes = _block_offsets(blob)
    starts = [0]
    for s in sizes:
        starts.append(starts[-1] + s)          # inflated offset at which each block begins
    text = bgzf.decompress(blob)

    def voff(upos):
        k = bisect.bisect_right(starts, upos) - 1
        while k < len(sizes) and sizes[k] == 0:          # never point into an empty block
            k += 1
        if k >= len(offs):
            return len(blob) << 16
        if upos - starts[k] == sizes[k] and k + 1 < len(offs):   # end of a block = start of the next one
            k += 1
        return (offs[k] << 16) | (upos - starts[k])

    names, refs = [], []
    pos = 0
    cur = None
    n = len(text)
    while pos < n:
        nl = text.find(b"\n", pos)
        if nl < 0:
            nl = n
        line = text[pos:nl]
        nxt = min(nl + 1, n)
        if line and not line.startswith(b"#"):
            f = line.split(b"\t")
            name, beg, end = f[0].decode(), int(f[1]), int(f[2])
            if cur is None or names[-1] != name:
                if name in names:
                    raise ValueError("fragments are not sorted by contig")
                names.append(name)
                cur = dict(bins={}, linear=[], last_bin=None)
                refs.append(cur)
            v0, v1 = voff(pos), voff(nxt)
            b = reg2bin(beg, max(end, beg + 1))
            chunks = cur["bins"].setdefault(b, [])
            if cur["last_bin"] == b and chunks:
                chunks[-1] = (chunks[-1][0], v1)           # the run of records in this bin goes on
            else:
                chunks.append((v0, v1))
            cur["last_bin"] = b
            lin = cur["linear"]
            for w in range(beg >> 14, ((max(end, beg + 1) - 1) >> 14) + 1):
                while len(lin) <= w:
                    lin.append(0)
                if lin[w] == 0:
                    lin[w] = v0
        pos = nxt
    out = [MAGIC]
    nm = b"".join(x.encode() + b"\x00" for x in names)
    out.append(struct.pack("<8i", len(names), TBX_UCSC, 1, 2, 3, ord("#"), 0, len(nm)))
    out.append(nm)
    for ref in refs:
        lin = ref["linear"]
        for w in range(1, len(lin)):                        # htslib fills empty windows from the left
            if lin[w] == 0:
                lin[w] = lin[w - 1]
        out.append(struct.pack("<i", len(ref["bins"])))
        for b in sorted(ref["bins"]):
            chunks = ref["bins"][b]
            out.append(struct.pack("<Ii", b, len(chunks)))
            for c in chunks:
                out.append(struct.pack("<QQ", c[0], c[1]))
        out.append(struct.pack("<i", len(lin)))
        out.append(struct.pack("<%dQ" % len(lin), *lin))
    bgzf.write(index_path or bgzf_path + ".tbi", b"".join(out), level=6)
    return index_path or bgzf_path + ".tbi"
