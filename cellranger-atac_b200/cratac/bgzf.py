"""Minimal BGZF writer/reader (the container of fragments.tsv.gz; SAM spec section 4.1).

The build image has no pysam/bgzip, so synthetic inputs are written with zlib by hand: a
series of gzip members of <= 64 KiB, each carrying its compressed size in a 'BC' extra
field, closed by the 28-byte empty EOF block.  Any gzip reader (and ``gzip -c -d`` as used
by lib/python/tools/io.py:197-217 of the reference) can read the result.
"""
import struct
import zlib

BLOCK_IN = 0xff00
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _block(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    comp = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    assert bsize < 65536
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    tail = struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))
    return head + body + tail


def compress(data, level=1, strategy=zlib.Z_DEFAULT_STRATEGY, block_in=BLOCK_IN):
    """bytes -> BGZF bytes."""
    mv = memoryview(data)
    out = [_block(bytes(mv[i:i + block_in]), level, strategy) for i in range(0, len(mv), block_in)]
    out.append(EOF_BLOCK)
    return b"".join(out)


def write(path, data, level=1):
    with open(path, "wb") as f:
        f.write(compress(data, level))


def write_parallel(path, data, level=1, threads=None, batch=256):
    """Same file as write(), compressed on `threads` threads (zlib releases the GIL): batches of `batch` blocks per
    task, written in order.  For the multi-GB synthetic inputs of bench.py."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    mv = memoryview(data)
    n = len(mv)
    step = BLOCK_IN * batch
    threads = threads or (os.cpu_count() or 1)

    def task(i):
        return b"".join(_block(bytes(mv[j:j + BLOCK_IN]), level) for j in range(i, min(n, i + step), BLOCK_IN))

    with open(path, "wb") as f, ThreadPoolExecutor(threads) as pool:
        starts = list(range(0, n, step))
        for k in range(0, len(starts), 4 * threads):            # bounded number of compressed batches in memory
            for blob in pool.map(task, starts[k:k + 4 * threads]):
                f.write(blob)
        f.write(EOF_BLOCK)


def decompress(blob):
    """BGZF / multi-member gzip bytes -> bytes (host zlib; tests only, the product uses cratac_inflate_file)."""
    out = []
    pos = 0
    while pos < len(blob):
        d = zlib.decompressobj(31)
        out.append(d.decompress(blob[pos:]))
        pos = len(blob) - len(d.unused_data)
    return b"".join(out)
