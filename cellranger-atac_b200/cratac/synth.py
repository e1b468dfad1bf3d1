"""Seeded synthetic fragments: GRCh38 / mm10 shaped inputs.

Fragments: fraction 0.5 start near a latent peak centre (N(0,150)), 0.5 uniform; length from
the nucleosomal mixture 0.50 N(120,35) + 0.35 N(320,45) + 0.15 N(520,60) clipped to [20,1200];
0 <= start < stop <= contig_len; dup_count = 1 + Geometric(0.5); barcodes: `cells` random
16-mers + "-1" with log-normal(0.5) depth carrying 80 % of fragments plus 20 x cells
background barcodes sharing 20 %.  Sorted by (contig order, start, stop).
"""
import numpy as np

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
          ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
          ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
          ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]
MM10 = [("chr1", 195471971), ("chr2", 182113224), ("chr3", 160039680), ("chr4", 156508116), ("chr5", 151834684),
        ("chr6", 149736546), ("chr7", 145441459), ("chr8", 129401213), ("chr9", 124595110), ("chr10", 130694993),
        ("chr11", 122082543), ("chr12", 120129022), ("chr13", 120421639), ("chr14", 124902244), ("chr15", 104043685),
        ("chr16", 98207768), ("chr17", 94987271), ("chr18", 90702639), ("chr19", 61431566), ("chrX", 171031299),
        ("chrY", 91744698)]

_BASES = np.frombuffer(b"ACTG", dtype=np.uint8)   # 2-bit code order used by libcratac ((ch >> 1) & 3)


def seq_to_str(seq):
    """uint32 2-bit packed 16-mer -> str (first base in the top bits)."""
    shifts = np.arange(30, -2, -2, dtype=np.uint32)
    codes = (np.asarray(seq, dtype=np.uint32)[..., None] >> shifts) & 3
    return _BASES[codes]


def generate(contigs, n_fragments, cells, n_peaks, seed, background_factor=20, in_peak_frac=0.5):
    """-> dict(chrom, start, end, dup (int32), bc_seq (uint32 per fragment), n_contigs)."""
    rng = np.random.default_rng(seed)
    lens = np.array([c[1] for c in contigs], dtype=np.int64)
    cum = np.concatenate(([0], np.cumsum(lens)))
    total = int(cum[-1])
    n = int(n_fragments)
    # latent peaks
    centres = np.sort(rng.integers(0, total, size=max(int(n_peaks), 1)))
    in_peak = rng.random(n) < in_peak_frac
    gpos = np.where(in_peak, centres[rng.integers(0, centres.size, size=n)] + rng.normal(0, 150, size=n).astype(np.int64),
                    rng.integers(0, total, size=n))
    gpos = np.clip(gpos, 0, total - 1)
    chrom = (np.searchsorted(cum, gpos, side="right") - 1).astype(np.int32)
    start = (gpos - cum[chrom]).astype(np.int64)
    comp = rng.random(n)
    length = np.where(comp < 0.5, rng.normal(120, 35, size=n), np.where(comp < 0.85, rng.normal(320, 45, size=n), rng.normal(520, 60, size=n)))
    length = np.clip(length.astype(np.int64), 20, 1200)
    L = lens[chrom]
    start = np.minimum(start, L - 1)
    end = np.minimum(start + length, L)
    start = np.minimum(start, end - 1)
    start = np.maximum(start, 0)
    dup = rng.geometric(0.5, size=n).astype(np.int32)
    # barcodes
    n_bg = background_factor * cells
    seqs = rng.integers(0, 1 << 32, size=cells + n_bg, dtype=np.uint64).astype(np.uint32)
    depth = rng.lognormal(0.0, 0.5, size=cells)
    p_cell = 0.8 * depth / depth.sum()
    p = np.concatenate((p_cell, np.full(n_bg, 0.2 / max(n_bg, 1)))) if n_bg else depth / depth.sum()
    bc_idx = rng.choice(cells + n_bg, size=n, p=p / p.sum())
    order = np.lexsort((end, start, chrom))
    return dict(chrom=chrom[order], start=start[order].astype(np.int32), end=end[order].astype(np.int32), dup=dup[order],
                bc_seq=seqs[bc_idx][order], contigs=contigs)


def to_text(frags, gem_group="1"):
    """SoA -> fragments.tsv bytes (vectorised; one np.char pass per column)."""
    names = np.array([c[0] for c in frags["contigs"]])
    n = frags["chrom"].size
    if n == 0:
        return b""
    bc = seq_to_str(frags["bc_seq"]).view("S16").ravel().astype("U16")
    cols = [names[frags["chrom"]], frags["start"].astype("U"), frags["end"].astype("U"),
            np.char.add(bc, "-" + gem_group), frags["dup"].astype("U")]
    line = cols[0]
    for c in cols[1:]:
        line = np.char.add(np.char.add(line, "\t"), c)
    return ("\n".join(line.tolist()) + "\n").encode("ascii")


def barcode_strings(frags, gem_group="1"):
    return np.char.add(seq_to_str(frags["bc_seq"]).view("S16").ravel().astype("U16"), "-" + gem_group)
