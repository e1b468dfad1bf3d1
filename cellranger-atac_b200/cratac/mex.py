"""Matrix Market export of the raw peak x barcode matrix: `CountMatrix.save_mex` (cellranger/matrix.py:768-823)
with the ATAC feature writer `save_features_bed` (cellranger/atac/feature_ref.py:15-21), uncompressed as
`atac.matrix.save_mex` calls it (cellranger/atac/matrix.py:11-33)."""
import json
import os

import numpy as np

from . import _ffi


def save_mex(base_dir, indptr, indices, data, n_features, barcodes, feature_ids, sw_version):
    if not os.path.isdir(base_dir):
        os.makedirs(base_dir)
    n_bc = len(barcodes)
    nnz = int(indptr[-1]) if len(indptr) else 0
    meta = json.dumps({"software_version": sw_version, "format_version": 2})
    header = ("%%MatrixMarket matrix coordinate integer general\n" + "%%metadata_json: %s\n" % meta + "%i %i %i\n" % (n_features, n_bc, nnz))
    # the entries ("row col value", 1-based, CSC -> COO order) are formatted natively on all cores (cratac_write_mtx)
    _ffi.write_mtx(os.path.join(base_dir, "matrix.mtx"), header, np.asarray(indptr, dtype=np.int64) if len(indptr) else np.zeros(1, dtype=np.int64),
                   indices, data)
    with open(os.path.join(base_dir, "peaks.bed"), "w") as f:
        for name in feature_ids:
            contig, span = name.split(":")
            s, e = span.split("-")
            f.write(contig + "\t" + s + "\t" + e + "\n")
    with open(os.path.join(base_dir, "barcodes.tsv"), "w") as f:
        for bc in barcodes:
            f.write(bc + "\n")
