#!/usr/bin/env python
"""Decode the reference's vignette input (inst/extdata/cellSNP.cells.vcf.gz) into tests/golden/vignette_cellsnp.npz
without R: the AD and DP FORMAT fields per cell, exactly what load_cellSNP_vcf() extracts with vcfR
(R/read_data.R:233-265 with its defaults min_count = 0, min_MAF = 0: every SNP whose total DP is positive passes;
missing calls become 0), and the "full" row names chr_pos_ref_alt (:268-281).  Run in the build container, where
/root/reference exists; the fixture is committed so that the GPU box never needs the reference."""
import gzip
import os
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/inst/extdata/cellSNP.cells.vcf.gz"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "vignette_cellsnp.npz")


def main():
    cells, names, A, D = None, [], [], []
    with gzip.open(SRC, "rt") as f:
        for line in f:
            if line.startswith("##"):
                continue
            col = line.rstrip("\n").split("\t")
            if line.startswith("#CHROM"):
                cells = col[9:]
                continue
            fmt = col[8].split(":")
            ia, idp = fmt.index("AD"), fmt.index("DP")
            a_row, d_row = [], []
            for g in col[9:]:
                parts = g.split(":")
                a_row.append(0.0 if parts[ia] in (".", "") else float(parts[ia]))
                d_row.append(0.0 if parts[idp] in (".", "") else float(parts[idp]))
            names.append("%s_%s_%s_%s" % (col[0], col[1], col[3], col[4]))
            A.append(a_row)
            D.append(d_row)
    A, D = np.array(A), np.array(D)
    with np.errstate(invalid="ignore", divide="ignore"):
        keep = (D.sum(1) >= 0) & ((A.sum(1) / D.sum(1)) >= 0)      # min_count = 0, min_MAF = 0; NaN (no reads) fails
    A, D, names = A[keep], D[keep], [n for n, k in zip(names, keep) if k]
    np.savez_compressed(OUT, A=A, D=D, variants=np.array(names), cells=np.array(cells))
    print("%d out of %d SNPs passed; %d cells; %d covered entries -> %s" % (keep.sum(), keep.size, A.shape[1],
                                                                          int((D > 0).sum()), OUT))


if __name__ == "__main__":
    main()
