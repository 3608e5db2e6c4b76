"""`stencil` -- apply GBRS genotype calls to a multiway alignment incidence matrix and save the restricted matrix.

Same call surface and output as the reference's `stencil` (/root/reference/src/gbrs/gbrs/emase_utils.py:110-177): the
genotype TSV (`gene<TAB>diplotype`, `#` comments skipped) becomes an H x T 0/1 mask -- 1 for the haplotypes named by the
diplotype letters on every transcript of the gene -- the incidence is multiplied by it, structural zeros are dropped
and the matrix is written back in EMASE format.  It is the same restriction `quantify -G` applies before the EM
(emase_utils.py:240-273), as a file-to-file step.

Host only: this is a data-format step next to the path (sparse masking of the per-haplotype CSC matrices), there is no
kernel behind it.  Two quirks of the reference are not reproduced: with `output_file=None` it computes a default name
into an unused variable and then fails in `save(h5file=None)` (:131-132, :175) -- here the default name
`gbrs.stenciled.<alignment file name>` is used; and its no-group branch indexes the mask with gene ids that do not exist
without a group file (:151, :169) -- here, without groups, the genotype file is keyed by locus name ("stenciled as is").
"""
from __future__ import annotations

import os
from itertools import dropwhile

import numpy as np

from . import utils
from .apm import AlignmentPropertyMatrix
from .quantify import load_genotype_mask

logger = utils.get_logger("gbrs")


def stencil(alignment_file: str, genotype_file: str, group_file: str = None, output_file: str = None) -> None:
    """Applying genotype calls to multi-way alignment incidence matrix (emase_utils.py:110-177)."""
    data_dir = os.getenv("GBRS_DATA", ".")
    if group_file is None:
        group_file = os.path.join(data_dir, "ref.gene2transcripts.tsv")  # :125-128
        if not os.path.exists(group_file):
            logger.info("A group file is *not* given. Genotype will be stenciled as is.")
            group_file = None
    if output_file is None:
        output_file = f"gbrs.stenciled.{os.path.basename(alignment_file)}"
    logger.info(f"Alignment File: {alignment_file}")
    logger.info(f"Genotype File: {genotype_file}")
    logger.info(f"Group File: {group_file}")
    logger.info(f"Output File: {output_file}")

    logger.info(f"Loading EMASE file: {alignment_file}")
    aln_mat = AlignmentPropertyMatrix(h5file=alignment_file, grpfile=group_file)
    logger.debug(f"Number Loci: {aln_mat.num_loci}")
    logger.debug(f"Number Haplotypes: {aln_mat.num_haplotypes}")
    logger.debug(f"Number Reads: {aln_mat.num_reads}")

    logger.info(f"Loading and processing genotype calls from: {genotype_file}")
    if group_file is not None:
        gtmask, _, _ = load_genotype_mask(aln_mat, genotype_file)  # :147-161, shared with quantify -G
    else:
        hid = dict(zip(aln_mat.hname, np.arange(aln_mat.num_haplotypes)))
        gtmask = np.zeros((aln_mat.num_haplotypes, aln_mat.num_loci))
        with open(genotype_file) as fh:
            for line in dropwhile(utils.is_comment, fh):
                item = line.rstrip().split("\t")
                t, gt = item[:2]
                gtmask[np.array([hid[c] for c in gt]), aln_mat.lid[t]] = 1.0

    aln_mat.multiply(gtmask, axis=2)  # :171-173
    aln_mat.eliminate_zeros()
    logger.info(f"Saving EMASE Formatted File: {output_file}")
    aln_mat.save(h5file=output_file)
    logger.info("Done")
