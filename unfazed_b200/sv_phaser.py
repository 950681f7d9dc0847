"""Drop-in for the reference's ``unfazed/sv_phaser.py``: ``phase_svs`` (:427-448) -- CNV
allele-balance phasing over the sites inside DEL/DUP events (``run_cnv_phasing`` :357-423,
``phase_by_snvs`` :71-85) plus read-backed phasing from breakpoint-supporting reads
(``run_read_phasing`` :176-266), merged as :484-492 -- batched on the GPU."""
from __future__ import annotations

from . import snv_phaser
from .snv_phaser import phase_by_reads  # noqa: F401  (sv_phaser.py:14-68 is the same function)


def phase_by_snvs(informative_sites):
    """``sv_phaser.py:71-85``: every CNV candidate site votes for ``site[site["kid_allele"]]``; returns
    ``{parent: [site, ...]}`` keyed by the parents of the FIRST site, or None without sites.  Host mirror of the
    vote count ``unfz_compact_sites`` keeps per DNM."""
    if not informative_sites:
        return None
    votes = {informative_sites[0]["ref_parent"]: [], informative_sites[0]["alt_parent"]: []}
    for site in informative_sites:
        votes[site[site["kid_allele"]]].append(site)
    return votes

def autophase(denovo, pedigrees, records, dad_id, mom_id, build):
    """``sv_phaser.py:304-354``: as ``snv_phaser.autophase`` -- the SEX-CHROM record is written -- but the reference's SV
    variant falls off its end and returns None where the SNV one returns True (False on the early exits), so its caller
    goes on to phase the event as well."""
    return None if snv_phaser.autophase(denovo, pedigrees, records, dad_id, mom_id, build) else False


def phase_svs(dnms, kids, pedigrees, sites, threads, build, no_extended, multiread_proc_min, quiet_mode,
              ab_homref, ab_homalt, ab_het, min_gt_qual, min_depth, search_dist, insert_size_max_sample,
              stdevs, min_map_qual, readlen, split_error_margin):
    snv_phaser.QUIET_MODE = quiet_mode
    return snv_phaser.run_batch([], dnms, pedigrees, sites, threads, build, no_extended, multiread_proc_min,
                                ab_homref, ab_homalt, ab_het, min_gt_qual, min_depth, search_dist,
                                insert_size_max_sample, stdevs, min_map_qual, readlen, split_error_margin)
