"""Constants the drop-in modules share.  Same names and values as the reference's
``unfazed/utils.py`` (genotype codes :2-5, SEX_KEY :6, type lists :7-9, CIGAR map :13-24, PAR tables
:26-43, ``get_prefix`` :46-52) because callers do ``from .utils import *``.  The pseudoautosomal
bounds are kept exactly as the reference spells them, swapped build labels included;
``plan.py`` bakes the same numbers in."""

HOM_REF, HET, GT_UNKNOWN, HOM_ALT = range(4)
SEX_KEY = dict(male=1, female=2)
VCF_TYPES = "vcf vcf.gz bcf".split()
SV_TYPES = "DEL DUP INV CNV DUP:TANDEM DEL:ME CPX CTX".split()
SNV_TYPES = "POINT SNV INDEL".split()
LABELS = "chrom start end kid vartype".split()
QUIET_MODE = False
CIGAR_MAP = {code: letter for code, letter in enumerate("MIDNSHP=XB")}   # BAM op code -> letter


def _par(x_lo, x_hi, y_lo, y_hi):
    return {"x": [x_lo, x_hi], "y": [y_lo, y_hi]}


grch37_par1 = _par(10001, 2781479, 10001, 2781479)
grch37_par2 = _par(155701383, 156030895, 56887903, 57217415)
grch38_par1 = _par(60001, 2699520, 10001, 2649520)
grch38_par2 = _par(154931044, 155260560, 59034050, 59363566)


def get_prefix(vcf):
    """First three characters of the first record's CHROM when it spells "chr" (any case), else ""."""
    first = next(iter(vcf), None)
    if first is None or "chr" not in first.CHROM.lower():
        return ""
    return first.CHROM[:3]
