"""Constants shared by the drop-in modules (mirror of the reference's ``unfazed/utils.py``:
genotype codes :2-5, SEX_KEY :6, variant type lists :7-9, CIGAR map :13-24, PAR tables :26-43,
``get_prefix`` :46-52).  The PAR tables are reproduced exactly as the reference spells them,
including the swapped build labels -- they are also baked into ``plan.py``."""

HOM_REF, HET, GT_UNKNOWN, HOM_ALT = 0, 1, 2, 3
SEX_KEY = {"male": 1, "female": 2}
VCF_TYPES = ["vcf", "vcf.gz", "bcf"]
SV_TYPES = ["DEL", "DUP", "INV", "CNV", "DUP:TANDEM", "DEL:ME", "CPX", "CTX"]
SNV_TYPES = ["POINT", "SNV", "INDEL"]
LABELS = ["chrom", "start", "end", "kid", "vartype"]
QUIET_MODE = False
CIGAR_MAP = dict(enumerate("MIDNSHP=XB"))

grch37_par1 = {"x": [10001, 2781479], "y": [10001, 2781479]}
grch37_par2 = {"x": [155701383, 156030895], "y": [56887903, 57217415]}
grch38_par1 = {"x": [60001, 2699520], "y": [10001, 2649520]}
grch38_par2 = {"x": [154931044, 155260560], "y": [59034050, 59363566]}


def get_prefix(vcf):
    """"chr"-style prefix of the first record the handle yields, "" when it yields nothing."""
    for var in vcf:
        return var.CHROM[:3] if "chr" in var.CHROM.lower() else ""
    return ""
