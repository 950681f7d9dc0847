"""``python -m unfazed_b200`` -- the unfazed command line: the same 26 flags, defaults and messages
as the reference's ``unfazed/__main__.py`` (:19-239)."""
from __future__ import annotations

import argparse
import sys

from . import __version__


def pair(arg):
    return arg.split(":")


def float_pair(arg):
    return [float(x) for x in arg.split(":")]


# (flags, kwargs) -- the flag names, types, defaults and choices are the reference's (__main__.py:23-223);
# the help texts are ours
FLAGS = [
    (("-d", "--dnms"), dict(required=True, help="de novo variants to phase: VCF/BCF, or BED with the columns chrom, start, end, kid, vartype")),
    (("-s", "--sites"), dict(required=True, help="indexed VCF/BCF with the genotypes of every kid and both parents (source of informative sites)")),
    (("-p", "--ped"), dict(required=True, type=str, help="pedigree file naming each kid's father and mother")),
    (("-b", "--bam-dir"), dict(type=str, required=False, help="folder with one {sample}.bam or {sample}.cram per kid (alternative: --bam-pairs)")),
    (("--bam-pairs",), dict(type=pair, nargs="*", required=False, help="explicit {sample}:{alignment file} pairs; override --bam-dir for those samples")),
    (("-t", "--threads"), dict(type=int, default=2, help="accepted for compatibility; the GPU batch does not use host threads")),
    (("-o", "--output-type"), dict(type=str, choices=["vcf", "bed"], help="bed or vcf (vcf only when --dnms is a VCF/BCF); default: same as the input")),
    (("--include-ambiguous",), dict(action="store_true", default=False, help="also report variants whose evidence is ambiguous")),
    (("--verbose",), dict(action="store_true", default=False, help="BED only: add the supporting sites and read names")),
    (("--outfile",), dict(default="/dev/stdout", help="where to write the result")),
    (("-r", "--reference"), dict(required=False, help="reference FASTA, needed to decode CRAM")),
    (("-g", "--build"), dict(choices=["37", "38", "na"], required=True, type=str, help="genome build for the pseudoautosomal bounds used by sex-chromosome auto-phasing; na disables it")),
    (("--no-extended",), dict(action="store_true", default=False, help="skip extended read-backed phasing (only reads overlapping the variant itself are used)")),
    (("--multiread-proc-min",), dict(type=int, default=1000, help="from this many variants on, sites are searched with the reference's find_many semantics")),
    (("-q", "--quiet"), dict(action="store_true", help="suppress per-variant messages")),
    (("--min-gt-qual",), dict(type=int, default=20, help="lowest genotype quality of an informative site; also the lowest base quality")),
    (("--min-depth",), dict(type=int, default=10, help="lowest read depth of an informative site")),
    (("--ab-homref",), dict(type=float_pair, default="0.0:0.2", help="lo:hi allele balance accepted for a homozygous-reference genotype")),
    (("--ab-homalt",), dict(type=float_pair, default="0.8:1.0", help="lo:hi allele balance accepted for a homozygous-alternate genotype")),
    (("--ab-het",), dict(type=float_pair, default="0.2:0.8", help="lo:hi allele balance accepted for a heterozygous genotype")),
    (("--evidence-min-ratio",), dict(type=int, default="10", help="evidence for one parent must outweigh the other by this factor for an unambiguous call")),
    (("--search-dist",), dict(type=int, default=5000, help="how far (bp) from the variant informative sites are searched")),
    (("--insert-size-max-sample",), dict(type=int, default=1000000, help="reads sampled from the head of the alignment file for the insert-size estimate")),
    (("--min-map-qual",), dict(type=int, default=1, help="lowest mapping quality of a usable read")),
    (("--stdevs",), dict(type=int, default=3, help="standard deviations above the mean insert size that make a pair discordant")),
    (("--readlen",), dict(type=int, default=151, help="nominal read length")),
    (("--split-error-margin",), dict(type=int, default=5, help="tolerance (bp) between a split read's clip position and an SV breakpoint")),
    (("--max-reads",), dict(type=int, default=100, help="accepted and ignored, exactly like the reference (SURVEY Q1)")),
]


def setup_args():
    p = argparse.ArgumentParser(prog="unfazed", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-v", "--version", action="version", version="%(prog)s " + str(__version__))
    for names, kw in FLAGS:
        p.add_argument(*names, **kw)
    return p


def main(argv=None):
    print("\nUNFAZED v{}".format(__version__), file=sys.stderr)
    parser = setup_args()
    args = parser.parse_args(argv)
    print("Genome Build: {}\n".format(args.build), file=sys.stderr)
    if args.bam_dir is None and args.bam_pairs is None:
        print("\nMissing required argument: --bam-dir or --bam-pairs must be set\n", file=sys.stderr)
        sys.exit(parser.print_help())
    from .unfazed import unfazed
    unfazed(args)


if __name__ == "__main__":
    sys.exit(main() or 0)
