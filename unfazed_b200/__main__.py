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


def setup_args():
    p = argparse.ArgumentParser(prog="unfazed", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    a = p.add_argument
    a("-v", "--version", action="version", version="%(prog)s " + str(__version__), help="Installed version ({})".format(__version__))
    a("-d", "--dnms", required=True, help="valid VCF OR BED file of the DNMs of interest> If BED, must contain chrom, start, end, kid_id, var_type columns")
    a("-s", "--sites", required=True, help="sorted/bgzipped/indexed VCF/BCF file of SNVs to identify informative sites. Must contain each kid and both parents")
    a("-p", "--ped", required=True, type=str, help="ped file including the kid and both parent IDs")
    a("-b", "--bam-dir", type=str, required=False, help="directory where bam/cram files (named {sample_id}.bam or {sample_id}.cram) are stored for offspring. If not included, --bam-pairs must be set")
    a("--bam-pairs", type=pair, nargs="*", required=False, help="space-delimited list of pairs in the format {sample_id}:{bam_path} where {sample_id} matches an offspring id from the dnm file. Can be used with --bam-dir arg, must be used in its absence")
    a("-t", "--threads", type=int, default=2, help="number of threads to use")
    a("-o", "--output-type", type=str, choices=["vcf", "bed"], help="choose output type. If --dnms is not a VCF/BCF, output must be to BED format. Defaults to match --dnms input file")
    a("--include-ambiguous", action="store_true", default=False, help="include ambiguous phasing results")
    a("--verbose", action="store_true", default=False, help="print verbose output including sites and reads used for phasing. Only applies to BED output")
    a("--outfile", default="/dev/stdout", help="name for output file. Defaults to stdout")
    a("-r", "--reference", required=False, help="reference fasta file (required for crams)")
    a("-g", "--build", choices=["37", "38", "na"], required=True, type=str, help="human genome build, used to determine sex chromosome pseudoautosomal regions. If `na` option is chosen, sex chromosomes will not be auto-phased. HG19/GRCh37 interchangeable")
    a("--no-extended", action="store_true", default=False, help="do not perform extended read-based phasing (default True)")
    a("--multiread-proc-min", type=int, default=1000, help="min number of variants required to perform multiple parallel reads of the sites file")
    a("-q", "--quiet", action="store_true", help="no logging of variant processing data")
    a("--min-gt-qual", type=int, default=20, help="min genotype and base quality for informative sites")
    a("--min-depth", type=int, default=10, help="min coverage for informative sites")
    a("--ab-homref", type=float_pair, default="0.0:0.2", help="allele balance range for homozygous reference informative sites")
    a("--ab-homalt", type=float_pair, default="0.8:1.0", help="allele balance range for homozygous alternate informative sites")
    a("--ab-het", type=float_pair, default="0.2:0.8", help="allele balance range for heterozygous informative sites")
    a("--evidence-min-ratio", type=int, default="10", help="minimum ratio of evidence for a parent to provide an unambiguous call. Default 10:1")
    a("--search-dist", type=int, default=5000, help="maximum search distance from variant for informative sites (in bases)")
    a("--insert-size-max-sample", type=int, default=1000000, help="maximum number of read inserts to sample in order to estimate concordant read insert size")
    a("--min-map-qual", type=int, default=1, help="minimum map quality for reads")
    a("--stdevs", type=int, default=3, help="number of standard deviations from the mean insert length to define a discordant read")
    a("--readlen", type=int, default=151, help="expected length of input reads")
    a("--split-error-margin", type=int, default=5, help="margin of error for the location of split read clipping in bases")
    a("--max-reads", type=int, default=100, help="maximum number of reads to collect for phasing a single variant (accepted and ignored, exactly like the reference: Q1)")
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
