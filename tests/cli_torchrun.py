"""Launched by test_gpu_multi.py under torchrun: the CLI on N GPUs over a golden case.  Every rank
registers the case's tables, runs ``python -m unfazed_b200``'s main with the reference's flags and
rank 0 writes the BED text to ``--out`` (strict, then ambiguous)."""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    from oracle.make_golden import cli_files
    from tests.golden_util import load_case
    from unfazed_b200 import datasource
    from unfazed_b200.__main__ import main
    ds, _ = load_case(a.case)
    tmp = tempfile.mkdtemp(prefix="unfz_rank%s_" % os.environ.get("RANK", "0"))
    bed, ped, pairs = cli_files(ds, tmp)
    vcf_name = "mem://golden.vcf"
    datasource.register_tables(vcf_name, sites=ds.sites)
    for kid, path in pairs:
        datasource.register_tables(path, reads=ds.reads)
    p = dict(threads=1, build="38", multiread_proc_min=1000, no_extended=False)
    p.update(ds.params)
    for amb in (False, True):
        out = "%s.%s" % (a.out, "ambiguous" if amb else "strict")
        argv = ["-d", bed, "-s", vcf_name, "-p", ped, "--bam-pairs"] + ["%s:%s" % (k, b) for k, b in pairs] + \
               ["-t", str(p["threads"]), "-g", p["build"], "--multiread-proc-min", str(p["multiread_proc_min"]),
                "--quiet", "--verbose", "-o", "bed", "--outfile", out]
        if p["no_extended"]:
            argv.append("--no-extended")
        if amb:
            argv.append("--include-ambiguous")
        main(argv)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    run()
