"""Throughput of the host packers over the in-memory cyvcf2 / pysam stand-ins of the
oracle: reads/s of packers.pack_reads (region fetch + hashed-name mate join + columnar packing) and
site rows/s of packers.pack_sites.  The stand-ins decode lazily and cache per read, so the first pass
pays their decoding and the second pass is (almost) the packer alone.  CPU only."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fakes  # noqa: E402
from unfazed_b200 import datasource  # noqa: E402
from unfazed_b200.synth import SynthConfig, make_dataset  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ds = make_dataset(SynthConfig(dnms_per_trio=n, seed=11))
fakes.install()
vcf = "mem://time.vcf"
fakes.register_vcf(vcf, ds.sites)
dnms = copy.deepcopy(ds.dnms)
for d in dnms:
    d["bam"], d["cram_ref"] = "mem://%s.bam" % d["kid"], None
    fakes.register_bam(d["bam"], ds.reads, ds.reads.kids.index(d["kid"]))
out = {"dnms": n}
for label in ("first_pass", "second_pass"):
    t0 = time.perf_counter()
    sites = datasource.load_sites(vcf, dnms, ds.pedigrees, 5000)
    t1 = time.perf_counter()
    reads = datasource.load_reads(dnms, 5000, 151, 1000000)
    t2 = time.perf_counter()
    out[label] = {"site_rows": sites.n_rows, "site_rows_per_s": sites.n_rows / (t1 - t0), "reads": reads.n_reads,
                  "reads_per_s": reads.n_reads / (t2 - t1), "sites_s": t1 - t0, "reads_s": t2 - t1}
print(json.dumps(out))
