"""CPU: the decode path real files take -- datasource -> packers over the cyvcf2 / pysam API (here the
in-memory fakes the oracle uses) -- must hand the phasing the same information as the original tables:
the oracle port gives identical records on the packed tables, and the .npz table files round-trip."""
import copy
import sys

import numpy as np
import pytest

from oracle import fakes, port
from tests.util import norm_record, port_params
from unfazed_b200 import datasource
from unfazed_b200.synth import SynthConfig, make_dataset
from unfazed_b200.tableio import load_tables, save_tables

CASES = [
    (SynthConfig(dnms_per_trio=8, seed=301, coverage=20.0), {}),
    (SynthConfig(dnms_per_trio=6, seed=302, coverage=20.0, n_trios=2, sv_frac=0.4, sv_max_len=20000, indel_frac=0.2), {}),
    (SynthConfig(dnms_per_trio=6, seed=303, coverage=20.0, chr_prefix="chr", sex_chrom_frac=0.3, male_frac=1.0), {}),
]


@pytest.fixture()
def fake_decoders():
    fakes.install()
    datasource._registry.clear()
    yield
    datasource._registry.clear()
    for m in ("cyvcf2", "pysam"):
        if getattr(sys.modules.get(m), "__fake__", False):
            del sys.modules[m]


def _phase(sites, reads, ds, params):
    ph = port.Phaser(sites, reads, ds.pedigrees, port_params(**params))
    return ph.phase(copy.deepcopy(ds.dnms))


@pytest.mark.parametrize("cfg,params", CASES, ids=[str(c[0].seed) for c in CASES])
def test_packed_tables_phase_like_the_originals(fake_decoders, cfg, params):
    ds = make_dataset(cfg)
    vcf_name = "mem://packers.vcf"
    fakes.register_vcf(vcf_name, ds.sites)
    dnms = copy.deepcopy(ds.dnms)
    for d in dnms:
        path = "mem://%s.bam" % d["kid"]
        d["bam"], d["cram_ref"] = path, None
        fakes.register_bam(path, ds.reads, ds.reads.kids.index(d["kid"]))
    sd = port_params(**params).search_dist if hasattr(port_params(**params), "search_dist") else 5000
    sites = datasource.load_sites(vcf_name, dnms, ds.pedigrees, sd)
    reads = datasource.load_reads(dnms, sd, 151, 1000000)
    assert sites.n_rows > 0 and reads.n_reads > 0
    assert reads.n_reads <= ds.reads.n_reads              # only around the DNMs (site rows: one per record AND trio)
    want = _phase(ds.sites, ds.reads, ds, params)
    got = _phase(sites, reads, ds, params)
    assert set(got) == set(want) and len(want) > 0
    for k in want:
        assert norm_record(got[k]) == norm_record(want[k]), k


def test_table_files_round_trip(tmp_path):
    ds = make_dataset(SynthConfig(dnms_per_trio=5, seed=304, coverage=16.0))
    path = str(tmp_path / "trio.npz")
    save_tables(path, ds.sites, ds.reads, meta={"note": "round trip"})
    out = load_tables(path)
    sites, reads = out[0], out[1]
    for f in ("pos", "ref", "alt", "flag", "gt", "gq", "rd", "ad", "blk_off"):
        assert np.array_equal(getattr(sites, f), getattr(ds.sites, f)), f
    assert np.array_equal(reads.hdr, ds.reads.hdr) and np.array_equal(reads.qual, ds.reads.qual)
    assert np.array_equal(reads.cigar, ds.reads.cigar) and np.array_equal(reads.seq2, ds.reads.seq2)
    assert list(reads.kids) == list(ds.reads.kids) and list(sites.contigs) == list(ds.sites.contigs)
