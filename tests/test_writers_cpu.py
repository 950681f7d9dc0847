"""CPU, build container only: the output writers. `write_vcf_output` of the drop-in against the reference's own
(`unfazed/unfazed.py:336-440`), both over one recording cyvcf2 stand-in: same header additions, same
phased genotypes, same UOPS / UET arrays for every variant and sample."""
import copy
import sys
import types

import numpy as np
import pytest

from oracle import port, ref_driver
from tests.util import port_params
from unfazed_b200.synth import SynthConfig, make_dataset

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_driver.available(), reason="reference checkout not mounted")]


class _Variant:
    def __init__(self, chrom, start, end, svtype, n_samples, het_index):
        self.CHROM, self.start, self.end = chrom, start, end
        self.INFO = {"SVTYPE": svtype} if svtype else {}
        self.genotypes = [[0, 1, False] if i == het_index else [0, 0, False] for i in range(n_samples)]
        self.gt_types = np.array([1 if i == het_index else 0 for i in range(n_samples)])
        self.formats = {}

    def set_format(self, name, arr):
        self.formats[name] = np.asarray(arr).tolist()


class _Log:
    def __init__(self):
        self.header, self.formats, self.records = [], [], []


def _cyvcf2(variants, samples, log):
    class VCF:
        def __init__(self, name):
            self.samples = list(samples)

        def add_to_header(self, line):
            log.header.append(line)

        def add_format_to_header(self, d):
            log.formats.append(dict(d))

        def __iter__(self):
            return iter(copy.deepcopy(variants))

    class Writer:
        def __init__(self, outfile, vcf):
            pass

        def write_record(self, v):
            log.records.append((v.CHROM, v.start, v.end, copy.deepcopy(v.genotypes), dict(v.formats)))

    return VCF, Writer


@pytest.mark.parametrize("include_ambiguous", [False, True])
def test_vcf_writer_matches_reference(include_ambiguous, monkeypatch):
    ds = make_dataset(SynthConfig(dnms_per_trio=14, seed=401, coverage=20.0, n_trios=2, sv_frac=0.3, sv_max_len=20000,
                                  sex_chrom_frac=0.2, male_frac=1.0))
    records = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port_params()).phase(copy.deepcopy(ds.dnms))
    assert len(records) >= 5
    samples = []
    for kid, ped in ds.pedigrees.items():
        samples += [kid, ped["dad"], ped["mom"]]
    variants = []
    for d in ds.dnms:
        vt = d["vartype"] if d["vartype"] in ("DEL", "DUP", "INV", "CNV", "DUP:TANDEM", "DEL:ME", "CPX", "CTX") else None
        variants.append(_Variant(d["chrom"], int(d["start"]), int(d["end"]), vt, len(samples), samples.index(d["kid"])))

    ref_mod = ref_driver.modules()["unfazed"]
    ref_version = ref_mod.__version__
    log_ref, log_new = _Log(), _Log()
    VCF, Writer = _cyvcf2(variants, samples, log_ref)
    monkeypatch.setattr(ref_mod, "VCF", VCF)
    monkeypatch.setattr(ref_mod, "Writer", Writer)
    ref_mod.write_vcf_output("dnms.vcf", copy.deepcopy(records), include_ambiguous, True, "out.vcf", 10)

    from unfazed_b200 import unfazed as new_mod
    VCF2, Writer2 = _cyvcf2(variants, samples, log_new)
    fake = types.ModuleType("cyvcf2")
    fake.VCF, fake.Writer, fake.__fake__ = VCF2, Writer2, True
    monkeypatch.setitem(sys.modules, "cyvcf2", fake)
    new_mod.write_vcf_output("dnms.vcf", copy.deepcopy(records), include_ambiguous, True, "out.vcf", 10)

    assert log_new.formats == log_ref.formats
    assert [h.replace(new_mod.__version__, "V") for h in log_new.header] == [h.replace(ref_version, "V") for h in log_ref.header]
    assert len(log_ref.records) == len(variants)
    assert log_new.records == log_ref.records
    phased = sum(1 for r in log_ref.records for g in r[3] if g[2] is True)
    assert phased >= 3


@pytest.mark.parametrize("include_ambiguous,verbose", [(False, False), (True, True), (False, True)])
def test_bed_writer_matches_reference(include_ambiguous, verbose, tmp_path):
    """`write_bed_output` (`unfazed.py:443-515`) on the same records: same header, same rows (the read-name
    columns are a joined Python set in the reference, Q23: compared as sets)."""
    from unfazed_b200 import unfazed as new_mod
    ds = make_dataset(SynthConfig(dnms_per_trio=14, seed=402, coverage=20.0, n_trios=2, sv_frac=0.3, sv_max_len=20000,
                                  sex_chrom_frac=0.2, male_frac=1.0))
    records = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port_params()).phase(copy.deepcopy(ds.dnms))
    ref_mod = ref_driver.modules()["unfazed"]
    a, b = str(tmp_path / "ref.bed"), str(tmp_path / "new.bed")
    ref_mod.write_bed_output(copy.deepcopy(records), include_ambiguous, verbose, a, 10)
    new_mod.write_bed_output(copy.deepcopy(records), include_ambiguous, verbose, b, 10)

    def rows(path):
        out = []
        for line in open(path).read().strip().split("\n"):
            f = line.split("\t")
            if not line.startswith("#") and len(f) == 13:
                f[10] = ",".join(sorted(f[10].split(",")))
                f[12] = ",".join(sorted(f[12].split(",")))
            out.append(f)
        return out

    ra, rb = rows(a), rows(b)
    assert ra[0] == rb[0] and ra[0][0].startswith("#")
    assert ra == rb and len(ra) > 3
