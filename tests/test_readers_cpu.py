"""CPU, build container only: the input readers of the orchestrator (`read_vars_bed`, `read_vars_vcf`,
`get_bam_names`, `parse_ped`, `unfazed/unfazed.py:21-160`) against the reference's own functions."""
import types

import numpy as np
import pytest

from oracle import ref_driver

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_driver.available(), reason="reference checkout not mounted")]


@pytest.fixture()
def mods():
    from unfazed_b200 import unfazed as new
    return ref_driver.modules()["unfazed"], new


def test_bed_reader(mods, tmp_path):
    ref, new = mods
    bed = tmp_path / "d.bed"
    bed.write_text("#chrom\tstart\tend\tkid\tvartype\n1\t100\t101\tk1\tPOINT\nchr2\t5\t900\tk2\tDEL\n"
                   "X\t7\t8\tk1\tweird\n3\t10\t40\tk3\tDUP\n4\t1\t2\tk1\tINDEL\n")
    assert list(new.read_vars_bed(str(bed))) == list(ref.read_vars_bed(str(bed)))
    bad = tmp_path / "bad.bed"
    bad.write_text("1\t100\t101\tk1\n")
    for m in (ref, new):
        with pytest.raises(SystemExit) as e:
            list(m.read_vars_bed(str(bad)))
        assert "must contain the following columns exactly" in str(e.value)


def test_vcf_reader(mods, monkeypatch):
    ref, new = mods
    import sys

    class V:
        def __init__(self, chrom, start, end, svtype, gts):
            self.CHROM, self.start, self.end = chrom, start, end
            self.INFO = {"SVTYPE": svtype} if svtype else {}
            self.gt_types = np.array(gts)

    variants = [V("1", 10, 11, None, [1, 0, 0]), V("2", 50, 900, "DEL", [0, 3, 1]), V("X", 5, 6, None, [2, 2, 0]),
                V("3", 1, 2, "INV", [3, 3, 3])]

    class VCF:
        samples = ["a", "b", "c"]

        def __init__(self, name):
            pass

        def __iter__(self):
            return iter(variants)

    monkeypatch.setattr(ref, "VCF", VCF)
    fake = types.ModuleType("cyvcf2")
    fake.VCF, fake.__fake__ = VCF, True
    monkeypatch.setitem(sys.modules, "cyvcf2", fake)
    got, want = list(new.read_vars_vcf("x.vcf")), list(ref.read_vars_vcf("x.vcf"))
    assert got == want and len(want) == 6


def test_bam_names_and_ped(mods, tmp_path, capsys):
    ref, new = mods
    d = tmp_path / "bams"
    d.mkdir()
    for n in ("k1.bam", "k2.bam", "k3.cram", "note.txt"):
        (d / n).write_text("")
    extra = tmp_path / "other.bam"
    extra.write_text("")
    fa = tmp_path / "ref.fa"
    fa.write_text(">1\nA\n")
    for args in ((str(d), None, str(fa)), (str(d), [["k1", str(extra)]], str(fa)), (None, [["k9", str(extra)]], None)):
        assert new.get_bam_names(*args) == ref.get_bam_names(*args)
    for m in (ref, new):
        with pytest.raises(SystemExit) as e:
            m.get_bam_names(str(d), None, None)
        assert str(e.value) == "Missing reference file for CRAM"
        with pytest.raises(SystemExit) as e:
            m.get_bam_names(str(d), None, str(tmp_path / "nope.fa"))
        assert str(e.value) == "Reference file is not valid"
        with pytest.raises(SystemExit) as e:
            m.get_bam_names(None, [["k1", str(tmp_path / "missing.bam")]], None)
        assert str(e.value).startswith("invalid filename")
    ped = tmp_path / "f.ped"
    ped.write_text("f\tk1\td1\tm1\t1\t2\nf\tk2\t0\tm2\t2\t2\nf\td1\t0\t0\t1\t1\nf\tk4\td4\tm4\t2\t2\n")
    for quiet in (True, False):
        ref.QUIET_MODE = new.QUIET_MODE = quiet
        capsys.readouterr()
        want = ref.parse_ped(str(ped), {"k1", "k2", "k3"})
        err_ref = capsys.readouterr().err
        got = new.parse_ped(str(ped), {"k1", "k2", "k3"})
        err_new = capsys.readouterr().err
        assert got == want == {"k1": {"kid": "k1", "dad": "d1", "mom": "m1", "sex": "1"}}
        assert sorted(err_new.splitlines()) == sorted(err_ref.splitlines())
    ref.QUIET_MODE = new.QUIET_MODE = False
