"""CPU, build container only: the oracle port against the UNMODIFIED reference run live over the
in-memory cyvcf2/pysam fakes, on seeds that are not in the golden set."""
import copy

import pytest

from oracle import port, ref_driver
from tests.util import norm_record, port_params
from unfazed_b200.synth import SynthConfig, make_dataset

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_driver.available(), reason="reference checkout not mounted")]

CASES = [
    (SynthConfig(dnms_per_trio=12, seed=201, coverage=24.0), {}),
    (SynthConfig(dnms_per_trio=12, seed=202, coverage=24.0, cluster_frac=0.4, indel_frac=0.3), {"multiread_proc_min": 1}),
    (SynthConfig(dnms_per_trio=10, seed=203, coverage=24.0, sv_frac=0.6, sv_max_len=30000), {}),
    (SynthConfig(dnms_per_trio=10, seed=204, coverage=24.0, n_trios=2, sex_chrom_frac=0.3, male_frac=1.0, chr_prefix="chr"),
     {"multiread_proc_min": 1, "threads": 2}),
    (SynthConfig(dnms_per_trio=10, seed=205, coverage=24.0), {"min_gt_qual": 30, "min_depth": 20, "ab_het": [0.3, 0.7], "readlen": 150}),
    # a wider sweep of the knobs that change control flow in the reference
    (SynthConfig(dnms_per_trio=10, seed=206, coverage=20.0), {"no_extended": True}),
    (SynthConfig(dnms_per_trio=10, seed=207, coverage=20.0, sv_frac=0.7, sv_max_len=200000), {"multiread_proc_min": 1, "threads": 2}),
    (SynthConfig(dnms_per_trio=10, seed=208, coverage=20.0, sv_frac=0.5, sv_max_len=5000, indel_frac=0.3), {"no_extended": True}),
    (SynthConfig(dnms_per_trio=8, seed=209, coverage=30.0, search_dist=1500), {"search_dist": 1500}),
    (SynthConfig(dnms_per_trio=6, seed=210, coverage=16.0, search_dist=12000, site_spacing=150), {"search_dist": 12000}),
    (SynthConfig(dnms_per_trio=10, seed=211, coverage=20.0, noise=False), {"evidence_min_ratio": 3}),
    (SynthConfig(dnms_per_trio=8, seed=212, coverage=20.0, n_trios=3, cluster_frac=0.6), {"multiread_proc_min": 1}),
    (SynthConfig(dnms_per_trio=10, seed=213, coverage=20.0, sex_chrom_frac=0.5, male_frac=0.5), {"build": "37"}),
    (SynthConfig(dnms_per_trio=10, seed=214, coverage=20.0, indel_frac=0.8), {"min_map_qual": 30, "split_error_margin": 10}),
    (SynthConfig(dnms_per_trio=8, seed=215, coverage=20.0, sv_frac=0.9, sv_max_len=1000000), {}),
]


@pytest.mark.parametrize("cfg,params", CASES, ids=[str(c[0].seed) for c in CASES])
def test_port_equals_reference(cfg, params):
    ds = make_dataset(cfg)
    want = ref_driver.phase(ds, **params)
    ph = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port_params(**params))
    got = ph.phase(copy.deepcopy(ds.dnms))
    assert set(got) == set(want)
    for k in want:
        assert norm_record(got[k]) == norm_record(want[k]), k
    sw = ref_driver.summarize(copy.deepcopy(want))
    for k in want:
        assert port.summarize_record(copy.deepcopy(got[k]), True, True, 10) == \
            {**sw[k], "origin_parent_reads": port.summarize_record(copy.deepcopy(got[k]), True, True, 10)["origin_parent_reads"],
             "other_parent_reads": port.summarize_record(copy.deepcopy(got[k]), True, True, 10)["other_parent_reads"]}, k
