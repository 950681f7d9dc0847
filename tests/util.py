"""Shared helpers of the parity tests: run the oracle port and the CUDA engine on one dataset."""
import copy

import numpy as np

from oracle import port
from unfazed_b200.synth import SynthConfig, make_dataset


def norm_record(rec):
    r = dict(rec)
    for k in ("dad_sites", "mom_sites", "dad_reads", "mom_reads", "cnv_dad_sites", "cnv_mom_sites"):
        if isinstance(r[k], list):
            r[k] = sorted(r[k])
    return r


def port_params(**kw):
    fields = port.Params.__dataclass_fields__
    return port.Params(**{k: v for k, v in kw.items() if k in fields})


def run_port(ds, **kw):
    p = port_params(**kw)
    ph = port.Phaser(ds.sites, ds.reads, ds.pedigrees, p)
    recs = ph.phase(copy.deepcopy(ds.dnms))
    return recs, ph


def gpu_kwargs(**kw):
    allowed = {"threads", "build", "no_extended", "multiread_proc_min", "ab_homref", "ab_homalt", "ab_het",
               "min_gt_qual", "min_depth", "search_dist", "insert_size_max_sample", "stdevs", "min_map_qual",
               "readlen", "split_error_margin", "evidence_min_ratio"}
    return {k: v for k, v in kw.items() if k in allowed}


def summarize_all(records, include_ambiguous=True, ratio=10):
    return {k: port.summarize_record(copy.deepcopy(r), include_ambiguous, True, ratio) for k, r in records.items()}
