import json
import os
import types

import numpy as np

from unfazed_b200.schema import SiteTable
from unfazed_b200.tableio import load_tables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["snv_noisy", "indel_cluster_many", "sv_cnv", "sexchrom_chr", "no_extended"]


def load_case(name):
    sites, reads, meta = load_tables(os.path.join(GOLDEN, name + ".npz"))
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        want = json.load(f)
    ds = types.SimpleNamespace(sites=sites, reads=reads, dnms=meta["dnms"], pedigrees=meta["pedigrees"],
                               params=meta["params"], truth=meta["truth"])
    return ds, want


def unit_vectors():
    with open(os.path.join(GOLDEN, "unit_vectors.json")) as f:
        return json.load(f)


def one_trio_table(rows):
    """SiteTable with one block; rows = list of dicts(pos, gt[3], gq[3], rd[3], ad[3])."""
    n = len(rows)
    t = SiteTable(
        trios=[("kid", "dad", "mom")], contigs=["1"], blk_trio=np.zeros(1, np.int32), blk_contig=np.zeros(1, np.int32),
        blk_off=np.array([0, n], np.int64), pos=np.array([r["pos"] for r in rows], np.int32),
        flag=np.ones(n, np.uint8), ref=np.full(n, ord("A"), np.uint8), alt=np.full(n, ord("C"), np.uint8),
        gt=np.array([r["gt"] for r in rows], np.uint8).T.copy(), gq=np.array([r["gq"] for r in rows], np.float32).T.copy(),
        rd=np.array([r["rd"] for r in rows], np.int32).T.copy(), ad=np.array([r["ad"] for r in rows], np.int32).T.copy())
    t.validate()
    return t
