"""Save / load the columnar tables as compressed .npz (fixtures, caches of decoded files)."""
from __future__ import annotations

import json

import numpy as np

from .schema import READ_HDR, ReadTable, SiteTable


def save_tables(path: str, sites: SiteTable, reads: ReadTable, meta: dict = None) -> None:
    extras = {str(k): [v[0], list(v[1])] for k, v in sites.extras.items()}
    np.savez_compressed(
        path,
        s_trios=np.array(json.dumps(sites.trios)), s_contigs=np.array(json.dumps(sites.contigs)),
        s_blk_trio=sites.blk_trio, s_blk_contig=sites.blk_contig, s_blk_off=sites.blk_off,
        s_pos=sites.pos, s_flag=sites.flag, s_ref=sites.ref, s_alt=sites.alt, s_gt=sites.gt, s_gq=sites.gq,
        s_rd=sites.rd, s_ad=sites.ad, s_extras=np.array(json.dumps(extras)),
        s_rec_id=sites.rec_id if sites.rec_id is not None else np.zeros(0, dtype=np.int64),
        r_kids=np.array(json.dumps(reads.kids)), r_contigs=np.array(json.dumps(reads.contigs)),
        r_blk_kid=reads.blk_kid, r_blk_contig=reads.blk_contig, r_blk_off=reads.blk_off,
        r_hdr=reads.hdr.view(np.uint8).reshape(-1), r_cigar=reads.cigar, r_qual=reads.qual, r_seq2=reads.seq2,
        r_names=np.array(json.dumps(reads.names)) if reads.names is not None else np.array("null"),
        meta=np.array(json.dumps(meta or {})),
    )


def load_tables(path: str):
    z = np.load(path, allow_pickle=False)
    js = lambda k: json.loads(str(z[k]))
    extras = {int(k): (v[0], list(v[1])) for k, v in js("s_extras").items()}
    rec_id = z["s_rec_id"]
    sites = SiteTable(
        trios=[tuple(t) for t in js("s_trios")], contigs=js("s_contigs"), blk_trio=z["s_blk_trio"],
        blk_contig=z["s_blk_contig"], blk_off=z["s_blk_off"], pos=z["s_pos"], flag=z["s_flag"], ref=z["s_ref"],
        alt=z["s_alt"], gt=z["s_gt"], gq=z["s_gq"], rd=z["s_rd"], ad=z["s_ad"], extras=extras,
        rec_id=rec_id if rec_id.shape[0] else None)
    reads = ReadTable(
        kids=js("r_kids"), contigs=js("r_contigs"), blk_kid=z["r_blk_kid"], blk_contig=z["r_blk_contig"],
        blk_off=z["r_blk_off"], hdr=z["r_hdr"].view(READ_HDR), cigar=z["r_cigar"], qual=z["r_qual"],
        seq2=z["r_seq2"], names=js("r_names"))
    sites.validate()
    reads.validate()
    return sites, reads, js("meta")
