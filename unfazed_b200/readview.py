"""Read objects handed back by the drop-in ``read_collector`` functions: views over the ReadTable
with the handful of pysam.AlignedSegment attributes the reference's callers touch
(snv_phaser.py:16-70, site_searcher.py:50-78)."""
from __future__ import annotations

import numpy as np

from .schema import BASE_CHARS, QUAL_ESCAPE, ReadTable


class ReadView:
    __slots__ = ("table", "idx", "_pos")

    def __init__(self, table: ReadTable, idx: int):
        self.table, self.idx, self._pos = table, int(idx), None

    def __repr__(self):
        return "<ReadView %s %d-%d>" % (self.query_name, self.reference_start, self.reference_end)

    @property
    def query_name(self):
        return self.table.name_of(self.idx)

    @property
    def reference_start(self):
        return int(self.table.hdr["start"][self.idx])

    @property
    def reference_end(self):
        return int(self.table.ref_ends()[self.idx])

    @property
    def flag(self):
        return int(self.table.hdr["flag"][self.idx])

    @property
    def mapping_quality(self):
        return int(self.table.hdr["mapq"][self.idx])

    @property
    def tlen(self):
        return int(self.table.hdr["tlen"][self.idx])

    @property
    def cigartuples(self):
        h = self.table.hdr[self.idx]
        o, n = int(h["cigar_off"]), int(h["n_cigar"])
        return [(int(w) & 15, int(w) >> 4) for w in self.table.cigar[o:o + n]]

    def get_reference_positions(self, full_length=False):
        if self._pos is None:
            out, p = [], self.reference_start
            for op, ln in self.cigartuples:
                if op in (0, 7, 8):
                    out.extend(range(p, p + ln))
                    p += ln
                elif op in (1, 4):
                    out.extend([None] * ln)
                elif op in (2, 3):
                    p += ln
            self._pos = out
        return list(self._pos) if full_length else [x for x in self._pos if x is not None]

    def _span(self):
        t = self.table
        q0, L = t.qoff(self.idx), int(t.hdr["l_seq"][self.idx])
        return q0, L

    @property
    def query_sequence(self):
        t = self.table
        q0, L = self._span()
        g = np.arange(q0, q0 + L)
        code = (t.seq2[g >> 2] >> ((g & 3) << 1).astype(np.uint8)) & 3
        ch = np.frombuffer(BASE_CHARS.encode(), dtype=np.uint8)[code]
        esc = (t.qual[q0:q0 + L] & QUAL_ESCAPE) != 0
        return np.where(esc, np.where(code == 0, ord("N"), ord("?")), ch).astype(np.uint8).tobytes().decode("ascii")

    @property
    def query_qualities(self):
        q0, L = self._span()
        return (self.table.qual[q0:q0 + L] & 0x7F).tolist()
