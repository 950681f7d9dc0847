"""TEST INFRASTRUCTURE ONLY -- in-memory stand-ins for ``cyvcf2`` and ``pysam``.

The reference (jbelyeu/unfazed) does all of its phasing arithmetic in its own Python; cyvcf2 and
pysam only decode files.  Neither wheel (nor htslib) exists in this image, so the oracle drives the
*unmodified* reference modules over these two fake modules, which are views over the very same
columnar tables (``unfazed_b200.schema``) the CUDA engine consumes.

Third-party semantics honoured (cyvcf2 0.31.0 / pysam 0.22.1 as pinned in the reference's
requirements.txt:1-2; stated from the libraries' documented behaviour):

* ``Variant.gt_types`` 0/1/2/3 = HOM_REF/HET/UNKNOWN/HOM_ALT, depths int32 with -1 missing,
  ``gt_quals`` float32, ``start`` 0-based, ``POS`` 1-based, region strings 1-based inclusive and
  answered by *overlap*;
* ``AlignmentFile.fetch(contig, start, stop)`` 0-based half-open overlap in file order,
  ``ValueError`` on an unknown contig; ``mate()`` raises ``ValueError`` when there is none;
* ``AlignedSegment.get_reference_positions(full_length=True)``: one entry per query base, a
  reference position for M/=/X, ``None`` for S/I; D/N advance the reference; H/P do nothing;
  ``reference_end`` is exclusive.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this.
"""
from __future__ import annotations

import sys
import types
from array import array
from typing import Dict, List, Optional, Tuple

import numpy as np

from unfazed_b200.schema import (
    AUX_HAS_SA, AUX_SAME_REF, BASE_CHARS, QUAL_ESCAPE, ReadTable, SiteTable,
)

_VCFS: Dict[str, SiteTable] = {}
_BAMS: Dict[str, Tuple[ReadTable, int]] = {}


def register_vcf(name: str, table: SiteTable) -> None:
    _VCFS[name] = table


def register_bam(name: str, table: ReadTable, kid: int) -> None:
    _BAMS[name] = (table, kid)


def clear() -> None:
    _VCFS.clear()
    _BAMS.clear()


# ----------------------------------------------------------------------------------------------
# cyvcf2
# ----------------------------------------------------------------------------------------------

class _Info:
    def __init__(self, d=None):
        self._d = d or {}

    def get(self, key, default=None):
        return self._d.get(key, default)


class FakeVariant:
    __slots__ = ("CHROM", "start", "POS", "end", "REF", "ALT", "INFO", "gt_types",
                 "gt_ref_depths", "gt_alt_depths", "gt_quals", "_row")

    def __init__(self, view: "_VcfView", g: int):
        t = view.table
        rows = view.rows[view.g_lo[g]: view.g_lo[g + 1]]
        row = int(rows[0])
        self._row = row
        self.CHROM = t.contigs[int(view.g_contig[g])]
        self.start = int(t.pos[row])
        self.POS = self.start + 1
        ref, alts = t.ref_alts(row)
        self.REF = ref
        self.ALT = list(alts)
        self.end = self.start + len(ref)
        self.INFO = _Info()
        ns = 3 * len(t.trios)
        gt = np.full(ns, 2, dtype=np.int64)
        rd = np.full(ns, -1, dtype=np.int32)
        ad = np.full(ns, -1, dtype=np.int32)
        gq = np.full(ns, -1.0, dtype=np.float32)
        for row in rows:
            b = 3 * int(t.blk_trio[view.row_blk[row]])
            gt[b:b + 3] = t.gt[:, row]
            rd[b:b + 3] = t.rd[:, row]
            ad[b:b + 3] = t.ad[:, row]
            gq[b:b + 3] = t.gq[:, row]
        self.gt_types, self.gt_ref_depths, self.gt_alt_depths, self.gt_quals = gt, rd, ad, gq


class _VcfView:
    """Joint-VCF view of a trio-major SiteTable: one record per rec_id, sorted by (contig, pos)."""

    def __init__(self, table: SiteTable):
        self.table = table
        V = table.n_rows
        blk = np.repeat(np.arange(table.n_blocks), np.diff(table.blk_off))
        self.row_blk = blk
        contig = table.blk_contig[blk].astype(np.int64)
        rid = table.record_ids().astype(np.int64)
        order = np.lexsort((blk, rid, table.pos.astype(np.int64), contig))
        self.rows = np.arange(V)[order]
        srid = rid[order]
        first = np.ones(V, dtype=bool)
        first[1:] = (srid[1:] != srid[:-1]) | (contig[order][1:] != contig[order][:-1])
        g_first = np.nonzero(first)[0]
        self.g_lo = np.concatenate([g_first, [V]]).astype(np.int64)
        lead = self.rows[g_first]
        self.g_contig = contig[lead]
        self.g_pos = table.pos[lead].astype(np.int64)
        self.g_key = (self.g_contig << 40) + self.g_pos
        reflen = np.ones(V, dtype=np.int64)
        for r, (ref, _alts) in table.extras.items():
            reflen[r] = len(ref)
        self.g_reflen = reflen[lead]
        self.max_reflen = int(reflen.max()) if V else 1
        self.n = int(g_first.shape[0])


def _view(table: SiteTable) -> _VcfView:
    v = getattr(table, "_fake_view", None)
    if v is None:
        v = _VcfView(table)
        object.__setattr__(table, "_fake_view", v)
    return v


class VCF:
    def __init__(self, name, *a, **k):
        if name not in _VCFS:
            raise IOError("fake cyvcf2: no such file %r" % (name,))
        self._view = _view(_VCFS[name])
        t = self._view.table
        self.samples = [s for trio in t.trios for s in trio]
        self._cursor = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self._cursor >= self._view.n:
            raise StopIteration
        v = FakeVariant(self._view, self._cursor)
        self._cursor += 1
        return v

    def __call__(self, region):
        v = self._view
        t = v.table
        contig, _, span = region.rpartition(":")
        b, _, e = span.partition("-")
        B, E = int(b), int(e)
        if contig not in t.contigs:
            return
        ci = t.contigs.index(contig)
        lo0, hi0 = B - 1, E                      # 0-based half-open
        k_lo = (ci << 40) + max(lo0 - v.max_reflen + 1, 0)
        k_hi = (ci << 40) + max(hi0, 0)
        g0 = int(np.searchsorted(v.g_key, k_lo, side="left"))
        g1 = int(np.searchsorted(v.g_key, k_hi, side="left"))
        for g in range(g0, g1):
            if v.g_pos[g] + v.g_reflen[g] > lo0:
                yield FakeVariant(v, g)

    # writer-side API used only by write_vcf_output; not part of the hot path
    def add_to_header(self, *a, **k):
        pass

    def add_format_to_header(self, *a, **k):
        pass


class Writer:
    def __init__(self, *a, **k):
        pass

    def write_record(self, v):
        pass

    def close(self):
        pass


# ----------------------------------------------------------------------------------------------
# pysam
# ----------------------------------------------------------------------------------------------

class FakeRead:
    __slots__ = ("_t", "idx", "_h", "_refpos", "_seq", "_qual", "_cig", "_end")

    def __init__(self, table: ReadTable, idx: int):
        self._t = table
        self.idx = idx
        self._h = table.hdr[idx]
        self._refpos = None
        self._seq = None
        self._qual = None
        self._cig = None
        self._end = None

    def __bool__(self):
        return True

    # flags -----------------------------------------------------------------------------
    @property
    def flag(self):
        return int(self._h["flag"])

    is_qcfail = property(lambda s: bool(s.flag & 0x200))
    is_unmapped = property(lambda s: bool(s.flag & 0x4))
    is_duplicate = property(lambda s: bool(s.flag & 0x400))
    is_secondary = property(lambda s: bool(s.flag & 0x100))
    is_supplementary = property(lambda s: bool(s.flag & 0x800))
    mate_is_unmapped = property(lambda s: bool(s.flag & 0x8))
    is_paired = property(lambda s: bool(s.flag & 0x1))

    @property
    def mapping_quality(self):
        return int(self._h["mapq"])

    @property
    def reference_id(self):
        return self._contig_id()

    @property
    def next_reference_id(self):
        return self._contig_id() if (int(self._h["aux"]) & AUX_SAME_REF) else -1

    def _contig_id(self):
        t = self._t
        b = int(np.searchsorted(t.blk_off, self.idx, side="right")) - 1
        return int(t.blk_contig[b])

    @property
    def tlen(self):
        return int(self._h["tlen"])

    template_length = tlen

    @property
    def query_name(self):
        return self._t.name_of(self.idx)

    @property
    def reference_start(self):
        return int(self._h["start"])

    @property
    def cigartuples(self):
        if self._cig is None:
            o, n = int(self._h["cigar_off"]), int(self._h["n_cigar"])
            w = self._t.cigar[o:o + n]
            self._cig = [(int(x) & 15, int(x) >> 4) for x in w]
        return self._cig

    @property
    def reference_end(self):
        if self._end is None:
            e = self.reference_start
            for op, ln in self.cigartuples:
                if op in (0, 2, 3, 7, 8):
                    e += ln
            self._end = e
        return self._end

    def get_reference_positions(self, full_length=False):
        if self._refpos is None:
            out = []
            p = self.reference_start
            for op, ln in self.cigartuples:
                if op in (0, 7, 8):
                    out.extend(range(p, p + ln))
                    p += ln
                elif op in (1, 4):
                    out.extend([None] * ln)
                elif op in (2, 3):
                    p += ln
            self._refpos = out
        if full_length:
            return list(self._refpos)
        return [x for x in self._refpos if x is not None]

    def _decode(self):
        t = self._t
        q0 = t.qoff(self.idx)
        L = int(self._h["l_seq"])
        q = t.qual[q0:q0 + L]
        g = np.arange(q0, q0 + L)
        code = (t.seq2[g >> 2] >> ((g & 3) << 1).astype(np.uint8)) & 3
        chars = np.frombuffer(BASE_CHARS.encode(), dtype=np.uint8)[code]
        esc = (q & QUAL_ESCAPE) != 0
        chars = np.where(esc, np.where(code == 0, ord("N"), ord("?")), chars).astype(np.uint8)
        self._seq = chars.tobytes().decode("ascii")
        self._qual = array("B", (q & 0x7F).astype(np.uint8).tobytes())

    @property
    def query_sequence(self):
        if self._seq is None:
            self._decode()
        return self._seq

    @property
    def query_qualities(self):
        if self._qual is None:
            self._decode()
        return self._qual

    def has_tag(self, tag):
        return tag == "SA" and bool(int(self._h["aux"]) & AUX_HAS_SA)


class AlignmentFile:
    def __init__(self, name, mode="rb", reference_filename=None, **k):
        if name not in _BAMS:
            raise IOError("fake pysam: no such file %r" % (name,))
        self._t, self._kid = _BAMS[name]
        t = self._t
        self._blocks = [b for b in range(t.n_blocks) if int(t.blk_kid[b]) == self._kid]
        cache = getattr(t, "_fake_reads", None)
        if cache is None:
            cache = {}
            object.__setattr__(t, "_fake_reads", cache)
        self._cache: Dict[int, FakeRead] = cache

    def _read(self, idx: int) -> FakeRead:
        r = self._cache.get(idx)
        if r is None:
            r = FakeRead(self._t, idx)
            self._cache[idx] = r
        return r

    def __iter__(self):
        t = self._t
        for b in self._blocks:
            for i in range(int(t.blk_off[b]), int(t.blk_off[b + 1])):
                yield self._read(i)

    def _block_ends(self, b: int) -> np.ndarray:
        t = self._t
        return t.ref_ends()[int(t.blk_off[b]): int(t.blk_off[b + 1])]

    def fetch(self, contig=None, start=None, stop=None, **k):
        t = self._t
        if contig not in t.contigs:
            raise ValueError("invalid contig `%s`" % contig)
        start, stop = int(start), int(stop)
        if start < 0 or stop < start:
            raise ValueError("invalid coordinates")
        b = t.block_of(self._kid, contig)
        if b < 0:
            return iter(())
        lo, hi = int(t.blk_off[b]), int(t.blk_off[b + 1])
        starts = t.hdr["start"][lo:hi]
        ends = self._block_ends(b)
        j1 = int(np.searchsorted(starts, stop, side="left"))
        sel = np.nonzero(ends[:j1] > start)[0]
        return (self._read(lo + int(i)) for i in sel)

    def mate(self, read: FakeRead) -> FakeRead:
        if not read.is_paired:
            raise ValueError("read %s: is unpaired" % read.query_name)
        if read.mate_is_unmapped:
            raise ValueError("mate %s: is unmapped" % read.query_name)
        m = int(read._h["mate"])
        if m < 0:
            raise ValueError("mate not found")
        return self._read(m)

    def close(self):
        pass


def install() -> None:
    """Put the fakes into ``sys.modules`` as ``cyvcf2`` and ``pysam``."""
    cy = types.ModuleType("cyvcf2")
    cy.VCF = VCF
    cy.Writer = Writer
    cy.__fake__ = True
    ps = types.ModuleType("pysam")
    ps.AlignmentFile = AlignmentFile
    ps.__fake__ = True
    sys.modules["cyvcf2"] = cy
    sys.modules["pysam"] = ps
