"""Flattened alignment-record buffer (the `svim_aln_soa` of include/svimgpu.h).

This is the host-side container the COLLECT kernels consume.  It replaces the
stream of `pysam.AlignedSegment` objects the reference iterates in
SVIM_COLLECT.py:133 (`bam.fetch(until_eof=True)`): every field the hot path
reads from a record (SVIM_COLLECT.py:143-154, SVIM_intra.py:37-47,
SVIM_inter.py:30-46,85,91) is stored once, in structure-of-arrays form, with
the variable-length parts (CIGAR, 4-bit SEQ, SA tag text) in three blobs.

Layout (all little-endian, numpy arrays):

    tid        int32[n]   reference id (BAM refID), -1 = none
    pos        int32[n]   0-based leftmost position
    flag       uint16[n]  SAM flag
    mapq       uint8[n]
    n_cigar    uint32[n]  number of CIGAR operations
    cigar_off  uint64[n]  offset into `cigar`, in uint32 words, multiple of 4
                          (every record's CIGAR starts 16-byte aligned so a
                          warp can stream it with 128-bit loads)
    l_seq      int32[n]   number of bases stored (0 when SEQ is '*')
    seq_off    uint64[n]  byte offset into `seq` (BAM 4-bit packing, high
                          nibble first)
    sa_off     uint64[n]  byte offset into `sa`
    sa_len     uint32[n]  length of the SA tag text, 0 = no SA tag
    qname_id   uint32[n]  dense id of the read name (equal names <=> equal ids)

    cigar      uint32[]   BAM encoding  len<<4 | op   (op: MIDNSHP=X -> 0..8)
    seq        uint8[]    4-bit packed bases, code table "=ACMGRSVTWYHKDBN"
    sa         uint8[]    SA tag text as in the BAM aux field (no NUL)
"""
from __future__ import annotations

import re
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

CIGAR_OPS = "MIDNSHP=XB"
_OP_CODE = {c: i for i, c in enumerate(CIGAR_OPS)}
SEQ_NT16 = "=ACMGRSVTWYHKDBN"
_NT16_CODE = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate(SEQ_NT16):
    _NT16_CODE[ord(_c)] = _i
    _NT16_CODE[ord(_c.lower())] = _i
_NT16_CHARS = np.frombuffer(SEQ_NT16.encode(), dtype=np.uint8)
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")   # pysam CIGAR_REGEX

FLAG_UNMAPPED = 0x4
FLAG_REVERSE = 0x10
FLAG_SECONDARY = 0x100
FLAG_SUPPLEMENTARY = 0x800


def parse_cigar_string(cigar: str) -> List[Tuple[int, int]]:
    """'10S5M' -> [(4,10),(0,5)]; '*' or '' -> []."""
    if not cigar or cigar == "*":
        return []
    return [(_OP_CODE[op], int(n)) for n, op in _CIGAR_RE.findall(cigar)]


def cigar_to_string(tuples: Sequence[Tuple[int, int]]) -> Optional[str]:
    if not tuples:
        return None
    return "".join("%d%s" % (n, CIGAR_OPS[op]) for op, n in tuples)


def encode_cigar(tuples: Sequence[Tuple[int, int]]) -> np.ndarray:
    if len(tuples) == 0:
        return np.zeros(0, dtype=np.uint32)
    a = np.asarray(tuples, dtype=np.int64).reshape(-1, 2)
    return ((a[:, 1] << 4) | a[:, 0]).astype(np.uint32)


def pack_seq(seq: Optional[str]) -> np.ndarray:
    """ASCII bases -> BAM 4-bit packing (high nibble = first base)."""
    if not seq or seq == "*":
        return np.zeros(0, dtype=np.uint8)
    codes = _NT16_CODE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]
    if len(codes) & 1:
        codes = np.concatenate([codes, np.zeros(1, dtype=np.uint8)])
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


def unpack_seq(packed: np.ndarray, l_seq: int) -> str:
    if l_seq <= 0:
        return ""
    b = np.asarray(packed[: (l_seq + 1) // 2], dtype=np.uint8)
    out = np.empty(len(b) * 2, dtype=np.uint8)
    out[0::2] = _NT16_CHARS[b >> 4]
    out[1::2] = _NT16_CHARS[b & 15]
    return out[:l_seq].tobytes().decode("ascii")


class AlignmentBatch:
    """Immutable SoA view of a run of alignment records (see module docstring)."""

    FIELDS = (
        ("tid", np.int32), ("pos", np.int32), ("flag", np.uint16), ("mapq", np.uint8),
        ("n_cigar", np.uint32), ("cigar_off", np.uint64), ("l_seq", np.int32),
        ("seq_off", np.uint64), ("sa_off", np.uint64), ("sa_len", np.uint32),
        ("qname_id", np.uint32),
    )
    #: bytes of the fixed-width row per alignment (40 + qname_id)
    ROW_BYTES = 4 + 4 + 2 + 1 + 1 + 4 + 8 + 4 + 8 + 8 + 4

    def __init__(self, contig_names, contig_lengths, arrays, cigar, seq, sa, qnames=None,
                 sort_order="coordinate"):
        self.contig_names = list(contig_names)
        self.contig_lengths = np.asarray(contig_lengths, dtype=np.int64)
        for name, dt in self.FIELDS:
            a = np.ascontiguousarray(arrays[name], dtype=dt)
            setattr(self, name, a)
        self.n = int(len(self.tid))
        self.cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.sa = np.ascontiguousarray(sa, dtype=np.uint8)
        self.qnames = qnames  # list[str] indexed by qname_id, or None (synthetic: "read<id>")
        self.sort_order = sort_order
        self.cigar16 = None       # optional 16-bit packed CIGAR stream + per-record offsets (pack_cigar16); what crosses PCIe when present
        self.cigar16_off = None
        self.cigar8 = None        # optional 8-bit packed CIGAR stream (pack_cigar8); taken in preference to cigar16
        self.cigar8_off = None
        self._tid_of = {n: i for i, n in enumerate(self.contig_names)}

    # -- header-like helpers (what the reference asks of `bam`) -------------
    def get_tid(self, name: str) -> int:
        return self._tid_of.get(name, -1)

    def getrname(self, tid: int) -> str:
        if not 0 <= tid < len(self.contig_names):
            raise ValueError("reference_id %d out of range" % tid)
        return self.contig_names[tid]

    def qname(self, qid: int) -> str:
        if self.qnames is None:
            return "read%d" % qid
        return self.qnames[qid]

    # -- per-record decoding (host tests / oracle only; never on the product path)
    def cigartuples(self, i: int) -> List[Tuple[int, int]]:
        o = int(self.cigar_off[i]); n = int(self.n_cigar[i])
        c = self.cigar[o:o + n]
        return list(zip((c & 15).tolist(), (c >> 4).tolist()))

    def sequence(self, i: int) -> Optional[str]:
        l = int(self.l_seq[i])
        if l == 0:
            return None
        o = int(self.seq_off[i])
        return unpack_seq(self.seq[o:o + (l + 1) // 2], l)

    def sa_tag(self, i: int) -> Optional[str]:
        l = int(self.sa_len[i])
        if l == 0:
            return None
        o = int(self.sa_off[i])
        return self.sa[o:o + l].tobytes().decode("ascii")

    def algorithmic_bytes(self) -> int:
        """Input part of SURVEY.md §8(d) B_aln summed over the batch."""
        return int(self.n * self.ROW_BYTES + 4 * int(self.n_cigar.sum(dtype=np.int64))
                   + int(self.sa_len.sum(dtype=np.int64)))

    def pack_cigar16(self, threads: int = 0) -> "AlignmentBatch":
        """Build the 16-bit packed CIGAR stream of include/svimgpu.h (svim_aln_soa.cigar16) for this batch's records; the C ABI then
        uploads it instead of the uint32 words (half the PCIe bytes) and expands it on the device.  Host-side re-encoding only."""
        from . import io as sio
        self.cigar16, self.cigar16_off = sio.pack_cigar16(self, threads)
        return self

    def pack_cigar8(self, threads: int = 0) -> "AlignmentBatch":
        """Build the 8-bit packed CIGAR stream (svim_aln_soa.cigar8): one byte per operation under 16 bases, extension bytes for
        the rest; about a quarter of the uint32 words' PCIe bytes on noisy long reads.  Host-side re-encoding only."""
        from . import io as sio
        self.cigar8, self.cigar8_off = sio.pack_cigar8(self, threads)
        return self

    def take(self, order, sort_order: str = "unknown") -> "AlignmentBatch":
        """Records re-ordered by `order` (index array); blobs are shared, only the row arrays are permuted."""
        order = np.asarray(order, dtype=np.int64)
        arrays = {name: getattr(self, name)[order] for name, _ in self.FIELDS}
        return AlignmentBatch(self.contig_names, self.contig_lengths, arrays, self.cigar, self.seq, self.sa, self.qnames, sort_order)

    def slice(self, lo: int, hi: int) -> "AlignmentBatch":
        """Records [lo, hi) sharing the blobs (offsets stay absolute)."""
        arrays = {name: getattr(self, name)[lo:hi] for name, _ in self.FIELDS}
        return AlignmentBatch(self.contig_names, self.contig_lengths, arrays,
                              self.cigar, self.seq, self.sa, self.qnames, self.sort_order)


def _compact_slice(batch: "AlignmentBatch", lo: int, hi: int) -> "AlignmentBatch":
    """Records [lo, hi) with their OWN blobs: the CIGAR / SEQ / SA sub-ranges are cut out and the offsets rebased, so a rank that
    uploads the result moves (and holds) only its shard.  Requires the blobs to be laid out in record order (true for every reader
    and generator in this package)."""
    arrays = {name: getattr(batch, name)[lo:hi].copy() for name, _ in batch.FIELDS}
    if hi <= lo:
        return AlignmentBatch(batch.contig_names, batch.contig_lengths, arrays, np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(0, np.uint8),
                              batch.qnames, batch.sort_order)
    blobs = {}
    for blob, off, length in (("cigar", "cigar_off", (batch.n_cigar[lo:hi].astype(np.int64) + 3) & ~3),
                              ("seq", "seq_off", (batch.l_seq[lo:hi].astype(np.int64) + 1) // 2),
                              ("sa", "sa_off", batch.sa_len[lo:hi].astype(np.int64))):
        o = getattr(batch, off)[lo:hi].astype(np.int64)
        b0 = int(o.min()); b1 = int((o + length).max())
        if (np.diff(o) < 0).any():
            raise ValueError("compact slice needs the %s blob in record order" % blob)
        blobs[blob] = getattr(batch, blob)[b0:b1].copy()
        arrays[off] = (o - b0).astype(np.uint64)
    return AlignmentBatch(batch.contig_names, batch.contig_lengths, arrays, blobs["cigar"], blobs["seq"], blobs["sa"], batch.qnames, batch.sort_order)


AlignmentBatch.compact_slice = _compact_slice


class BatchBuilder:
    """Accumulates records one at a time (SAM/BAM readers, tests)."""

    def __init__(self, contig_names: Iterable[str], contig_lengths: Iterable[int],
                 sort_order: str = "coordinate"):
        self.contig_names = list(contig_names)
        self.contig_lengths = list(contig_lengths)
        self.sort_order = sort_order
        self._rows = {name: [] for name, _ in AlignmentBatch.FIELDS}
        self._cigar: List[np.ndarray] = []
        self._seq: List[np.ndarray] = []
        self._sa: List[bytes] = []
        self._cigar_words = 0
        self._seq_bytes = 0
        self._sa_bytes = 0
        self._qid = {}
        self._qnames: List[str] = []

    def add(self, qname: str, flag: int, tid: int, pos: int, mapq: int, cigar, seq: Optional[str],
            sa: Optional[str] = None, packed_seq: Optional[np.ndarray] = None, l_seq: Optional[int] = None):
        if isinstance(cigar, str):
            cigar = parse_cigar_string(cigar)
        enc = cigar if isinstance(cigar, np.ndarray) else encode_cigar(cigar)
        qid = self._qid.get(qname)
        if qid is None:
            qid = len(self._qnames)
            self._qid[qname] = qid
            self._qnames.append(qname)
        if packed_seq is None:
            packed_seq = pack_seq(seq)
            l_seq = 0 if (not seq or seq == "*") else len(seq)
        sa_b = sa.encode("ascii") if sa else b""
        r = self._rows
        r["tid"].append(tid); r["pos"].append(pos); r["flag"].append(flag); r["mapq"].append(mapq)
        r["n_cigar"].append(len(enc)); r["cigar_off"].append(self._cigar_words)
        r["l_seq"].append(l_seq); r["seq_off"].append(self._seq_bytes)
        r["sa_off"].append(self._sa_bytes); r["sa_len"].append(len(sa_b)); r["qname_id"].append(qid)
        pad = (-len(enc)) % 4
        if pad:
            enc = np.concatenate([enc, np.zeros(pad, dtype=np.uint32)])
        self._cigar.append(enc); self._cigar_words += len(enc)
        self._seq.append(packed_seq); self._seq_bytes += len(packed_seq)
        self._sa.append(sa_b); self._sa_bytes += len(sa_b)

    def finish(self) -> AlignmentBatch:
        arrays = {name: np.asarray(self._rows[name], dtype=dt) for name, dt in AlignmentBatch.FIELDS}
        cigar = np.concatenate(self._cigar) if self._cigar else np.zeros(0, np.uint32)
        seq = np.concatenate(self._seq) if self._seq else np.zeros(0, np.uint8)
        sa = np.frombuffer(b"".join(self._sa), dtype=np.uint8) if self._sa else np.zeros(0, np.uint8)
        return AlignmentBatch(self.contig_names, self.contig_lengths, arrays, cigar, seq, sa,
                              self._qnames, self.sort_order)
