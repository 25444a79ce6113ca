"""ORACLE — test infrastructure, NOT product code.

CPU restatement (plain Python + scipy/statistics/random, exactly the third-party
pieces the reference itself calls) of the reference's COLLECT -> CLUSTER path,
working on the flattened record buffer (`svim_b200.records.AlignmentBatch`)
instead of pysam objects.  Every function names the reference lines it follows.

Pinning: `tools/make_golden.py` runs the UNMODIFIED reference (imported from
/root/reference with the pysam/edlib shims under tools/ref_shim) and this oracle
on the same seeded inputs in the build container, asserts equality, and commits
the reference's outputs as fixtures under tests/golden/.  `tests/test_oracle_*`
re-check the oracle against those fixtures and against the known-answer vectors
of the reference's own tests (tests/test_intra.py:8-22, tests/test_inter.py:8-11,
tests/test_Signature.py, tests/test_satag.py + chimeric_read.sam).

Third-party semantics restated here (pysam/htslib; "parity unpinned" where the
reference has no test – see SURVEY.md §8c): reference_end, query_alignment_*,
infer_read_length, FASTA fetch clamping.
"""
from __future__ import annotations

import logging
import random
from statistics import mean, stdev
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import editdist

TYPE_ORDER = ("DEL", "INS", "INV", "DUP_TAN", "BND", "DUP_INT")  # call order SVIM_CLUSTER.py:19-24


class Params:
    """Hot-path options and their defaults (SVIM_input_parsing.py:279-371)."""

    def __init__(self, **kw):
        self.min_mapq = 20
        self.min_sv_size = 40
        self.max_sv_size = 100000
        self.segment_gap_tolerance = 10
        self.segment_overlap_tolerance = 5
        self.partition_max_distance = 1000
        self.position_distance_normalizer = 900
        self.edit_distance_normalizer = 1.0
        self.cluster_max_distance = 0.5
        self.all_bnds = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


# ---------------------------------------------------------------------------
# Signature value object (fields of SVSignature.py:3-233, one class for all types)
# ---------------------------------------------------------------------------
class Sig:
    __slots__ = ("type", "contig", "start", "end", "contig2", "pos", "dir1", "dir2",
                 "direction", "copies", "fully_covered", "signature", "read", "sequence")

    def __init__(self, type, contig, start, end, signature, read, contig2=None, pos=None,
                 dir1=None, dir2=None, direction=None, copies=None, fully_covered=None, sequence=None):
        self.type = type; self.contig = contig; self.start = start; self.end = end
        self.contig2 = contig2; self.pos = pos; self.dir1 = dir1; self.dir2 = dir2
        self.direction = direction; self.copies = copies; self.fully_covered = fully_covered
        self.signature = signature; self.read = read; self.sequence = sequence

    # SVSignature.py:18,120,219 – BND source is (contig1,pos1,pos1+1)
    def source(self):
        return (self.contig, self.start, self.end)

    # SVSignature.py:124-129 (DUP_INT), :177-180 (DUP_TAN), :222-223 (BND)
    def destination(self):
        if self.type == "DUP_INT":
            return (self.contig2, self.pos, self.pos + (self.end - self.start))
        if self.type == "DUP_TAN":
            return (self.contig, self.end, self.end + self.copies * (self.end - self.start))
        if self.type == "BND":
            return (self.contig2, self.pos, self.pos + 1)
        raise AttributeError("no destination for " + self.type)

    # SVSignature.py:21-23, :70-72, :132-135, :232-233
    def key(self):
        if self.type == "INS":
            return (self.type, self.contig, self.start)
        if self.type == "DUP_INT":
            return (self.type, self.contig2, self.contig, self.pos)
        if self.type == "BND":
            return (self.type, self.contig, self.start)
        return (self.type, self.contig, self.end)

    # SVSignature.py:26-33, :75-82, :137-148
    def gap_to(self, other):
        if self.type != other.type:
            return float("inf")
        if self.type == "DUP_INT":
            if self.contig2 == other.contig2 and self.contig == other.contig:
                return max(0, other.pos - self.pos)
            return float("inf")
        if self.contig != other.contig:
            return float("inf")
        if self.type == "INS":
            return max(0, other.start - self.start)
        return max(0, other.start - self.end)

    def as_tuple(self):
        return tuple(getattr(self, f) for f in self.__slots__)

    def __repr__(self):
        return "Sig" + repr(self.as_tuple())


def make_bnd(contig1, pos1, dir1, contig2, pos2, dir2, signature, read) -> Sig:
    """SignatureTranslocation.__init__ (SVSignature.py:193-214): order the two
    breakends by (contig name STRING, position); flipping swaps and inverts the
    directions.  Stored as start=pos1, end=pos1+1, pos=pos2."""
    if contig1 < contig2 or (contig1 == contig2 and pos1 < pos2):
        return Sig("BND", contig1, pos1, pos1 + 1, signature, read, contig2=contig2, pos=pos2, dir1=dir1, dir2=dir2)
    flip = {"fwd": "rev", "rev": "fwd"}
    return Sig("BND", contig2, pos2, pos2 + 1, signature, read, contig2=contig1, pos=pos1,
               dir1=flip[dir2], dir2=flip[dir1])


# ---------------------------------------------------------------------------
# COLLECT
# ---------------------------------------------------------------------------
_ADV_REF = (1, 0, 1, 0, 0, 0, 0, 1, 1)    # SVIM_intra.py:14-29: N(3) does NOT advance pos_ref
_ADV_READ = (1, 1, 0, 0, 1, 0, 0, 1, 1)


def cigar_indels(tuples: Sequence[Tuple[int, int]], min_length: int):
    """analyze_cigar_indel (SVIM_intra.py:8-30)."""
    ref = read = 0
    out = []
    for op, ln in tuples:
        if op == 1 and ln >= min_length:
            out.append((ref, read, ln, "INS"))
        elif op == 2 and ln >= min_length:
            out.append((ref, read, ln, "DEL"))
        if op < 9:
            ref += ln * _ADV_REF[op]
            read += ln * _ADV_READ[op]
    return out


class Segment:
    """What SVIM_inter.py:28-47 reads from one alignment (primary or SA-derived)."""
    __slots__ = ("tid", "ref_start", "ref_end", "reverse", "mapq", "qas", "qae", "read_len")


def _cigar_summary(tuples, l_seq):
    """pysam semantics restated (module docstring): returns
    (ref_len, query_alignment_start, query_alignment_end, infer_read_length, hard_clipped_bases)."""
    ref_len = sum(n for op, n in tuples if op in (0, 2, 3, 7, 8))
    read_len = sum(n for op, n in tuples if op in (0, 1, 4, 5, 7, 8))
    hard = sum(n for op, n in tuples if op == 5)
    qas = 0
    for op, n in tuples:
        if op == 5:
            continue
        if op != 4:
            break
        qas += n
    if l_seq == 0:
        qae = 0
        for op, n in tuples:
            if op in (0, 1, 7, 8) or (op == 4 and qae == 0):
                qae += n
    else:
        qae = l_seq
        for op, n in reversed(tuples[1:]):
            if op == 5:
                continue
            if op != 4:
                break
            qae -= n
    return ref_len, qas, qae, (read_len if read_len > 0 else None), hard


def segment_of_record(batch, i) -> Tuple[Segment, int]:
    s = Segment()
    tuples = batch.cigartuples(i)
    ref_len, s.qas, s.qae, s.read_len, hard = _cigar_summary(tuples, int(batch.l_seq[i]))
    s.tid = int(batch.tid[i]); s.ref_start = int(batch.pos[i])
    s.ref_end = s.ref_start + (ref_len if ref_len else 1)
    s.reverse = bool(int(batch.flag[i]) & 0x10); s.mapq = int(batch.mapq[i])
    return s, hard


def sa_segments(batch, i, hard_clipped: int) -> List[Segment]:
    """retrieve_other_alignments (SVIM_COLLECT.py:44-93) reduced to the fields
    analyze_read_segments consumes.  SA-derived segments inherit the main
    record's SEQ, hence its l_seq, for query_alignment_end."""
    if hard_clipped > 0:
        return []
    tag = batch.sa_tag(i)
    if tag is None:
        return []
    from svim_b200.records import parse_cigar_string
    out = []
    l_seq = int(batch.l_seq[i])
    for element in tag.split(";"):
        if element == "":
            continue
        fields = element.split(",")
        if len(fields) != 6:
            logging.warning("SA tag does not consist of 6 fields.")
            continue
        mapq = int(fields[4])
        if not 0 <= mapq <= 255:
            mapq = 0                      # OverflowError branch, SVIM_COLLECT.py:81-84
        int(fields[5])                    # nm must parse
        tuples = parse_cigar_string(fields[3])
        s = Segment()
        ref_len, s.qas, s.qae, s.read_len, _ = _cigar_summary(tuples, l_seq)
        s.tid = batch.get_tid(fields[0]); s.ref_start = int(fields[1]) - 1
        s.ref_end = s.ref_start + (ref_len if ref_len else 1)
        s.reverse = fields[2] != "+"; s.mapq = mapq
        out.append(s)
    return out


def similar(chr1, start1, end1, chr2, start2, end2, threshold=0.3):
    """is_similar (SVIM_inter.py:11-21)."""
    span1 = end1 - start1
    span2 = end2 - start2
    c1 = (start1 + end1) // 2
    c2 = (start2 + end2) // 2
    d = abs(c1 - c2) / 900 + abs(span1 - span2) / max(span1, span2)
    return chr1 == chr2 and d < threshold


def _py_slice(seq: Optional[str], a, n):
    """`seq[a:a+n]` with the reference's `except TypeError -> ""` (SVIM_inter.py:84-94)."""
    try:
        return seq[a:a + n]
    except TypeError:
        return ""


def segment_signatures(batch, primary: Segment, supplementaries: List[Segment], primary_seq: Optional[str],
                       read_name: str, p: Params):
    """analyze_read_segments (SVIM_inter.py:24-302), restated around one
    observation: in every branch a breakend is the END of `cur` as the read
    traverses it (ref_end-1,'fwd' on the forward strand; ref_start,'rev' on the
    reverse strand) joined to the BEGINNING of `nxt` (ref_start,'fwd' |
    ref_end-1,'rev')."""
    name = batch.getrname
    chain = []
    for seg in [primary] + supplementaries:
        if seg.reverse:
            if seg.read_len is None:            # SVIM_inter.py:31-34
                logging.warning("Skipping alignment: cannot infer read length")
                continue
            qs, qe = seg.read_len - seg.qae, seg.read_len - seg.qas
        else:
            qs, qe = seg.qas, seg.qae
        chain.append((qs, qe, seg))
    chain.sort(key=lambda t: (t[0], t[1]))      # stable, SVIM_inter.py:49

    sigs: List[Sig] = []
    twins: List[Sig] = []                       # --all_bnds extras
    tandems = []                                # (chr, start, end, fully_covered, forward)
    junctions = []                              # (dir1, dir2, chr1, pos1, chr2, pos2)
    tol_o, tol_g = p.segment_overlap_tolerance, p.segment_gap_tolerance
    lo, hi = p.min_sv_size, p.max_sv_size

    def leave(seg):   # breakend where the read leaves `seg`
        return (seg.ref_start, "rev") if seg.reverse else (seg.ref_end - 1, "fwd")

    def enter(seg):   # breakend where the read enters `seg`
        return (seg.ref_end - 1, "rev") if seg.reverse else (seg.ref_start, "fwd")

    for (cqs, cqe, cur), (nqs, nqe, nxt) in zip(chain, chain[1:]):
        dr = nqs - cqe
        (p1, d1), (p2, d2) = leave(cur), enter(nxt)

        def junction(c1=None, c2=None):
            c1 = c1 or name(cur.tid); c2 = c2 or name(nxt.tid)
            sigs.append(make_bnd(c1, p1, d1, c2, p2, d2, "suppl", read_name))
            junctions.append((d1, d2, c1, p1, c2, p2))

        def twin(c):
            if p.all_bnds:
                twins.append(make_bnd(c, p1, d1, c, p2, d2, "suppl", read_name))

        if cur.tid != nxt.tid:                                   # SVIM_inter.py:205-240
            c1, c2 = name(cur.tid), name(nxt.tid)
            if -tol_o <= dr <= tol_g:
                junction(c1, c2)
            continue
        chrom = name(cur.tid)
        if cur.reverse == nxt.reverse:                           # SVIM_inter.py:66-150
            rev = cur.reverse
            dref = (cur.ref_start - nxt.ref_end) if rev else (nxt.ref_start - cur.ref_end)
            if dr < -tol_o:
                continue
            if dref >= -tol_o:
                dev = dr - dref
                if dev >= lo:
                    if dref <= tol_g:                            # :80-94
                        if not rev:
                            seq = _py_slice(primary_seq, cqe, dev)
                            at = cur.ref_end
                        else:
                            try:
                                a = primary.read_len - nqs
                                seq = primary_seq[a:a + dev]
                            except TypeError:
                                seq = ""
                            at = cur.ref_start
                        sigs.append(Sig("INS", chrom, at, at + dev, "suppl", read_name, sequence=seq))
                elif -hi <= dev <= -lo:
                    if dr <= tol_g:                              # :96-106
                        at = nxt.ref_end if rev else cur.ref_end
                        assert at - dev >= at
                        sigs.append(Sig("DEL", chrom, at, at - dev, "suppl", read_name))
                        if p.all_bnds:
                            twins.append(make_bnd(chrom, at - 1, "fwd", chrom, at - dev, "fwd", "suppl", read_name))
                elif dev < -hi:
                    if dr <= tol_g:                              # :108-116
                        junction(chrom, chrom)
            elif dref <= -lo:                                    # :117-150
                if not rev:
                    td = (chrom, nxt.ref_start, cur.ref_end)
                    full = nxt.ref_end > cur.ref_start
                else:
                    td = (chrom, cur.ref_start, nxt.ref_end)
                    full = nxt.ref_start < cur.ref_end
                if full or dref >= -hi:
                    tandems.append(td + (full, not rev))
                    twin(chrom)
                else:
                    junction(chrom, chrom)
        else:                                                    # SVIM_inter.py:152-204
            if not (-tol_o <= dr <= tol_g):
                continue
            # inversion anchors: forward->reverse compares the segment ENDs
            # ("left_*"), reverse->forward the segment STARTs ("right_*")
            if not cur.reverse:
                a, b, side = cur.ref_end, nxt.ref_end, "left"
            else:
                a, b, side = cur.ref_start, nxt.ref_start, "right"
            if nxt.ref_start - cur.ref_end >= -tol_o:
                size, inv = b - a, (a, b, side + "_fwd")
            elif cur.ref_start - nxt.ref_end >= -tol_o:
                size, inv = a - b, (b, a, side + "_rev")
            else:
                continue
            if lo <= size <= hi:
                assert inv[1] >= inv[0]
                sigs.append(Sig("INV", chrom, inv[0], inv[1], "suppl", read_name, direction=inv[2]))
                twin(chrom)
            elif size > hi:
                junction(chrom, chrom)

    # tandem duplications, SVIM_inter.py:242-272 (direction of the FIRST group is
    # never refreshed – :255 vs :265-269 – kept)
    if tandems:
        grp_chr, starts, ends, covered = None, [], [], []
        direction = None
        for chrom, s, e, full, fwd in tandems:
            if grp_chr is None:
                grp_chr, starts, ends, covered, direction = chrom, [s], [e], [full], fwd
            elif similar(grp_chr, mean(starts), mean(ends), chrom, s, e) and direction == fwd:
                starts.append(s); ends.append(e); covered.append(full)
            else:
                sigs.append(_tandem(grp_chr, starts, ends, covered, read_name))
                grp_chr, starts, ends, covered = chrom, [s], [e], [full]
        sigs.append(_tandem(grp_chr, starts, ends, covered, read_name))

    # interspersed duplications, SVIM_inter.py:274-300
    for t, (td1, td2, tc1, tp1, tc2, tp2) in enumerate(junctions):
        for bd1, bd2, bc1, bp1, bc2, bp2 in junctions[:t]:
            if not (bd1 == td2 and bd2 == td1):
                continue
            if not similar(bc1, bp1, bp1 + 1, tc2, tp2, tp2 + 1, 0.1):
                continue
            if bc2 != tc1 or bd1 != bd2:
                continue
            if bd1 == "fwd":
                if lo <= tp1 - bp2 + 1 <= hi:
                    sigs.append(Sig("DUP_INT", bc2, bp2, tp1 + 1, "suppl", read_name,
                                    contig2=bc1, pos=int(mean([bp1 + 1, tp2]))))
            else:
                if lo <= bp2 - tp1 <= hi:
                    sigs.append(Sig("DUP_INT", bc2, tp1, bp2 + 1, "suppl", read_name,
                                    contig2=bc1, pos=int(mean([bp1, tp2 + 1]))))
    return sigs, twins


def _tandem(chrom, starts, ends, covered, read_name):
    s, e = int(mean(starts)), int(mean(ends))
    assert e >= s
    return Sig("DUP_TAN", chrom, s, e, "suppl", read_name, copies=len(starts), fully_covered=any(covered))


def record_indel_signatures(batch, i, read_name, p: Params):
    """analyze_alignment_indel (SVIM_intra.py:33-51)."""
    chrom = batch.getrname(int(batch.tid[i]))
    ref_start = int(batch.pos[i])
    sigs, twins = [], []
    seq = None
    for pos_ref, pos_read, ln, typ in cigar_indels(batch.cigartuples(i), p.min_sv_size):
        s = ref_start + pos_ref
        if typ == "DEL":
            sigs.append(Sig("DEL", chrom, s, s + ln, "cigar", read_name))
            if p.all_bnds:
                twins.append(make_bnd(chrom, s, "fwd", chrom, s + ln, "fwd", "cigar", read_name))
        else:
            if seq is None:
                seq = batch.sequence(i) or ""
            sigs.append(Sig("INS", chrom, s, s + ln, "cigar", read_name, sequence=seq[pos_read:pos_read + ln]))
    return sigs, twins


def collect(batch, p: Params):
    """analyze_alignment_file_coordsorted (SVIM_COLLECT.py:132-167)."""
    sigs: List[Sig] = []
    twins: List[Sig] = []
    for i in range(batch.n):
        flag = int(batch.flag[i])
        if flag & 0x4 or flag & 0x100 or int(batch.mapq[i]) < p.min_mapq:
            continue
        read_name = batch.qname(int(batch.qname_id[i]))
        a, b = record_indel_signatures(batch, i, read_name, p)
        sigs += a; twins += b
        if flag & 0x800:
            continue
        primary, hard = segment_of_record(batch, i)
        good = [s for s in sa_segments(batch, i, hard) if s.mapq >= p.min_mapq]
        a, b = segment_signatures(batch, primary, good, batch.sequence(i), read_name, p)
        sigs += a; twins += b
    return sigs, twins


def read_groups(batch):
    """bam_iterator (SVIM_COLLECT.py:8-41): runs of consecutive records with the same read name ->
    (primary indices, supplementary indices, secondary indices)."""
    groups = []
    qid = batch.qname_id
    i = 0
    while i < batch.n:
        j = i
        prim, sup, sec = [], [], []
        while j < batch.n and qid[j] == qid[i]:
            f = int(batch.flag[j])
            (sec if f & 0x100 else sup if f & 0x800 else prim).append(j)
            j += 1
        groups.append((prim, sup, sec))
        i = j
    return groups


def collect_querysorted(batch, p: Params):
    """analyze_alignment_file_querysorted (SVIM_COLLECT.py:96-129): the REAL supplementary records of a read are its
    segments (no SA parsing); CIGAR analysis runs on the primary and on every good supplementary."""
    sigs: List[Sig] = []
    twins: List[Sig] = []
    for prim, sup, _sec in read_groups(batch):
        if len(prim) != 1:
            continue
        i = prim[0]
        if int(batch.flag[i]) & 0x4 or int(batch.mapq[i]) < p.min_mapq:
            continue
        good = [j for j in sup if not (int(batch.flag[j]) & 0x4) and int(batch.mapq[j]) >= p.min_mapq]
        read_name = batch.qname(int(batch.qname_id[i]))
        for j in [i] + good:
            a, b = record_indel_signatures(batch, j, batch.qname(int(batch.qname_id[j])), p)
            sigs += a; twins += b
        primary, _ = segment_of_record(batch, i)
        a, b = segment_signatures(batch, primary, [segment_of_record(batch, j)[0] for j in good], batch.sequence(i), read_name, p)
        sigs += a; twins += b
    return sigs, twins


# ---------------------------------------------------------------------------
# CLUSTER
# ---------------------------------------------------------------------------
def form_partitions(sigs: List[Sig], max_distance):
    """form_partitions (SVIM_clustering.py:17-29)."""
    parts: List[List[Sig]] = []
    for s in sorted(sigs, key=Sig.key):
        if parts and parts[-1][-1].gap_to(s) <= max_distance:
            parts[-1].append(s)
        else:
            parts.append([s])
    return parts


def haplotype_edit_distance(a: Sig, b: Sig, genome, pad=100):
    """compute_haplotype_edit_distance (SVIM_clustering.py:32-45)."""
    ws = min(a.start, b.start) - pad
    we = max(a.start, b.start) + pad

    def hap(s):
        return (genome.fetch(s.contig, max(0, ws), max(0, s.start)).upper() + s.sequence.upper()
                + genome.fetch(s.contig, max(0, s.start), max(0, we)).upper())
    return editdist.edit_distance(hap(a), hap(b))


def span_position_distance(a: Sig, b: Sig, genome, p: Params):
    """span_position_distance (SVIM_clustering.py:47-96); operation order kept."""
    N = p.position_distance_normalizer
    t = a.type
    if t == "BND":
        d1 = abs(a.start - b.start)
        d2 = abs(a.pos - b.pos)
        if a.dir1 == b.dir1 and a.dir2 == b.dir2:
            return (d1 + d2) / 3000
        return 99999
    span1, span2 = a.end - a.start, b.end - b.start
    if t == "INS":
        pd = abs(a.start - b.start) / N
        if pd > 2 * p.cluster_max_distance:
            return pd + abs(span1 - span2) / max(span1, span2)
        ed = haplotype_edit_distance(a, b, genome)
        return pd + ed / max(span1, span2) / p.edit_distance_normalizer
    c1, c2 = (a.start + a.end) // 2, (b.start + b.end) // 2
    pd = abs(c1 - c2) / N
    if t == "DUP_INT":
        pdd = abs(a.pos - b.pos) / N
        return pd + pdd + abs(span1 - span2) / max(span1, span2)
    return pd + abs(span1 - span2) / max(span1, span2)


def clusters_from_partitions(parts, genome, p: Params, stats=None):
    """clusters_from_partitions (SVIM_clustering.py:122-180)."""
    from scipy.cluster.hierarchy import linkage, fcluster
    out = []
    large = dup = 0
    random.seed(1524)
    for part in parts:
        if len(part) > 100:
            sample = random.sample(part, 100); large += 1
        else:
            sample = part
        t = sample[0].type
        if t == "INV":
            kept = sample
        else:
            dups = set()
            for i in range(len(sample) - 1):
                for j in range(i + 1, len(sample)):
                    if sample[i].read == sample[j].read and \
                            span_position_distance(sample[i], sample[j], genome, p) <= p.cluster_max_distance:
                        dups.add(j)
            dup += len(dups)
            kept = [s for k, s in enumerate(sample) if k not in dups]
        if len(kept) == 1:
            out.append([kept[0]])
            continue
        dist = []
        for i in range(len(kept) - 1):
            for j in range(i + 1, len(kept)):
                if t != "INV" and kept[i].read == kept[j].read:
                    dist.append(99999)
                else:
                    dist.append(span_position_distance(kept[i], kept[j], genome, p))
        Z = linkage(np.array(dist), method="average")
        ids = list(fcluster(Z, p.cluster_max_distance, criterion="distance"))
        groups = [[] for _ in range(max(ids))]
        for s, c in zip(kept, ids):
            groups[c - 1].append(s)
        out.extend(groups)
    if stats is not None:
        stats["large_partitions"] = large
        stats["duplicate_signatures"] = dup
    return out


class Cluster:
    """SignatureClusterUniLocal / BiLocal value object (SVSignature.py:236-310)."""
    __slots__ = ("type", "contig", "start", "end", "dest_contig", "dest_start", "dest_end", "score", "size",
                 "members", "std_span", "std_pos", "dir1", "dir2")

    def __init__(self, **kw):
        for f in self.__slots__:
            setattr(self, f, kw.get(f))


def score_of(members, std_span, std_pos, span, t):
    """calculate_score (SVIM_clustering.py:183-211)."""
    if std_span is None or std_pos is None:
        ss = ps = 0
    else:
        ss = 1 - min(1, std_span / span)
        ps = 1 - min(1, std_pos / span)
    if t == "INV":
        left = sum(1 for m in members if m.direction in ("left_fwd", "left_rev"))
        right = sum(1 for m in members if m.direction in ("right_fwd", "right_rev"))
        n = min(80, min(left, right) + sum(1 for m in members if m.direction == "all"))
    else:
        n = min(80, len(members))
    return n + ss * (n / 8) + ps * (n / 8)


def consolidate(groups, bilocal: bool):
    """consolidate_clusters_unilocal / _bilocal (SVIM_clustering.py:214-303)."""
    out = []
    for g in groups:
        n = len(g)
        t = g[0].type
        src = [m.source() for m in g]
        a_s = sum(s[1] for s in src) / n
        a_e = sum(s[2] for s in src) / n
        if n > 1:
            sd_span = stdev([s[2] - s[1] for s in src])
            sd_pos = stdev([(s[2] + s[1]) / 2 for s in src])
        else:
            sd_span = sd_pos = None
        rs, re_ = int(round(a_s)), int(round(a_e))
        if not bilocal:
            out.append(Cluster(type=t, contig=src[0][0], start=rs, end=re_, size=n, members=g,
                               score=score_of(g, sd_span, sd_pos, a_e - a_s, t), std_span=sd_span, std_pos=sd_pos))
            continue
        if t == "DUP_TAN":
            mc = max(m.copies for m in g)
            out.append(Cluster(type=t, contig=src[0][0], start=rs, end=re_, dest_contig=src[0][0],
                               dest_start=re_, dest_end=re_ + mc * (re_ - rs), size=n, members=g,
                               score=score_of(g, sd_span, sd_pos, a_e - a_s, t), std_span=sd_span, std_pos=sd_pos))
            continue
        dst = [m.destination() for m in g]
        d_s = sum(d[1] for d in dst) / n
        d_e = sum(d[2] for d in dst) / n
        c = Cluster(type=t, contig=src[0][0], start=rs, end=re_, dest_contig=dst[0][0],
                    dest_start=int(round(d_s)), dest_end=int(round(d_e)), size=n, members=g)
        if t == "DUP_INT":
            span = mean([a_e - a_s, d_e - d_s])
            if n > 1:
                dd_span = stdev([d[2] - d[1] for d in dst])
                dd_pos = stdev([(d[2] + d[1]) / 2 for d in dst])
                c.std_span = mean([sd_span, dd_span]); c.std_pos = mean([sd_pos, dd_pos])
            c.score = score_of(g, c.std_span, c.std_pos, span, t)
        else:  # BND
            d1 = {m.dir1 for m in g}; d2 = {m.dir2 for m in g}
            assert len(d1) == 1 and len(d2) == 1
            c.dir1, c.dir2 = d1.pop(), d2.pop()
            if n > 1:
                c.std_span = sd_pos
                c.std_pos = stdev([(d[2] + d[1]) / 2 for d in dst])
            c.score = score_of(g, c.std_span, c.std_pos, 500, t)
        out.append(c)
    return out


def partition_and_cluster(sigs, genome, p: Params, stats=None):
    """partition_and_cluster (SVIM_clustering.py:375-385) for one type."""
    if not sigs:
        return []
    parts = form_partitions(sigs, p.partition_max_distance)
    groups = clusters_from_partitions(parts, genome, p, stats)
    if stats is not None:
        stats["partitions"] = len(parts); stats["clusters"] = len(groups)
    t = sigs[0].type
    if t in ("DEL", "INS", "INV"):
        return sorted(consolidate(groups, False), key=lambda c: (c.contig, (c.end + c.start) / 2))
    return consolidate(groups, True)


def cluster(sigs, genome, p: Params, stats=None):
    """cluster_sv_signatures (SVIM_CLUSTER.py:7-26).  Returns the 6-tuple in the
    reference's RETURN order (DEL, INS, INV, DUP_TAN, DUP_INT, BND)."""
    res = {}
    for t in TYPE_ORDER:
        st = {} if stats is not None else None
        res[t] = partition_and_cluster([s for s in sigs if s.type == t], genome, p, st)
        if stats is not None:
            stats[t] = st
    return (res["DEL"], res["INS"], res["INV"], res["DUP_TAN"], res["DUP_INT"], res["BND"])


class Cand:
    """DUP_INT candidate value object (SVCandidate.py:424-453 fields used by the clustering twin)."""
    __slots__ = ("type", "contig", "start", "end", "dest_contig", "dest_start", "dest_end", "members", "score", "std_span", "std_pos", "cutpaste")

    def __init__(self, contig, start, end, dest_contig, dest_start, dest_end, members, score, std_span, std_pos, cutpaste=False):
        self.type = "DUP_INT"
        self.contig, self.start, self.end = contig, max(0, start), end
        self.dest_contig, self.dest_start, self.dest_end = dest_contig, max(0, dest_start), dest_end
        self.members, self.score, self.std_span, self.std_pos, self.cutpaste = members, score, std_span, std_pos, cutpaste

    def key(self):                       # Candidate.get_key, SVCandidate.py:24-26
        return (self.type, self.contig, self.end)

    def gap_to(self, other):             # Candidate.downstream_distance_to, SVCandidate.py:29-36
        if self.type == other.type and self.contig == other.contig:
            return max(0, other.start - self.end)
        return float("inf")

    def as_tuple(self):
        return tuple(getattr(self, f) for f in self.__slots__)


def candidate_distance(a: Cand, b: Cand, p: Params):
    """span_position_distance_intdup_candidates (SVIM_clustering.py:110-119)."""
    N = p.position_distance_normalizer
    span1, span2 = a.end - a.start, b.end - b.start
    c1, c2 = (a.start + a.end) // 2, (b.start + b.end) // 2
    return abs(c1 - c2) / N + abs(a.dest_start - b.dest_start) / N + abs(span1 - span2) / max(span1, span2)


def partition_and_cluster_candidates(cands, p: Params):
    """partition_and_cluster_candidates (SVIM_clustering.py:306-372)."""
    from scipy.cluster.hierarchy import linkage, fcluster
    parts = []
    for c in sorted(cands, key=Cand.key):
        if parts and parts[-1][-1].gap_to(c) <= p.partition_max_distance:
            parts[-1].append(c)
        else:
            parts.append([c])
    groups = []
    random.seed(1524)
    for part in parts:
        if len(part) == 1:
            groups.append([part[0]]); continue
        sample = random.sample(part, 100) if len(part) > 100 else part
        dist = [candidate_distance(sample[i], sample[j], p) for i in range(len(sample) - 1) for j in range(i + 1, len(sample))]
        ids = list(fcluster(linkage(np.array(dist), method="average"), p.cluster_max_distance, criterion="distance"))
        new = [[] for _ in range(max(ids))]
        for c, k in zip(sample, ids):
            new[k - 1].append(c)
        groups.extend(new)
    out = []
    for g in groups:
        n = len(g)
        spans = [c.std_span for c in g if c.std_span is not None]
        poss = [c.std_pos for c in g if c.std_pos is not None]
        if g[0].type == "DUP_INT":
            out.append(Cand(g[0].contig, int(round(sum(c.start for c in g) / n)), int(round(sum(c.end for c in g) / n)),
                            g[0].dest_contig, int(round(sum(c.dest_start for c in g) / n)), int(round(sum(c.dest_end for c in g) / n)),
                            [m for c in g for m in c.members], max(c.score for c in g),
                            mean(spans) if spans else None, mean(poss) if poss else None, any(c.cutpaste for c in g)))
    return out


# ---------------------------------------------------------------------------
# scipy restated (spec for the CUDA linkage kernel; verified against scipy in tests)
# ---------------------------------------------------------------------------
def linkage_average_restated(dist: Sequence[float], n: int):
    """scipy.cluster.hierarchy.linkage(y, 'average') = _hierarchy.nn_chain + sort +
    label (scipy is a third-party dependency of the reference, setup.py:39, call
    site SVIM_clustering.py:170).  Returns Z as a list of [a, b, d, size]."""
    D = [float(x) for x in dist]

    def ix(i, j):
        if i > j:
            i, j = j, i
        return n * i - i * (i + 1) // 2 + (j - i - 1)
    size = [1] * n
    Z = []
    chain = []
    for _ in range(n - 1):
        if not chain:
            chain.append(next(i for i in range(n) if size[i] > 0))
        while True:
            x = chain[-1]
            if len(chain) > 1:
                y = chain[-2]; cur = D[ix(x, y)]
            else:
                y = -1; cur = float("inf")
            for i in range(n):
                if size[i] == 0 or i == x:
                    continue
                d = D[ix(x, i)]
                if d < cur:
                    cur = d; y = i
            if len(chain) > 1 and y == chain[-2]:
                break
            chain.append(y)
        chain.pop(); chain.pop()
        if x > y:
            x, y = y, x
        nx, ny = size[x], size[y]
        Z.append([x, y, cur, nx + ny])
        size[x] = 0; size[y] = nx + ny
        for i in range(n):
            if size[i] == 0 or i == y:
                continue
            D[ix(i, y)] = (nx * D[ix(i, x)] + ny * D[ix(i, y)]) / (nx + ny)
    order = sorted(range(n - 1), key=lambda k: Z[k][2])      # stable (mergesort in scipy)
    Z = [Z[k] for k in order]
    parent = list(range(2 * n - 1))

    def find(a):
        r = a
        while parent[r] != r:
            r = parent[r]
        while parent[a] != r:
            parent[a], a = r, parent[a]
        return r
    nxt = n
    for row in Z:
        a, b = find(int(row[0])), find(int(row[1]))
        row[0], row[1] = (a, b) if a < b else (b, a)
        parent[a] = parent[b] = nxt
        nxt += 1
    return Z


def fcluster_distance_restated(Z, t: float, n: int):
    """scipy fcluster(Z, t, 'distance') = get_max_dist_for_each_cluster +
    cluster_monocrit (_hierarchy.pyx); ids numbered in DFS order, left first."""
    MD = [0.0] * (n - 1)
    for k in range(n - 1):   # children always precede parents after the sort
        m = Z[k][2]
        for c in (int(Z[k][0]), int(Z[k][1])):
            if c >= n and MD[c - n] > m:
                m = MD[c - n]
        MD[k] = m
    T = [0] * n
    visited = [False] * (2 * n - 1)
    stack = [2 * n - 2]
    leader = -1
    ncl = 0
    while stack:
        r = stack[-1]
        k = r - n
        left, right = int(Z[k][0]), int(Z[k][1])
        if leader == -1 and MD[k] <= t:
            leader = r; ncl += 1
        if left >= n and not visited[left]:
            visited[left] = True; stack.append(left); continue
        if right >= n and not visited[right]:
            visited[right] = True; stack.append(right); continue
        if left < n:
            if leader == -1:
                ncl += 1
            T[left] = ncl
        if right < n:
            if leader == -1:
                ncl += 1
            T[right] = ncl
        if leader == r:
            leader = -1
        stack.pop()
    return T


# ---------------------------------------------------------------------------
# GENOTYPE (SVIM_genotyping.py:34-93) — SURVEY.md §8f rank 2
# ---------------------------------------------------------------------------
class GenoParams:
    """Genotyping options and defaults (SVIM_input_parsing.py:404-437)."""

    def __init__(self, **kw):
        self.min_mapq = 20
        self.minimum_score = 3
        self.minimum_depth = 4
        self.homozygous_threshold = 0.8
        self.heterozygous_threshold = 0.2
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class GenoCand:
    """What genotype() reads from a candidate (SVIM_genotyping.py:39-51): the locus — get_source() for DEL/INV,
    get_destination() for INS/DUP_INT (SVCandidate.py:20-21,216-217,451-452) —, the score and the member read names;
    and the four attributes it writes (:75-93)."""
    __slots__ = ("contig", "start", "end", "score", "reads", "support_fraction", "genotype", "ref_reads", "alt_reads")

    def __init__(self, contig, start, end, score, reads):
        self.contig, self.start, self.end, self.score, self.reads = contig, start, end, score, list(reads)
        self.support_fraction, self.genotype, self.ref_reads, self.alt_reads = ".", "./.", None, None   # constructor defaults

    def result(self):
        return [self.support_fraction, self.genotype, self.ref_reads, self.alt_reads]


def record_reference_ends(batch) -> np.ndarray:
    """htslib bam_endpos per record: pos + reference length (M,D,N,=,X) for mapped records with a CIGAR, a zero length
    counting as 1; pos + 1 otherwise.  pysam's `reference_end` is this value, or None for unmapped / CIGAR-less records."""
    pos = np.asarray(batch.pos, dtype=np.int64)
    consumes = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1] + [0] * 7, dtype=np.int64)
    off = batch.cigar_off.astype(np.int64)
    cnt = batch.n_cigar.astype(np.int64)
    lo = int(off.min()) if batch.n else 0                   # a slice shares the blob of its parent: only its own range is summed
    hi = int((off + cnt).max()) if batch.n else 0
    words = batch.cigar[lo:hi]
    cs = np.zeros(words.size + 1, dtype=np.int64)
    np.cumsum((words >> 4).astype(np.int64) * consumes[words & 15], out=cs[1:])
    rlen = cs[off - lo + cnt] - cs[off - lo]
    rlen[((batch.flag & 0x4) != 0) | (batch.n_cigar == 0)] = 0
    return pos + np.maximum(rlen, 1)


def fetch_region(batch, ends, tid, start, stop):
    """pysam AlignmentFile.fetch(contig, start, stop) on a coordinate-sorted, indexed file (htslib bam_itr): the records of
    the contig with pos < stop and bam_endpos > start, in file order."""
    if start > stop:
        raise ValueError("invalid coordinates: start (%i) > stop (%i)" % (start, stop))
    idx = np.nonzero((batch.tid == tid) & (batch.pos < stop) & (ends > start))[0]
    return idx.tolist()


def genotype(cands: List[GenoCand], batch, type: str, gp: GenoParams, ends=None):
    """genotype (SVIM_genotyping.py:34-93).  Read identity is the read NAME (sets of query_name, :51,:53,:73,:77)."""
    if ends is None:
        ends = record_reference_ends(batch)
    for cand in cands:
        if cand.score < gp.minimum_score:                                     # :39-40
            continue
        contig, start, end = cand.contig, cand.start, cand.end
        if type in ("INS", "DUP_INT"):
            end = start                                                       # :45
        tid = batch.get_tid(contig)
        if tid < 0:
            raise KeyError(contig)
        contig_length = int(batch.contig_lengths[tid])
        variant = set(cand.reads)                                             # :51
        reference = set()
        aln_no = 0
        for i in fetch_region(batch, ends, tid, max(0, start - 1000), min(contig_length, end + 1000)):   # :49
            if aln_no >= 500:                                                 # :57
                break
            name = batch.qname(int(batch.qname_id[i]))
            if name in variant:                                               # :63-64
                continue
            flag = int(batch.flag[i])
            if flag & 0x4 or flag & 0x100 or int(batch.mapq[i]) < gp.min_mapq:   # :65-66
                continue
            aln_no += 1
            rs = int(batch.pos[i])
            re_ = int(ends[i]) if int(batch.n_cigar[i]) > 0 else None         # pysam: None without a CIGAR -> TypeError if compared
            if type in ("DEL", "INV"):                                        # :69-73
                min_overlap = min((end - start) / 2, 2000)
                if (rs < (end - min_overlap) and re_ > (end + 100)) or (rs < (start - 100) and re_ > (start + min_overlap)):
                    reference.add(name)
            if type in ("INS", "DUP_INT"):                                    # :74-76
                if rs < (start - 100) and re_ > (end + 100):
                    reference.add(name)
        n_var, n_ref = len(variant), len(reference)
        if n_var + n_ref >= gp.minimum_depth:                                 # :78-88
            cand.support_fraction = n_var / (n_var + n_ref)
            if cand.support_fraction >= gp.homozygous_threshold:
                cand.genotype = "1/1"
            elif gp.heterozygous_threshold <= cand.support_fraction < gp.homozygous_threshold:
                cand.genotype = "0/1"
            elif cand.support_fraction < gp.heterozygous_threshold:
                cand.genotype = "0/0"
            else:
                cand.genotype = "./."
        elif n_var + n_ref > 0:                                               # :89-91
            cand.support_fraction = n_var / (n_var + n_ref)
            cand.genotype = "./."
        else:                                                                 # :92-94
            cand.support_fraction = "."
            cand.genotype = "./."
        cand.ref_reads = n_ref
        cand.alt_reads = n_var
    return cands


# ---------------------------------------------------------------------------
# cut&paste search (SVIM_merging.py:12-29) — SURVEY.md §8f rank 3, second half
# ---------------------------------------------------------------------------
def cluster_source_distance(a_start, a_end, b_start, b_end, N):
    """span_position_distance_clusters (SVIM_clustering.py:99-107) on two source intervals.  Contigs are NOT compared."""
    span1, span2 = a_end - a_start, b_end - b_start
    c1, c2 = (a_start + a_end) // 2, (b_start + b_end) // 2
    return abs(c1 - c2) / N + abs(span1 - span2) / max(span1, span2)


def flag_cutpaste(ins_sources, del_sources, N=900, max_distance=1.0):
    """flag_cutpaste_candidates (SVIM_merging.py:12-29) reduced to what it decides: for every DUP_INT cluster source
    (start, end) the index of and the distance to the closest deletion cluster — first of the minima, because
    `sorted(..., key=distance)[0]` is stable — and the cut&paste flag `closest <= del_ins_dup_max_distance`.
    No deletion cluster -> IndexError, zero spans on both sides -> ZeroDivisionError, like the reference."""
    out = []
    for (s, e) in ins_sources:
        distances = [(j, cluster_source_distance(ds, de, s, e, N)) for j, (ds, de) in enumerate(del_sources)]
        j, d = sorted(distances, key=lambda o: o[1])[0]
        out.append((j, d, d <= max_distance))
    return out
