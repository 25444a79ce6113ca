"""Minimal pysam stand-in so the UNMODIFIED reference modules import and run in
this container (golden-vector generation and oracle validation only; never used
by the product or on the GPU box).

Restates the pysam/htslib semantics the hot path relies on:
  cigartuples op codes MIDNSHP=X -> 0..8
  reference_end        = reference_start + sum(M,D,N,=,X)   (htslib bam_endpos:
                         a zero reference length counts as 1)
  query_alignment_start= leading soft clips (hard clips skipped)
  query_alignment_end  = l_seq - trailing soft clips; without SEQ the sum of
                         M,I,=,X plus the leading soft clip
  infer_read_length()  = sum(M,I,S,H,=,X), None if 0
  get_cigar_stats()[0] = per-op base counts
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from svim_b200.records import parse_cigar_string, cigar_to_string  # noqa: E402
from svim_b200 import io as _io  # noqa: E402


class AlignedSegment:
    def __init__(self):
        self.query_name = None
        self._seq = None
        self.flag = 0
        self.reference_id = -1
        self.reference_start = -1
        self._mapq = 0
        self._cigar = []
        self.next_reference_id = -1
        self.next_reference_start = -1
        self.template_length = 0
        self.query_qualities = None
        self._tags = {}

    # --- simple attributes -------------------------------------------------
    @property
    def query_sequence(self):
        return self._seq

    @query_sequence.setter
    def query_sequence(self, s):
        self._seq = s if s else None

    @property
    def mapping_quality(self):
        return self._mapq

    @mapping_quality.setter
    def mapping_quality(self, v):
        if not 0 <= v <= 255:
            raise OverflowError("value too large to convert to uint8_t")
        self._mapq = v

    @property
    def cigarstring(self):
        return cigar_to_string(self._cigar)

    @cigarstring.setter
    def cigarstring(self, s):
        self._cigar = parse_cigar_string(s)

    @property
    def cigartuples(self):
        return list(self._cigar) if self._cigar else None

    @cigartuples.setter
    def cigartuples(self, t):
        self._cigar = list(t) if t else []

    # --- flags ---------------------------------------------------------------
    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_secondary(self):
        return bool(self.flag & 0x100)

    @property
    def is_supplementary(self):
        return bool(self.flag & 0x800)

    # --- tags ----------------------------------------------------------------
    def get_tag(self, name):
        return self._tags[name]  # KeyError when absent, like pysam

    def set_tag(self, name, value, value_type=None):
        self._tags[name] = value

    def set_tags(self, tags):
        self._tags = {t[0]: t[1] for t in tags}

    def has_tag(self, name):
        return name in self._tags

    # --- derived coordinates ------------------------------------------------------
    def _l_seq(self):
        return len(self._seq) if self._seq else 0

    @property
    def reference_end(self):
        if not self._cigar or self.is_unmapped or self.reference_start < 0:
            return None
        rlen = sum(n for op, n in self._cigar if op in (0, 2, 3, 7, 8))
        return self.reference_start + (rlen if rlen else 1)

    @property
    def query_alignment_start(self):
        off = 0
        for op, n in self._cigar:
            if op == 5:
                continue
            if op == 4:
                off += n
            else:
                break
        return off

    @property
    def query_alignment_end(self):
        end = self._l_seq()
        if end == 0:
            for op, n in self._cigar:
                if op in (0, 1, 7, 8) or (op == 4 and end == 0):
                    end += n
            return end
        for k in range(len(self._cigar) - 1, 0, -1):
            op, n = self._cigar[k]
            if op == 5:
                continue
            if op == 4:
                end -= n
            else:
                break
        return end

    @property
    def query_alignment_sequence(self):
        if self._seq is None:
            return None
        return self._seq[self.query_alignment_start:self.query_alignment_end]

    def infer_read_length(self):
        l = sum(n for op, n in self._cigar if op in (0, 1, 4, 5, 7, 8))
        return l if l > 0 else None

    def get_cigar_stats(self):
        bases = [0] * 11
        blocks = [0] * 11
        for op, n in self._cigar:
            bases[op] += n
            blocks[op] += 1
        return bases, blocks


class AlignmentFile:
    """File-backed (SAM/BAM path) or batch-backed (`from_batch`) record source."""

    def __init__(self, path=None, mode="r", batch=None):
        self._batch = batch if batch is not None else _io.read_alignments(path)
        b = self._batch
        self.header = {"HD": {"SO": b.sort_order}, "SQ": [{"SN": n, "LN": int(l)} for n, l in zip(b.contig_names, b.contig_lengths)]}
        self.references = tuple(b.contig_names)
        self.lengths = tuple(int(l) for l in b.contig_lengths)

    @classmethod
    def from_batch(cls, batch):
        return cls(batch=batch)

    def get_tid(self, name):
        return self._batch.get_tid(name)

    def check_index(self):
        """svim:93-98 only asks whether an index exists; region fetch below does not need one."""
        return True

    def getrname(self, tid):
        return self._batch.getrname(tid)

    get_reference_name = getrname

    def segment(self, i):
        b = self._batch
        a = AlignedSegment()
        a.query_name = b.qname(int(b.qname_id[i]))
        a._seq = b.sequence(i)
        a.flag = int(b.flag[i])
        a.reference_id = int(b.tid[i])
        a.reference_start = int(b.pos[i])
        a._mapq = int(b.mapq[i])
        a._cigar = b.cigartuples(i)
        sa = b.sa_tag(i)
        if sa is not None:
            a._tags["SA"] = sa
        return a

    def fetch(self, *args, contig=None, start=None, stop=None, **kwargs):
        """No region: every record in file order.  Region (htslib bam_itr semantics on a coordinate-sorted file): records of
        `contig` with pos < stop and bam_endpos > start, in file order; bam_endpos = pos + reference length for mapped
        records with a CIGAR (a zero length counts as 1), else pos + 1."""
        b = self._batch
        if contig is None:
            for i in range(b.n):
                yield self.segment(i)
            return
        tid = b.get_tid(contig)
        if tid < 0:
            raise ValueError("invalid contig `%s`" % contig)
        if start is None:
            start = 0
        if stop is None:
            stop = int(b.contig_lengths[tid])
        if start > stop:
            raise ValueError("invalid coordinates: start (%i) > stop (%i)" % (start, stop))
        for i in range(b.n):
            if int(b.tid[i]) != tid or int(b.pos[i]) >= stop:
                continue
            pos = int(b.pos[i])
            end = pos + 1
            if not (int(b.flag[i]) & 0x4) and int(b.n_cigar[i]) > 0:
                rlen = sum(n for op, n in b.cigartuples(i) if op in (0, 2, 3, 7, 8))
                end = pos + (rlen if rlen else 1)
            if end > start:
                yield self.segment(i)

    def get_reference_length(self, contig):
        tid = self._batch.get_tid(contig)
        if tid < 0:
            raise KeyError(contig)
        return int(self._batch.contig_lengths[tid])

    def close(self):
        pass


class FastaFile:
    def __init__(self, path):
        self._g = None
        self._path = path

    def _genome(self):
        if self._g is None:
            self._g = _genome_cache(self._path)
        return self._g

    def fetch(self, contig, start, end):
        return self._genome().fetch(contig, start, end)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_GENOMES = {}


def register_genome(path, genome):
    """Let in-memory genomes be 'opened' by path (avoids re-parsing per type)."""
    _GENOMES[path] = genome


def _genome_cache(path):
    g = _GENOMES.get(path)
    if g is None:
        g = _io.Genome.from_fasta(path)
        _GENOMES[path] = g
    return g
