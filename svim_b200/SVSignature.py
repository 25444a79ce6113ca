"""Signature and signature-cluster objects handed to the downstream stages.

Attribute surface of the reference's SVSignature.py (classes :36-233 and :236-310):
COMBINE, genotyping and the BED/VCF writers only read attributes and call
get_source / get_destination / get_key / downstream_distance_to / as_string, so these
classes are duck-type compatible.  Behaviour is table-driven here: every signature
type declares where its source/destination live and which coordinate is its sort key.
"""
import logging

_INF = float("inf")


class Signature:
    """A structural-variant signature observed in one read.

    Every class declares its attributes as __slots__ (plus __dict__, so downstream code may still attach attributes like the
    reference's plain classes allow): svim_b200/csrc_host/fastobj.c builds hundreds of thousands of these by storing straight
    into the slots.  `type` is a class attribute of the signature classes (the reference sets the same constant per instance)."""
    __slots__ = ("__dict__",)
    type = None

    def __init__(self, contig, start, end, signature, read):
        self.contig, self.start, self.end = contig, start, end
        self.signature, self.read = signature, read
        if end < start:
            logging.warning("Signature with invalid coordinates (end < start): " + self.as_string())

    # -- geometry ---------------------------------------------------------------------
    def get_source(self):
        return (self.contig, self.start, self.end)

    def _partition_anchor(self):
        """(grouping tuple, coordinate the next signature is measured from, own start)."""
        c, s, e = self.get_source()
        return (self.type, c), e, s

    def get_key(self):
        c, s, e = self.get_source()
        return (self.type, c, e)

    def downstream_distance_to(self, signature2):
        """Gap (>= 0) from this signature to `signature2`; inf across types/contigs."""
        grp, frm, _ = self._partition_anchor()
        grp2, _, to = signature2._partition_anchor()
        return max(0, to - frm) if grp == grp2 else _INF

    # -- text -------------------------------------------------------------------------------
    def _label(self):
        return "{0};{1}".format(self.type, self.signature)

    def as_string(self, sep="\t"):
        c, s, e = self.get_source()
        return sep.join(str(x) for x in (c, s, e, self._label(), self.read))

    def _as_string_bilocal(self, sep):
        sc, ss, se = self.get_source()
        dc, ds, de = self.get_destination()
        return sep.join(("{0}:{1}-{2}".format(sc, ss, se), "{0}:{1}-{2}".format(dc, ds, de), self._label(), str(self.read)))


def _checked_interval(start, end):
    assert end >= start
    return start, end


class SignatureDeletion(Signature):
    """contig:start-end (0-based, end exclusive) is missing from the sample."""
    __slots__ = ("contig", "start", "end", "signature", "read")
    type = "DEL"

    def __init__(self, contig, start, end, signature, read):
        self.contig = contig
        self.start, self.end = _checked_interval(start, end)
        self.signature, self.read = signature, read


class SignatureInsertion(Signature):
    """end-start bases (`sequence`) inserted before contig:start."""
    __slots__ = ("contig", "start", "end", "signature", "read", "sequence")
    type = "INS"

    def __init__(self, contig, start, end, signature, read, sequence):
        self.contig = contig
        self.start, self.end = _checked_interval(start, end)
        self.signature, self.read, self.sequence = signature, read, sequence

    def get_key(self):
        return (self.type, self.contig, self.start)

    def _partition_anchor(self):
        return (self.type, self.contig), self.start, self.start


class SignatureInversion(Signature):
    """contig:start-end is inverted; `direction` names the breakpoint seen."""
    __slots__ = ("contig", "start", "end", "signature", "read", "direction")
    type = "INV"

    def __init__(self, contig, start, end, signature, read, direction):
        self.contig = contig
        self.start, self.end = _checked_interval(start, end)
        self.signature, self.read, self.direction = signature, read, direction

    def _label(self):
        return "{0};{1};{2}".format(self.type, self.direction, self.signature)


class SignatureInsertionFrom(Signature):
    """contig1:start-end was copied to contig2:pos (interspersed duplication)."""
    __slots__ = ("contig1", "start", "end", "contig2", "pos", "signature", "read")
    type = "DUP_INT"

    def __init__(self, contig1, start, end, contig2, pos, signature, read):
        self.contig1 = contig1
        self.start, self.end = _checked_interval(start, end)
        self.contig2, self.pos = contig2, pos
        self.signature, self.read = signature, read

    def get_source(self):
        return (self.contig1, self.start, self.end)

    def get_destination(self):
        return (self.contig2, self.pos, self.pos + (self.end - self.start))

    def get_key(self):
        return (self.type, self.contig2, self.contig1, self.pos)

    def _partition_anchor(self):
        return (self.type, self.contig2, self.contig1), self.pos, self.pos

    def as_string(self, sep="\t"):
        return self._as_string_bilocal(sep)


class SignatureDuplicationTandem(Signature):
    """contig:start-end repeated `copies` more times right after `end`."""
    __slots__ = ("contig", "start", "end", "copies", "fully_covered", "signature", "read")
    type = "DUP_TAN"

    def __init__(self, contig, start, end, copies, fully_covered, signature, read):
        self.contig = contig
        self.start, self.end = _checked_interval(start, end)
        self.copies, self.fully_covered = copies, fully_covered
        self.signature, self.read = signature, read

    def get_destination(self):
        return (self.contig, self.end, self.end + self.copies * (self.end - self.start))

    def _label(self):
        return "{0};{1};{2}".format(self.type, self.signature, self.copies)

    def as_string(self, sep="\t"):
        return self._as_string_bilocal(sep)


_FLIP = {"fwd": "rev", "rev": "fwd"}


class SignatureTranslocation(Signature):
    """Novel adjacency contig1:pos1 -- contig2:pos2; the breakend with the smaller
    (contig name, position) comes first, which flips the directions when swapped."""
    __slots__ = ("contig1", "pos1", "direction1", "contig2", "pos2", "direction2", "signature", "read")
    type = "BND"

    def __init__(self, contig1, pos1, direction1, contig2, pos2, direction2, signature, read):
        if not (contig1 < contig2 or (contig1 == contig2 and pos1 < pos2)):
            contig1, pos1, direction1, contig2, pos2, direction2 = \
                contig2, pos2, _FLIP[direction2], contig1, pos1, _FLIP[direction1]
        self.contig1, self.pos1, self.direction1 = contig1, pos1, direction1
        self.contig2, self.pos2, self.direction2 = contig2, pos2, direction2
        self.signature, self.read = signature, read

    def get_source(self):
        return (self.contig1, self.pos1, self.pos1 + 1)

    def get_destination(self):
        return (self.contig2, self.pos2, self.pos2 + 1)

    def get_key(self):
        return (self.type, self.contig1, self.pos1)

    def as_string(self, sep="\t"):
        return self._as_string_bilocal(sep)


class SignatureClusterUniLocal(Signature):
    """Cluster of DEL / INS / INV signatures (one locus)."""
    __slots__ = ("contig", "start", "end", "score", "size", "members", "type", "std_span", "std_pos")

    def __init__(self, contig, start, end, score, size, members, type, std_span, std_pos):
        self.contig, self.start, self.end = contig, start, end
        self.score, self.size, self.members, self.type = score, size, members, type
        self.std_span, self.std_pos = std_span, std_pos

    def get_length(self):
        return self.end - self.start

    def _members_text(self):
        return "[" + "][".join(m.as_string("|") for m in self.members) + "]"

    def get_bed_entry(self):
        name = ";".join(str(x) for x in (self.type, self.size, self.std_span, self.std_pos))
        return "\t".join(str(x) for x in (self.contig, self.start, self.end, name, self.score, self._members_text()))

    def get_vcf_entry(self):
        if self.type not in ("DEL", "INS", "INV"):
            return None
        info = "SVTYPE={0};END={1};SVLEN={2};STD_SPAN={3};STD_POS={4}".format(self.type, self.end, self.end - self.start,
                                                                              self.std_span, self.std_pos)
        return "\t".join(str(x) for x in (self.contig, self.start + 1, ".", "N", "<" + self.type + ">", ".", "PASS", info))


class SignatureClusterBiLocal(Signature):
    """Cluster of DUP_TAN / DUP_INT / BND signatures (source and destination locus)."""
    __slots__ = ("source_contig", "source_start", "source_end", "dest_contig", "dest_start", "dest_end", "score", "size", "members", "type",
                 "std_span", "std_pos")

    def __init__(self, source_contig, source_start, source_end, dest_contig, dest_start, dest_end, score, size, members,
                 type, std_span, std_pos):
        self.source_contig, self.source_start, self.source_end = source_contig, source_start, source_end
        self.dest_contig, self.dest_start, self.dest_end = dest_contig, dest_start, dest_end
        self.score, self.size, self.members, self.type = score, size, members, type
        self.std_span, self.std_pos = std_span, std_pos

    def get_source(self):
        return (self.source_contig, self.source_start, self.source_end)

    def get_destination(self):
        return (self.dest_contig, self.dest_start, self.dest_end)

    def get_source_length(self):
        return self.source_end - self.source_start

    def get_destination_length(self):
        return self.dest_end - self.dest_start

    def _members_text(self):
        return "[" + "][".join(m.as_string("|") for m in self.members) + "]"

    def get_bed_entries(self):
        src_name = "{0}_source;{1}:{2}-{3};{4};{5};{6}".format(self.type, self.dest_contig, self.dest_start, self.dest_end,
                                                              self.size, self.std_span, self.std_pos)
        dst_name = "{0}_dest;{1}:{2}-{3};{4}".format(self.type, self.source_contig, self.source_start, self.source_end, self.size)
        mem = self._members_text()
        src = "\t".join(str(x) for x in (self.source_contig, self.source_start, self.source_end, src_name, self.score, mem))
        dst = "\t".join(str(x) for x in (self.dest_contig, self.dest_start, self.dest_end, dst_name, self.score, mem))
        return (src, dst)

    def get_vcf_entry(self):
        if self.type != "DUP_TAN":
            return None
        info = "SVTYPE=DUP:TANDEM;END={0};SVLEN={1};STD_SPAN={2};STD_POS={3}".format(
            self.source_end, self.source_end - self.source_start, self.std_span, self.std_pos)
        return "\t".join(str(x) for x in (self.source_contig, self.source_start + 1, ".", "N", "<DUP:TANDEM>", ".", "PASS", info))
