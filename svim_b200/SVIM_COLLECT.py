"""COLLECT host mirror: `analyze_alignment_file_coordsorted(bam, options)` of the
reference (SVIM_COLLECT.py:132-167) over the CUDA path.

`bam` is a flattened record buffer (`records.AlignmentBatch`), a SAM/BAM path, or
any object with a `.batch` attribute.  Returns `(sv_signatures,
translocation_signatures_all_bnds)` exactly like the reference: two Python lists of
`SVSignature` objects in the reference's emission order.  The first list is a
`SignatureList`: as long as it is passed on unmodified, `cluster_sv_signatures`
clusters the device-resident records instead of re-uploading them.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _lib, runtime
from .io import read_alignments
from .records import AlignmentBatch
from .SVSignature import (SignatureDeletion, SignatureInsertion, SignatureInversion, SignatureDuplicationTandem,
                          SignatureInsertionFrom, SignatureTranslocation)


class SignatureList(list):
    """A list that remembers it mirrors device-resident records until it is mutated."""
    _MUTATORS = ("append", "extend", "insert", "pop", "remove", "clear", "sort", "reverse", "__setitem__", "__delitem__",
                 "__iadd__", "__imul__")

    def __init__(self, items=(), token=None):
        super().__init__(items)
        self._svimgpu_token = token


def _drop_token(name):
    base = getattr(list, name)

    def method(self, *a, **k):
        self._svimgpu_token = None
        return base(self, *a, **k)
    method.__name__ = name
    return method


for _m in SignatureList._MUTATORS:
    setattr(SignatureList, _m, _drop_token(_m))


def as_batch(bam) -> AlignmentBatch:
    if isinstance(bam, AlignmentBatch):
        return bam
    if isinstance(bam, str):
        return read_alignments(bam)
    if hasattr(bam, "batch"):
        return bam.batch
    if hasattr(bam, "_batch"):
        return bam._batch
    raise TypeError("expected an AlignmentBatch, a SAM/BAM path or an object with .batch")


def materialize_signatures(sigs: np.ndarray, ins: np.ndarray, batch: AlignmentBatch):
    """svim_sig records (+ INS blob) -> SVSignature objects."""
    names = batch.contig_names
    qname = batch.qname
    ins_b = ins.tobytes()
    out = []
    cols = [sigs[f].tolist() for f in ("type", "flags", "contig1", "start", "end", "contig2", "pos", "qname_id", "seq_off",
                                      "seq_len", "copies")]
    for t, fl, c1, s, e, c2, pos, qid, so, sl, cp in zip(*cols):
        src = "suppl" if fl & _lib.F_SUPPL else "cigar"
        read = qname(qid)
        if t == 0:
            o = SignatureDeletion(names[c1], s, e, src, read)
        elif t == 1:
            o = SignatureInsertion(names[c1], s, e, src, read, ins_b[so:so + sl].decode("ascii"))
        elif t == 2:
            o = SignatureInversion(names[c1], s, e, src, read, _lib.INV_DIRECTIONS[(fl >> _lib.F_INVDIR_SHIFT) & 7])
        elif t == 3:
            o = SignatureDuplicationTandem(names[c1], s, e, cp, bool(fl & _lib.F_FULLY_COVERED), src, read)
        elif t == 4:
            # records are already in canonical breakend order; bypass the constructor's swap
            o = SignatureTranslocation.__new__(SignatureTranslocation)
            o.contig1, o.pos1, o.direction1 = names[c1], s, "rev" if fl & _lib.F_DIR1_REV else "fwd"
            o.contig2, o.pos2, o.direction2 = names[c2], pos, "rev" if fl & _lib.F_DIR2_REV else "fwd"
            o.signature, o.read, o.type = src, read, "BND"
        else:
            o = SignatureInsertionFrom(names[c1], s, e, names[c2], pos, src, read)
        out.append(o)
    return out


def collect_arrays(batch: AlignmentBatch, options=None, ctx=None, resident=False):
    """Run the COLLECT kernels; returns (ctx, stats, (sigs, ins), (twin_sigs, twin_ins))."""
    ctx = ctx or runtime.context()
    ctx.set_params(_lib.Params.from_options(options))
    key = (id(batch), tuple(batch.contig_names))
    if getattr(ctx, "contigs_key", None) != tuple(batch.contig_names):
        ctx.set_contigs(batch.contig_names)
        ctx.contigs_key = tuple(batch.contig_names)
    if resident and getattr(ctx, "resident", None) == key:
        stats = ctx.collect()
    else:
        stats = ctx.collect_host(batch)
        ctx.resident = key
    if stats.n_data_errors:
        raise _lib.SvimGpuError(-5, "%d reads carry SA tags the reference would raise on "
                                    "(unknown contig / non-integer field / empty CIGAR)" % stats.n_data_errors)
    for _ in range(stats.n_sa_bad_fields):
        logging.warning('SA tag does not consist of 6 fields. This could be a sign of invalid characters (e.g. commas or '
                        'semicolons) in a chromosome name of the reference genome.')
    for _ in range(stats.n_no_read_length):
        logging.warning('Skipping alignment because pysam was unable to infer length of read from CIGAR string')
    main = ctx.fetch_signatures(0, stats)
    twins = ctx.fetch_signatures(1, stats) if stats.n_twin_signatures else (np.zeros(0, _lib.SIG_DTYPE), np.zeros(0, np.uint8))
    return ctx, stats, main, twins


def analyze_alignment_file_coordsorted(bam, options):
    batch = as_batch(bam)
    ctx, stats, (sigs, ins), (tsigs, tins) = collect_arrays(batch, options)
    token = object()
    ctx.collect_token = token
    ctx.collect_batch = batch
    main = SignatureList(materialize_signatures(sigs, ins, batch), token=token)
    main._svimgpu_which = 0
    twins = SignatureList(materialize_signatures(tsigs, tins, batch), token=token)
    twins._svimgpu_which = 1
    return main, twins
