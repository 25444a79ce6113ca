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
    """svim_sig records (+ INS blob) -> SVSignature objects, in record order.

    The per-object cost dominates once the kernels are fast, so objects are built by filling `__dict__` directly
    (same attributes the constructors set) from column lists, one comprehension per type."""
    n = len(sigs)
    if n == 0:
        return []
    import gc
    was_enabled = gc.isenabled()
    gc.disable()            # hundreds of thousands of small containers: generational GC passes would dominate
    try:
        return _materialize(sigs, ins, batch, n)
    finally:
        if was_enabled:
            gc.enable()


def _materialize(sigs, ins, batch, n):
    names = batch.contig_names
    typ = sigs["type"]
    flags = sigs["flags"]
    src_l = np.where(flags & _lib.F_SUPPL, "suppl", "cigar").tolist()
    qid = sigs["qname_id"]
    if batch.qnames is None:
        reads = ["read%d" % q for q in qid.tolist()]
    else:
        qn = batch.qnames
        reads = [qn[q] for q in qid.tolist()]
    c1 = [names[t] for t in sigs["contig1"].tolist()] if len(names) > 1 else [names[0]] * n
    start = sigs["start"].tolist(); end = sigs["end"].tolist()
    out = [None] * n
    new = object.__new__

    def build(cls, idx, dicts):
        for k, d in zip(idx, dicts):
            o = new(cls)
            o.__dict__ = d
            out[k] = o

    idx = np.nonzero(typ == 0)[0].tolist()
    build(SignatureDeletion, idx, [{"contig": c1[k], "start": start[k], "end": end[k], "signature": src_l[k], "read": reads[k], "type": "DEL"} for k in idx])
    idx = np.nonzero(typ == 1)[0].tolist()
    if idx:
        text = ins.tobytes().decode("ascii")
        so = sigs["seq_off"].tolist(); sl = sigs["seq_len"].tolist()
        build(SignatureInsertion, idx, [{"contig": c1[k], "start": start[k], "end": end[k], "signature": src_l[k], "read": reads[k],
                                         "sequence": text[so[k]:so[k] + sl[k]], "type": "INS"} for k in idx])
    idx = np.nonzero(typ == 2)[0].tolist()
    if idx:
        fl = flags.tolist()
        build(SignatureInversion, idx, [{"contig": c1[k], "start": start[k], "end": end[k], "signature": src_l[k], "read": reads[k], "type": "INV",
                                         "direction": _lib.INV_DIRECTIONS[(fl[k] >> _lib.F_INVDIR_SHIFT) & 7]} for k in idx])
    idx = np.nonzero(typ == 3)[0].tolist()
    if idx:
        fl = flags.tolist(); cp = sigs["copies"].tolist()
        build(SignatureDuplicationTandem, idx, [{"contig": c1[k], "start": start[k], "end": end[k], "copies": cp[k],
                                                 "fully_covered": bool(fl[k] & _lib.F_FULLY_COVERED), "signature": src_l[k], "read": reads[k],
                                                 "type": "DUP_TAN"} for k in idx])
    idx = np.nonzero(typ >= 4)[0].tolist()
    if idx:
        fl = flags.tolist(); pos = sigs["pos"].tolist(); c2 = sigs["contig2"].tolist(); tl = typ.tolist()
        for k in idx:
            if tl[k] == 4:      # records are already in canonical breakend order (no constructor swap)
                o = new(SignatureTranslocation)
                o.__dict__ = {"contig1": c1[k], "pos1": start[k], "direction1": "rev" if fl[k] & _lib.F_DIR1_REV else "fwd",
                              "contig2": names[c2[k]], "pos2": pos[k], "direction2": "rev" if fl[k] & _lib.F_DIR2_REV else "fwd",
                              "signature": src_l[k], "read": reads[k], "type": "BND"}
            else:
                o = new(SignatureInsertionFrom)
                o.__dict__ = {"contig1": c1[k], "start": start[k], "end": end[k], "contig2": names[c2[k]], "pos": pos[k],
                              "signature": src_l[k], "read": reads[k], "type": "DUP_INT"}
            out[k] = o
    return out


def bam_iterator(bam):
    """bam_iterator (SVIM_COLLECT.py:8-41) over the flattened buffer: yields (primary, supplementary, secondary) lists of
    RECORD INDICES for each run of consecutive records with the same read name (host helper, nothing is computed)."""
    batch = as_batch(bam)
    qid = batch.qname_id; flag = batch.flag
    i = 0
    while i < batch.n:
        j = i
        prim, sup, sec = [], [], []
        while j < batch.n and qid[j] == qid[i]:
            f = int(flag[j])
            (sec if f & 0x100 else sup if f & 0x800 else prim).append(j)
            j += 1
        yield (prim, sup, sec)
        i = j


def collect_arrays(batch: AlignmentBatch, options=None, ctx=None, querysorted=False):
    """Run the COLLECT kernels from host buffers; returns (ctx, stats, (sigs, ins), (twin_sigs, twin_ins))."""
    ctx = ctx or runtime.context()
    ctx.set_params(_lib.Params.from_options(options))
    if getattr(ctx, "contigs_key", None) != tuple(batch.contig_names):
        ctx.set_contigs(batch.contig_names)
        ctx.contigs_key = tuple(batch.contig_names)
    stats = ctx.collect_host_querysorted(batch) if querysorted else ctx.collect_host(batch)
    ctx.resident = batch            # rows + CIGAR stay in HBM (GENOTYPE reuses them, svim_b200/SVIM_genotyping.py)
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


def analyze_alignment_file_querysorted(bam, options):
    """analyze_alignment_file_querysorted (SVIM_COLLECT.py:96-129) over the CUDA path."""
    return _analyze(bam, options, True)


def analyze_alignment_file_coordsorted(bam, options):
    return _analyze(bam, options, False)


def _analyze(bam, options, querysorted):
    batch = as_batch(bam)
    ctx, stats, (sigs, ins), (tsigs, tins) = collect_arrays(batch, options, querysorted=querysorted)
    token = object()
    ctx.collect_token = token
    ctx.collect_batch = batch
    main = SignatureList(materialize_signatures(sigs, ins, batch), token=token)
    main._svimgpu_which = 0
    twins = SignatureList(materialize_signatures(tsigs, tins, batch), token=token)
    twins._svimgpu_which = 1
    return main, twins
