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


_SIG_CLASSES = (SignatureDeletion, SignatureInsertion, SignatureInversion, SignatureDuplicationTandem, SignatureTranslocation,
                SignatureInsertionFrom)          # enum order of include/svimgpu.h


def materialize_signatures(sigs: np.ndarray, ins: np.ndarray, batch: AlignmentBatch):
    """svim_sig records (+ INS blob) -> SVSignature objects, in record order.

    The per-object cost dominates once the kernels are fast (380 ms of Python for 257 k signatures on BASELINE configs[1]), so
    the objects are built in one C loop (svim_b200/csrc_host/fastobj.c: tp_alloc + stores into the classes' __slots__).
    `materialize_signatures_py` is the same thing in Python; the tests compare the two attribute by attribute."""
    if len(sigs) == 0:
        return []
    from . import fastobj
    fo = fastobj.module()
    import gc
    was_enabled = gc.isenabled()
    gc.disable()            # hundreds of thousands of small containers: generational GC passes would dominate
    try:
        return fo.signatures(np.ascontiguousarray(sigs).view(np.uint8), np.ascontiguousarray(ins), list(batch.contig_names), batch.qnames,
                             _SIG_CLASSES, tuple(_lib.INV_DIRECTIONS))
    finally:
        if was_enabled:
            gc.enable()


def materialize_signatures_py(sigs: np.ndarray, ins: np.ndarray, batch: AlignmentBatch):
    """Reference implementation of `materialize_signatures` in Python (test oracle for the C loop; not on the product path)."""
    names = batch.contig_names
    text = ins.tobytes().decode("ascii")
    out = []
    for s in sigs:
        t = int(s["type"]); fl = int(s["flags"])
        src = "suppl" if fl & _lib.F_SUPPL else "cigar"
        read = batch.qname(int(s["qname_id"]))
        c1 = names[int(s["contig1"])]; st, en = int(s["start"]), int(s["end"])
        if t == 0:
            o = SignatureDeletion(c1, st, en, src, read)
        elif t == 1:
            o = SignatureInsertion(c1, st, en, src, read, text[int(s["seq_off"]):int(s["seq_off"]) + int(s["seq_len"])])
        elif t == 2:
            o = SignatureInversion(c1, st, en, src, read, _lib.INV_DIRECTIONS[(fl >> _lib.F_INVDIR_SHIFT) & 7])
        elif t == 3:
            o = SignatureDuplicationTandem(c1, st, en, int(s["copies"]), bool(fl & _lib.F_FULLY_COVERED), src, read)
        elif t == 4:      # records are already in canonical breakend order (SVSignature.py:193-214 ran on the device): no constructor swap
            o = object.__new__(SignatureTranslocation)
            o.contig1, o.pos1, o.direction1 = c1, st, "rev" if fl & _lib.F_DIR1_REV else "fwd"
            o.contig2, o.pos2, o.direction2 = names[int(s["contig2"])], int(s["pos"]), "rev" if fl & _lib.F_DIR2_REV else "fwd"
            o.signature, o.read = src, read
        else:
            o = SignatureInsertionFrom(c1, st, en, names[int(s["contig2"])], int(s["pos"]), src, read)
        out.append(o)
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


class _ClusterPrefetch:
    """CLUSTER of the list COLLECT is about to return, started on a host thread while the Python objects are built.

    `svim` calls cluster_sv_signatures(sv_signatures, options) right after analyze_alignment_file_* with the same options object
    (svim:102,132).  The kernels need nothing from the host but the options, so the device works on CLUSTER (ctypes releases the
    GIL for the call) while the interpreter materialises a quarter of a million Signature objects; cluster_sv_signatures takes the
    finished records if list, options and genome are still the ones this was started with, and recomputes otherwise."""

    def __init__(self, ctx, options, which):
        import threading
        self.key = self.options_key(options)
        self.which = which
        self.result = self.error = None
        self.t = threading.Thread(target=self._run, args=(ctx, options), daemon=True)
        self.t.start()

    @staticmethod
    def options_key(options):
        return tuple((k, getattr(options, k, None)) for k in ("partition_max_distance", "position_distance_normalizer", "edit_distance_normalizer",
                                                                "cluster_max_distance", "genome"))

    def _run(self, ctx, options):
        try:
            genome = runtime.genome_for(options.genome)
            runtime.ensure_genome(ctx, genome, ctx.collect_batch.contig_names)
            ctx.use_collected(self.which)
            self.result = ctx.cluster(view=True)       # consumed by build_clusters before the next cluster on this context
        except BaseException as e:          # surfaces (again) when cluster_sv_signatures recomputes
            self.error = e

    def take(self, options, which):
        self.t.join()
        if self.error is not None or which != self.which or self.options_key(options) != self.key:
            return None
        return self.result


def _analyze(bam, options, querysorted):
    try:
        return _analyze_inner(bam, options, querysorted)
    except KeyboardInterrupt:
        # The reference stops collecting at the record it had reached and goes on with what it has (SVIM_COLLECT.py:126-128, 164-166).
        # COLLECT here is one device pass of a fraction of a second, so the granularity is the whole pass: an interrupt that
        # arrives before its records are on the host leaves nothing to go on with.
        logging.warning('Execution interrupted by user. Stop detection and continue with next step..')
        return SignatureList([]), SignatureList([])


def _analyze_inner(bam, options, querysorted):
    batch = as_batch(bam)
    ctx, stats, (sigs, ins), (tsigs, tins) = collect_arrays(batch, options, querysorted=querysorted)
    token = object()
    ctx.collect_token = token
    ctx.collect_batch = batch
    ctx.cluster_prefetch = None
    if len(sigs) and getattr(options, "genome", None) is not None:
        # (sigs / ins are views of the context's pinned mirrors: use_collected + cluster never touch those, only the next collect does)
        ctx.cluster_prefetch = _ClusterPrefetch(ctx, options, 0)
    main = SignatureList(materialize_signatures(sigs, ins, batch), token=token)
    main._svimgpu_which = 0
    twins = SignatureList(materialize_signatures(tsigs, tins, batch), token=token)
    twins._svimgpu_which = 1
    return main, twins
