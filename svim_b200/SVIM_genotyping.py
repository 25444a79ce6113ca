"""GENOTYPE host mirror: `genotype(candidates, bam, type, options)` of the reference
(SVIM_genotyping.py:34-93; called for DEL, INV, INS, DUP_INT at svim:161-170) over the CUDA path.

`candidates` are the reference's own `SVCandidate` objects (anything with `.score`, `.members[*].read`,
`get_source()` / `get_destination()`); the four attributes the reference writes — `support_fraction`, `genotype`,
`ref_reads`, `alt_reads` — are set in place, candidates below `options.minimum_score` are left untouched (:39-40).
`bam` is the flattened record buffer COLLECT ran on (an `AlignmentBatch`, a path, or an object with `.batch`): when it is
the batch whose rows are still in HBM from `analyze_alignment_file_coordsorted`, nothing is uploaded again.
This module only marshals; the region fetch, the read walk and the genotype decision run in `svimgpu_genotype`.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _lib, runtime
from .SVIM_COLLECT import as_batch

GENOTYPED_TYPES = ("DEL", "INV", "INS", "DUP_INT")


def _qname_index(batch):
    idx = getattr(batch, "_qname_index", None)
    if idx is None:
        if batch.qnames is None:
            idx = None
        else:
            idx = {n: i for i, n in enumerate(batch.qnames)}
        batch._qname_index = idx
    return idx


def _read_id(batch, index, name):
    """qname id of a read name, or None when no record carries that name (such a read can never be skipped at :63)."""
    if index is not None:
        return index.get(name)
    if name.startswith("read") and name[4:].isdigit() and len(name) <= 14:      # synthetic batches: names are "read<id>"
        i = int(name[4:])
        return i if i < 0xFFFFFFFF else None
    return None


def ensure_resident(ctx, batch):
    """Rows + CIGAR of `batch` in HBM: reuse what COLLECT left there, else upload."""
    if getattr(ctx, "resident", None) is batch:
        return
    ctx.upload(batch)
    ctx.resident = batch
    ctx.collect_batch = None        # the collected lists belonged to the previous buffer
    ctx.collect_token = None


def genotype_arrays(candidates, batch, type):
    """-> (GENO_CAND_DTYPE[n], variant id blob) for candidates that all pass the score filter."""
    index = _qname_index(batch)
    ins_like = type in ("INS", "DUP_INT")
    cands = np.zeros(len(candidates), dtype=_lib.GENO_CAND_DTYPE)
    blob = []
    off = 0
    for k, cand in enumerate(candidates):
        contig, start, end = cand.get_destination() if ins_like else cand.get_source()
        tid = batch.get_tid(contig)
        if tid < 0:
            raise KeyError(contig)                           # bam.get_reference_length raises (:48)
        names = set(sig.read for sig in cand.members)        # :51
        ids = sorted(i for i in (_read_id(batch, index, n) for n in names) if i is not None)
        cands[k] = (start, end, tid, len(names), off)
        # n_variant_reads counts NAMES (alt_reads, :93); names without a record contribute no id but still count
        if len(ids) != len(names):
            ids += [0xFFFFFFFF] * (len(names) - len(ids))    # sentinel: never equals a record's id
        blob.extend(ids)
        off += len(ids)
    return cands, np.asarray(blob, dtype=np.uint32)


def candidate_arrays_from_clusters(clusters, members, sigs, type_code, minimum_score=3):
    """Array form of the candidates COMBINE makes 1:1 from DEL / INS signature clusters (CandidateDeletion and, with
    --skip_consensus, CandidateNovelInsertion: SVIM_COMBINE.py:462-466, :265-273): locus = (contig, max(0, start), end) of the
    cluster, variant reads = distinct reads of its members.  Vectorised (no Python objects); used by bench.py to drive
    `svimgpu_genotype` at full scale.  -> (GENO_CAND_DTYPE[n], variant ids, indices of the clusters used)."""
    sel = np.nonzero((clusters["type"] == type_code) & (clusters["score"] > 0) & ~(clusters["score"] < minimum_score))[0]
    cl = clusters[sel]
    size = cl["size"].astype(np.int64)
    which = np.repeat(np.arange(len(cl), dtype=np.int64), size)
    pos_in = np.arange(int(size.sum()), dtype=np.int64) - np.repeat(np.cumsum(size) - size, size)
    mem = members[np.repeat(cl["member_off"].astype(np.int64), size) + pos_in]
    ids = sigs["qname_id"][mem].astype(np.int64)
    order = np.lexsort((ids, which))
    which_s, ids_s = which[order], ids[order]
    first = np.ones(len(ids_s), dtype=bool)
    first[1:] = (which_s[1:] != which_s[:-1]) | (ids_s[1:] != ids_s[:-1])
    n_var = np.bincount(which_s[first], minlength=len(cl)).astype(np.int64)
    cands = np.zeros(len(cl), dtype=_lib.GENO_CAND_DTYPE)
    cands["start"] = np.maximum(cl["start"], 0)
    cands["end"] = cl["end"]
    cands["tid"] = sigs["contig1"][members[cl["member_off"].astype(np.int64)]] if len(cl) else 0
    cands["n_variant_reads"] = n_var
    cands["variant_off"] = np.cumsum(n_var) - n_var
    return cands, ids_s[first].astype(np.uint32), sel


def genotype(candidates, bam, type, options):
    if type not in GENOTYPED_TYPES:
        raise ValueError("genotype() is defined for %s (svim:161-170), not %r" % (", ".join(GENOTYPED_TYPES), type))
    batch = as_batch(bam)
    if batch.sort_order != "coordinate":
        raise ValueError("fetch called on bamfile without index")        # what pysam raises for an unindexed file
    todo = [c for c in candidates if not (c.score < options.minimum_score)]      # :39-40
    if not todo:
        return
    ctx = runtime.context()
    ensure_resident(ctx, batch)
    cands, variant_ids = genotype_arrays(todo, batch, type)
    res = ctx.genotype(_lib.TYPE_CODE[type], _lib.GenoParams.from_options(options), cands, variant_ids, batch.contig_lengths)
    status = res["status"]
    if status.any():
        k = int(np.nonzero(status)[0][0])
        st = int(status[k])
        if st == 1:
            raise TypeError("'>' not supported between instances of 'NoneType' and 'int'")   # reference_end of a CIGAR-less record
        if st == 2:
            raise ValueError("invalid coordinates: start > stop for candidate at %s:%d" % (batch.contig_names[int(cands[k]["tid"])], int(cands[k]["start"])))
        raise ZeroDivisionError("division by zero")
    frac = res["support_fraction"].tolist(); gt = res["genotype"].tolist(); rr = res["ref_reads"].tolist(); ar = res["alt_reads"].tolist()
    for k, cand in enumerate(todo):
        f = frac[k]
        cand.support_fraction = "." if f != f else f
        cand.genotype = _lib.GENOTYPES[gt[k]]
        cand.ref_reads = rr[k]
        cand.alt_reads = ar[k]
    n = len(candidates)
    if n >= 10000:
        logging.info("Processed {0} of {1} candidates".format(n - n % 10000, n))
