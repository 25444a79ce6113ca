"""COMBINE-stage cut&paste search: `flag_cutpaste_candidates(insertion_from_signature_clusters,
deletion_signature_clusters, options)` of the reference (SVIM_merging.py:12-29) over the CUDA path
(`svimgpu_closest_source`).  The O(#DUP_INT x #DEL) distance search runs on the GPU; this module builds the
`CandidateDuplicationInterspersed` objects (the reference's own class when it is importable) exactly as :21-28 do.
"""
from __future__ import annotations

import numpy as np

from . import runtime
from .SVIM_clustering import _candidate_class


def closest_deletion(insertion_from_signature_clusters, deletion_signature_clusters, options, ctx=None):
    """-> (index array, distance array): per DUP_INT cluster the first-closest deletion cluster (SVIM_merging.py:17-20)."""
    ins_src = [c.get_source() for c in insertion_from_signature_clusters]
    del_src = [c.get_source() for c in deletion_signature_clusters]
    ctx = ctx or runtime.context()
    return ctx.closest_source([s[1] for s in ins_src], [s[2] for s in ins_src], [s[1] for s in del_src], [s[2] for s in del_src],
                              options.position_distance_normalizer)


def flag_cutpaste_candidates(insertion_from_signature_clusters, deletion_signature_clusters, options):
    """Flag duplication signature clusters if they overlap a deletion."""
    if len(insertion_from_signature_clusters) == 0:
        return []
    if len(deletion_signature_clusters) == 0:
        raise IndexError("list index out of range")          # sorted([])[0] at SVIM_merging.py:20
    try:
        _idx, dist = closest_deletion(insertion_from_signature_clusters, deletion_signature_clusters, options)
    except Exception as e:
        if getattr(e, "code", None) == -5:
            raise ZeroDivisionError("division by zero") from None
        raise
    cls = _candidate_class()
    near = (dist <= options.del_ins_dup_max_distance).tolist()
    out = []
    for ins_cluster, cutpaste in zip(insertion_from_signature_clusters, near):
        source_contig, source_start, source_end = ins_cluster.get_source()
        dest_contig, dest_start, dest_end = ins_cluster.get_destination()
        out.append(cls(source_contig, source_start, source_end, dest_contig, dest_start, dest_end, ins_cluster.members, ins_cluster.score,
                       ins_cluster.std_span, ins_cluster.std_pos, cutpaste=cutpaste))
    return out
