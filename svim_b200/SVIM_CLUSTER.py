"""CLUSTER façade: `cluster_sv_signatures(sv_signatures, options)` of the reference
(SVIM_CLUSTER.py:7-26).  Returns the 6-tuple in the reference's return order
(DEL, INS, INV, DUP_TAN, DUP_INT, BND) of plain, mutable Python lists."""
from __future__ import annotations

from . import _lib, runtime
from .SVIM_COLLECT import SignatureList
from .SVIM_clustering import build_clusters, cluster_objects, _log_stats


def cluster_sv_signatures(sv_signatures, options):
    ctx = runtime.context()
    token = getattr(sv_signatures, "_svimgpu_token", None)
    if isinstance(sv_signatures, SignatureList) and token is not None and token is getattr(ctx, "collect_token", None):
        # the list is exactly what COLLECT returned: cluster the device-resident records
        which = getattr(sv_signatures, "_svimgpu_which", 0)
        pre = getattr(ctx, "cluster_prefetch", None)
        ctx.cluster_prefetch = None
        done = pre.take(options, which) if pre is not None else None      # CLUSTER already ran while COLLECT's objects were being built
        if done is None:
            ctx.set_params(_lib.Params.from_options(options))
            batch = ctx.collect_batch
            genome = runtime.genome_for(options.genome)
            runtime.ensure_genome(ctx, genome, batch.contig_names)
            ctx.use_collected(which)
            done = ctx.cluster(view=True)
        stats, clusters, members = done
        per_type = build_clusters(clusters, members, sv_signatures)
    else:
        if len(sv_signatures) == 0:
            runtime.genome_for(options.genome)
            return ([], [], [], [], [], [])
        ctx, stats, per_type = cluster_objects(list(sv_signatures), options, ctx)
    _log_stats(stats, range(6))
    d, i, v, t, b, f = per_type
    return (d, i, v, t, f, b)
