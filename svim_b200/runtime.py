"""Process-wide GPU context and genome registry shared by the host mirror modules."""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from .io import Genome

_CTX = {}
_GENOMES = {}


def context(device: int = None) -> "_lib.Context":
    if device is None:
        device = int(os.environ.get("SVIM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _CTX.get(device)
    if ctx is None:
        ctx = _lib.Context(device=device)
        _CTX[device] = ctx
    pre = getattr(ctx, "cluster_prefetch", None)
    if pre is not None:           # a CLUSTER started behind COLLECT's object building (SVIM_COLLECT._ClusterPrefetch) owns the context until it ends
        pre.t.join()
    return ctx


def shutdown():
    for c in _CTX.values():
        c.close()
    _CTX.clear()


def register_genome(path: str, genome: Genome):
    """Make an in-memory genome available under `options.genome == path`."""
    _GENOMES[os.path.abspath(path) if os.path.sep in path else path] = genome


def genome_for(path: str) -> Genome:
    key = os.path.abspath(path) if os.path.sep in path else path
    g = _GENOMES.get(key) or _GENOMES.get(path)
    if g is None:
        g = Genome.from_fasta(path)     # FileNotFoundError like pysam.FastaFile on a missing file
        _GENOMES[key] = g
    return g


def ensure_genome(ctx, genome: Genome, contig_names):
    """Upload `genome` laid out in the order of `contig_names` (BAM header order);
    contigs missing from the FASTA become empty."""
    key = (id(genome), tuple(contig_names))
    if ctx.genome_key == key:
        return
    idx = {n: i for i, n in enumerate(genome.names)}
    if list(genome.names) == list(contig_names):
        ordered = genome
    else:
        seqs = []
        for n in contig_names:
            i = idx.get(n)
            seqs.append(genome.blob[genome.offsets[i]:genome.offsets[i + 1]] if i is not None else np.zeros(0, np.uint8))
        ordered = Genome(list(contig_names), seqs)
    ctx.set_genome(ordered)
    ctx.genome_key = key
