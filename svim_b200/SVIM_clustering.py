"""CLUSTER host mirror (SVIM_clustering.py of the reference): `form_partitions`,
`partition_and_cluster` and the machinery `cluster_sv_signatures` shares with them.

Arbitrary lists of signature objects (the reference's public seam — its unit tests
call `partition_and_cluster` with float coordinates, tests/test_clustering.py:53) are
marshalled into `svim_csig` records and clustered by the CUDA kernels; nothing is
computed on the host except the marshalling and the object construction.
"""
from __future__ import annotations

import logging
import math

import numpy as np

from . import _lib, runtime
from .SVSignature import SignatureClusterUniLocal, SignatureClusterBiLocal

_INV_CODE = {d: i for i, d in enumerate(_lib.INV_DIRECTIONS)}
_LABEL = {"DEL": "deleted regions", "INS": "inserted regions", "INV": "inverted regions",
          "DUP_TAN": "tandem duplicated regions", "BND": "translocation breakpoints",
          "DUP_INT": "inserted regions with detected region of origin"}


def marshal_signatures(signatures):
    """Signature objects -> (csig array, INS blob, contig names sorted = rank order)."""
    n = len(signatures)
    cs = np.zeros(n, dtype=_lib.CSIG_DTYPE)
    contigs = set()
    for s in signatures:
        contigs.add(s.get_source()[0])
        if s.type in ("DUP_INT", "BND"):
            contigs.add(s.get_destination()[0])
    names = sorted(contigs)
    rank = {c: i for i, c in enumerate(names)}
    reads = {}
    blob = bytearray()
    for k, s in enumerate(signatures):
        t = s.type
        r = cs[k]
        c, st, en = s.get_source()
        r["type"] = _lib.TYPE_CODE[t]
        r["start"], r["end"] = st, en
        r["contig_a"] = rank[c]
        r["contig_b"] = -1
        r["read_id"] = reads.setdefault(s.read, len(reads))
        if t == "INS":
            seq = s.sequence.encode("latin-1")
            r["seq_off"], r["seq_len"] = len(blob), len(seq)
            blob += seq
        elif t == "INV":
            r["dirs"] = _INV_CODE[s.direction]
        elif t == "DUP_TAN":
            r["copies"] = min(int(s.copies), 65535)
        elif t == "DUP_INT":
            dc, ds, _ = s.get_destination()
            r["dpos"], r["contig_b"] = ds, rank[dc]
        elif t == "BND":
            dc, ds, _ = s.get_destination()
            r["dpos"], r["contig_b"] = ds, rank[dc]
            r["dirs"] = (1 if s.direction1 == "rev" else 0) | (2 if s.direction2 == "rev" else 0)
    return cs, np.frombuffer(bytes(blob), dtype=np.uint8), names


def _none_if_nan(x):
    return None if math.isnan(x) else x


def build_clusters(clusters: np.ndarray, members: np.ndarray, signatures):
    """svim_cluster records -> per-type lists of SignatureCluster objects (enum order), built in C (csrc_host/fastobj.c)."""
    from . import fastobj
    fo = fastobj.module()
    if not isinstance(signatures, list):
        signatures = list(signatures)
    import gc
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        return fo.clusters(np.ascontiguousarray(clusters).view(np.uint8), np.ascontiguousarray(members, dtype=np.uint32).view(np.uint8), signatures,
                           SignatureClusterUniLocal, SignatureClusterBiLocal, tuple(_lib.TYPE_NAMES))
    finally:
        if was_enabled:
            gc.enable()


def build_clusters_py(clusters, members, signatures):
    """Reference implementation of `build_clusters` in Python (test oracle for the C loop; not on the product path)."""
    out = [[] for _ in range(6)]
    mem = members.tolist()
    cols = [clusters[f].tolist() for f in ("type", "start", "end", "dest_start", "dest_end", "score", "std_span", "std_pos",
                                          "member_off", "size", "dir1_rev", "dir2_rev")]
    for t, s, e, ds, de, score, sd_span, sd_pos, off, size, d1, d2 in zip(*cols):
        ms = [signatures[i] for i in mem[off:off + size]]
        name = _lib.TYPE_NAMES[t]
        first = ms[0]
        sd_span, sd_pos = _none_if_nan(sd_span), _none_if_nan(sd_pos)
        if t <= 2:
            c = SignatureClusterUniLocal(first.get_source()[0], s, e, score, size, ms, name, sd_span, sd_pos)
        elif t == 3:
            contig = first.get_source()[0]
            c = SignatureClusterBiLocal(contig, s, e, contig, ds, de, score, size, ms, name, sd_span, sd_pos)
        else:
            c = SignatureClusterBiLocal(first.get_source()[0], s, e, first.get_destination()[0], ds, de, score, size, ms, name,
                                        sd_span, sd_pos)
            if t == 4:
                c.direction1 = "rev" if d1 else "fwd"
                c.direction2 = "rev" if d2 else "fwd"
        out[t].append(c)
    return out


def _log_stats(stats, types):
    for t in types:
        logging.debug("%d out of %d partitions for %s exceeded 100 elements." % (stats.large_partitions[t], stats.n_partitions[t], _lib.TYPE_NAMES[t]))
        logging.debug("%d %s signatures were removed due to similarity to another signature from the same read." % (stats.duplicate_signatures[t], _lib.TYPE_NAMES[t]))
        logging.info("Clustered {0}: {1} partitions and {2} clusters".format(_LABEL[_lib.TYPE_NAMES[t]], stats.n_partitions[t], stats.n_clusters[t]))


def cluster_objects(signatures, options, ctx=None):
    """Upload an arbitrary signature list and cluster it; returns (ctx, stats, per-type lists)."""
    ctx = ctx or runtime.context()
    ctx.set_params(_lib.Params.from_options(options))
    cs, blob, names = marshal_signatures(signatures)
    rank_to_tid = None
    if (cs["type"] == 1).any():
        genome = runtime.genome_for(options.genome)
        runtime.ensure_genome(ctx, genome, genome.names)
        idx = {n: i for i, n in enumerate(genome.names)}
        rank_to_tid = np.array([idx.get(n, -1) for n in names], dtype=np.int32)
    elif getattr(options, "genome", None) is not None:
        runtime.genome_for(options.genome)   # the reference opens the FASTA for every type (SVIM_clustering.py:377)
    ctx.set_signatures(cs, blob, rank_to_tid)
    stats, clusters, members = ctx.cluster(view=True)
    return ctx, stats, build_clusters(clusters, members, signatures)


def form_partitions(sv_signatures, max_distance):
    """form_partitions (SVIM_clustering.py:17-29) on the GPU: key sort + gap split."""
    if len(sv_signatures) == 0:
        return []
    ctx = runtime.context()
    ctx.set_params(_lib.Params.from_options(None, partition_max_distance=max_distance))
    if hasattr(sv_signatures[0], "source_contig") and hasattr(sv_signatures[0], "dest_contig") and not hasattr(sv_signatures[0], "read"):
        cs = marshal_candidates(sv_signatures)     # COMBINE partitions DUP_INT candidates with the same function
    else:
        cs, blob, names = marshal_signatures(sv_signatures)
    ctx.set_signatures(cs, None, None)       # partitions do not depend on the inserted sequences
    ctx.partition()
    order, off = ctx.fetch_partitions(len(cs))
    order = order.tolist(); off = off.tolist()
    return [[sv_signatures[i] for i in order[a:b]] for a, b in zip(off[:-1], off[1:])]


def partition_and_cluster(signatures, options, type):
    """partition_and_cluster (SVIM_clustering.py:375-385)."""
    known = set(_LABEL.values())
    if type not in known:
        logging.error("Unknown parameter type={0} to function partition_and_cluster.")
        return None
    if len(signatures) == 0:
        runtime.genome_for(options.genome)
        logging.info("Clustered {0}: {1} partitions and {2} clusters".format(type, 0, 0))
        return []
    ctx, stats, per_type = cluster_objects(signatures, options)
    present = sorted({_lib.TYPE_CODE[s.type] for s in signatures})
    _log_stats(stats, present)
    out = []
    for t in present:
        out.extend(per_type[t])
    return out


# ---- COMBINE-stage twin --------------------------------------------------------------
class CandidateDuplicationInterspersed:
    """Minimal stand-in used when the reference package is not importable: the attributes
    partition_and_cluster_candidates reads and writes (SVCandidate.py:424-453)."""
    type = "DUP_INT"

    def __init__(self, source_contig, source_start, source_end, dest_contig, dest_start, dest_end, members, score, std_span, std_pos,
                 cutpaste=False):
        self.source_contig, self.source_start, self.source_end = source_contig, max(0, source_start), source_end
        self.dest_contig, self.dest_start, self.dest_end = dest_contig, max(0, dest_start), dest_end
        self.members, self.score, self.std_span, self.std_pos, self.cutpaste = members, score, std_span, std_pos, cutpaste
        self.type = "DUP_INT"

    def get_source(self):
        return (self.source_contig, self.source_start, self.source_end)

    def get_destination(self):
        return (self.dest_contig, self.dest_start, self.dest_end)

    def get_key(self):
        return (self.type, self.source_contig, self.source_end)

    def downstream_distance_to(self, other):
        if self.type == other.type and self.source_contig == other.source_contig:
            return max(0, other.source_start - self.source_end)
        return float("inf")


def _candidate_class():
    try:
        from svim.SVCandidate import CandidateDuplicationInterspersed as ref_cls      # reference installed: return its own class
        return ref_cls
    except Exception:
        return CandidateDuplicationInterspersed


def marshal_candidates(candidates):
    """DUP_INT candidates -> svim_csig records of type SVIM_DUP_INT_CAND (unique read ids: no same-read rules)."""
    n = len(candidates)
    cs = np.zeros(n, dtype=_lib.CSIG_DTYPE)
    names = sorted({c.get_source()[0] for c in candidates} | {c.get_destination()[0] for c in candidates})
    rank = {c: i for i, c in enumerate(names)}
    dest_end = np.zeros(n, dtype=np.float64)
    for k, c in enumerate(candidates):
        sc, ss, se = c.get_source()
        dc, ds, de = c.get_destination()
        r = cs[k]
        r["type"] = _lib.TYPE_DUP_INT_CAND
        r["start"], r["end"], r["dpos"] = ss, se, ds
        r["contig_a"], r["contig_b"] = rank[sc], rank[dc]
        r["read_id"] = k
        dest_end[k] = de
    cs["seq_off"] = dest_end.view(np.uint64)          # bits of the destination end (svimgpu.h, SVIM_DUP_INT_CAND)
    return cs


def partition_and_cluster_candidates(candidates, options, type):
    """partition_and_cluster_candidates (SVIM_clustering.py:306-372): key sort, partitions, host-RNG sampling,
    span_position_distance_intdup_candidates matrix, average linkage and flat cut on the GPU; the merged
    candidate's fields (max score, concatenated members, mean of the stds, cut&paste flag) on the host."""
    from statistics import mean
    if len(candidates) == 0:
        logging.info("Clustered {0}: {1} partitions and {2} clusters".format(type, 0, 0))
        return []
    ctx = runtime.context()
    ctx.set_params(_lib.Params.from_options(options))
    ctx.set_signatures(marshal_candidates(candidates), None, None)
    stats, clusters, members = ctx.cluster(view=True)
    logging.debug("%d out of %d partitions for %s exceeded 100 elements." % (stats.large_partitions[5], stats.n_partitions[5], candidates[0].type))
    logging.info("Clustered {0}: {1} partitions and {2} clusters".format(type, stats.n_partitions[5], stats.n_clusters[5]))
    cls = _candidate_class()
    mem = members.tolist()
    final = []
    for off, size, s, e, ds, de in zip(*[clusters[f].tolist() for f in ("member_off", "size", "start", "end", "dest_start", "dest_end")]):
        group = [candidates[i] for i in mem[off:off + size]]
        if group[0].type != "DUP_INT":
            continue
        spans = [c.std_span for c in group if c.std_span is not None]
        poss = [c.std_pos for c in group if c.std_pos is not None]
        final.append(cls(group[0].get_source()[0], s, e, group[0].get_destination()[0], ds, de,
                         [m for c in group for m in c.members], max(c.score for c in group),
                         mean(spans) if spans else None, mean(poss) if poss else None, any(c.cutpaste for c in group)))
    return final
