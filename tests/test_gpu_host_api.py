"""-m gpu: the Python host mirror (same names / argument meaning as the reference's modules),
written after the reference's own tests (tests/test_clustering.py, tests/test_Collect.py)."""
import argparse
import random

import pytest

from svim_b200.SVSignature import SignatureDeletion, SignatureInsertion
from svim_b200.SVIM_clustering import form_partitions, partition_and_cluster
from svim_b200.SVIM_COLLECT import analyze_alignment_file_coordsorted
from svim_b200.SVIM_CLUSTER import cluster_sv_signatures
from svim_b200 import runtime

pytestmark = pytest.mark.gpu


def _options(genome_key="test-genome"):
    return argparse.Namespace(min_mapq=20, min_sv_size=40, max_sv_size=100000, segment_gap_tolerance=10, segment_overlap_tolerance=5,
                              partition_max_distance=1000, position_distance_normalizer=900, edit_distance_normalizer=1.0,
                              cluster_max_distance=0.5, all_bnds=False, genome=genome_key)


@pytest.fixture()
def float_deletions():
    # tests/test_clustering.py:11-30: three groups of ten deletions with FLOAT coordinates
    rng = random.Random(0)
    sigs = []
    for grp, (center0, half0) in enumerate([(100000, 1000), (200000, 1000), (100000, 2000)]):
        for i in range(10):
            center = center0 + rng.uniform(-100, 100)
            half = half0 + rng.uniform(-100, 100)
            sigs.append(SignatureDeletion("chr1", center - half, center + half, "cigar", str(grp * 10 + i)))
    from svim_b200 import synth
    runtime.register_genome("test-genome", synth.random_genome(["chr1"], [300000], 1))
    return sigs


def test_partitioning(float_deletions):
    parts = form_partitions(float_deletions, 100)
    assert len(parts) == 2
    for p in parts:
        assert {int(m.read) // 10 for m in p} in ({0, 2}, {1})
    parts = form_partitions(float_deletions, 100000)
    assert len(parts) == 1 and {int(m.read) // 10 for m in parts[0]} == {0, 1, 2}


def test_clustering_and_scores(float_deletions):
    clusters = partition_and_cluster(float_deletions, options=_options(), type="deleted regions")
    assert len(clusters) == 3
    for c in clusters:
        assert len({int(m.read) // 10 for m in c.members}) == 1
        assert 10 <= c.score <= 10 + 20 / 8
    # identical to the oracle on the same float inputs (order, membership, rounded coordinates)
    from oracle import svim_oracle as orc
    osigs = [orc.Sig("DEL", s.contig, s.start, s.end, "cigar", s.read) for s in float_deletions]
    want = orc.partition_and_cluster(osigs, None, orc.Params())
    assert [(c.start, c.end, [m.read for m in c.members]) for c in clusters] == [(c.start, c.end, [m.read for m in c.members]) for c in want]
    for a, b in zip(clusters, want):
        assert a.score == pytest.approx(b.score, rel=1e-9) and a.std_span == pytest.approx(b.std_span, rel=1e-9)


def test_collect_then_cluster_object_surface(golden):
    from oracle import svim_oracle as orc
    batch, genome, exp = golden("mini_mixed")
    runtime.register_genome("mini-mixed-genome", genome)
    opts = _options("mini-mixed-genome")
    sigs, twins = analyze_alignment_file_coordsorted(batch, opts)
    assert twins == [] and len(sigs) == len(exp["signatures"])
    # attribute surface the downstream stages read
    for s, row in zip(sigs, exp["signatures"]):
        assert s.type == row[0] and s.signature == row[11] and s.read == row[12]
        assert s.get_source()[0] == row[1]
        if s.type == "BND":
            assert (s.contig1, s.pos1, s.direction1, s.contig2, s.pos2, s.direction2) == (row[1], row[2], row[6], row[4], row[5], row[7])
        elif s.type == "DUP_INT":
            assert (s.contig1, s.start, s.end, s.contig2, s.pos) == (row[1], row[2], row[3], row[4], row[5])
        else:
            assert (s.contig, s.start, s.end) == (row[1], row[2], row[3])
        if s.type == "INS":
            assert s.sequence == row[13]
    res = cluster_sv_signatures(sigs, opts)                  # device-resident fast path
    assert isinstance(res, tuple) and len(res) == 6 and all(type(x) is list for x in res)
    names = ("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND")
    index_of = {id(s): i for i, s in enumerate(sigs)}
    for name, cl in zip(names, res):
        want = exp["clusters"][name]
        assert [[index_of[id(m)] for m in c.members] for c in cl] == [w[13] for w in want]
        assert [c.type for c in cl] == [name] * len(want)
    # the generic path (list was copied => re-marshalled and uploaded) gives the same answer
    res2 = cluster_sv_signatures(list(sigs), opts)
    for a, b in zip(res, res2):
        assert [[index_of[id(m)] for m in c.members] for c in a] == [[index_of[id(m)] for m in c.members] for c in b]
        assert [c.score for c in a] == [c.score for c in b]
    bnd = res[5]
    assert all(hasattr(c, "direction1") and hasattr(c, "direction2") for c in bnd)
    assert all(len(c.get_bed_entries()) == 2 for c in res[3] + res[4] + res[5])
    assert all(c.get_bed_entry().count("\t") == 5 for c in res[0] + res[1] + res[2])


def test_candidate_clustering_twin_on_gpu():
    """partition_and_cluster_candidates (SVIM_clustering.py:306-372) through the same kernels as the signature path."""
    import gzip, json, os
    from conftest import GOLDEN
    from svim_b200.SVIM_clustering import partition_and_cluster_candidates, CandidateDuplicationInterspersed, form_partitions
    g = json.load(gzip.open(os.path.join(GOLDEN, "candidates.golden.json.gz"), "rt"))
    cands = [CandidateDuplicationInterspersed(*r[:10], cutpaste=r[10]) for r in g["input"]]
    got = partition_and_cluster_candidates(cands, _options(), "interspersed duplication candidates")
    rows = [[c.source_contig, c.source_start, c.source_end, c.dest_contig, c.dest_start, c.dest_end, c.members, c.score, c.std_span, c.std_pos,
             c.cutpaste] for c in got]
    assert rows == g["output"]
    # form_partitions on candidates (used by COMBINE, SVIM_COMBINE.py:13): same partitions as the oracle's sort + gap split
    from oracle import svim_oracle as orc
    parts = form_partitions(cands, 1000)
    ocands = [orc.Cand(*r) for r in g["input"]]
    want = []
    for c in sorted(range(len(ocands)), key=lambda i: ocands[i].key()):
        if want and ocands[want[-1][-1]].gap_to(ocands[c]) <= 1000:
            want[-1].append(c)
        else:
            want.append([c])
    index_of = {id(c): i for i, c in enumerate(cands)}
    assert [[index_of[id(c)] for c in p] for p in parts] == want
