"""-m gpu: BASELINE.json configs[0] — a 1-contig, 1 000-read synthetic coordinate-sorted BAM goes through the
BAM writer/reader and the reference-shaped host API; results must equal the oracle bit for bit."""
import argparse

import pytest

from svim_b200 import synth, runtime
from svim_b200.io import write_bam, read_alignments
from svim_b200.SVIM_COLLECT import analyze_alignment_file_coordsorted
from svim_b200.SVIM_CLUSTER import cluster_sv_signatures

pytestmark = pytest.mark.gpu


def test_config1_bam_to_clusters(tmp_path):
    from oracle import svim_oracle as orc
    batch, genome, _ = synth.make_config("config1")
    bam = str(tmp_path / "config1.bam")
    fa = str(tmp_path / "genome.fa")
    write_bam(bam, batch)
    genome.write_fasta(fa)
    opts = argparse.Namespace(min_mapq=20, min_sv_size=40, max_sv_size=100000, segment_gap_tolerance=10, segment_overlap_tolerance=5,
                              partition_max_distance=1000, position_distance_normalizer=900, edit_distance_normalizer=1.0,
                              cluster_max_distance=0.5, all_bnds=False, genome=fa, bam_file=bam)
    decoded = read_alignments(bam)
    assert decoded.n == batch.n
    sigs, twins = analyze_alignment_file_coordsorted(bam, opts)           # path -> BAM decode -> GPU
    res = cluster_sv_signatures(sigs, opts)
    p = orc.Params()
    osigs, _ = orc.collect(decoded, p)
    assert len(sigs) == len(osigs) > 100
    for a, b in zip(sigs, osigs):
        assert (a.type, a.get_source(), a.signature, a.read) == (b.type, b.source(), b.signature, b.read)
        if a.type == "INS":
            assert a.sequence == b.sequence
    want = orc.cluster(osigs, genome, p)
    index_a = {id(s): i for i, s in enumerate(sigs)}
    index_b = {id(s): i for i, s in enumerate(osigs)}
    for got, exp in zip(res, want):
        assert [(c.type, [index_a[id(m)] for m in c.members]) for c in got] == [(c.type, [index_b[id(m)] for m in c.members]) for c in exp]
        for c, e in zip(got, exp):
            if hasattr(c, "start"):
                assert (c.contig, c.start, c.end) == (e.contig, e.start, e.end)
            assert c.score == pytest.approx(e.score, rel=1e-6)
    assert len(res[0]) > 5 and len(res[1]) > 5


def test_cli_writes_signature_beds(tmp_path):
    import subprocess, sys, os
    from conftest import ROOT
    batch, genome, _ = synth.make_config("config1", 0.3)
    bam = str(tmp_path / "c.bam"); fa = str(tmp_path / "g.fa"); wd = str(tmp_path / "out")
    write_bam(bam, batch); genome.write_fasta(fa)
    r = subprocess.run([sys.executable, "-m", "svim_b200", "alignment", wd, bam, fa], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    dels = open(os.path.join(wd, "signatures", "del.bed")).read().splitlines()
    assert len(dels) > 3 and all(l.split("\t")[3].startswith("DEL;") for l in dels)
    assert os.path.exists(os.path.join(wd, "signatures", "trans.bed"))
