"""GENOTYPE (SVIM_genotyping.py:34-93) on the GPU through the host mirror `svim_b200.SVIM_genotyping.genotype` (C ABI
`svimgpu_genotype`) against the reference's outputs committed under tests/golden/geno_* and against the oracle."""
import types

import numpy as np
import pytest

from conftest import GOLDEN_GENOTYPE
from oracle import svim_oracle as orc

pytestmark = pytest.mark.gpu
GENO_TYPES = ("DEL", "INV", "INS", "DUP_INT")


class Member:
    __slots__ = ("read",)

    def __init__(self, read):
        self.read = read


class Candidate:
    """The slice of SVCandidate.py genotype() touches: score, members[*].read, get_source / get_destination."""

    def __init__(self, contig, start, end, score, reads):
        self.locus = (contig, start, end)
        self.score = score
        self.members = [Member(r) for r in reads]
        self.support_fraction, self.genotype, self.ref_reads, self.alt_reads = ".", "./.", None, None

    def get_source(self):
        return self.locus

    def get_destination(self):
        return self.locus

    def result(self):
        return [self.support_fraction, self.genotype, self.ref_reads, self.alt_reads]


def options_of(**kw):
    d = dict(min_mapq=20, minimum_score=3, minimum_depth=4, homozygous_threshold=0.8, heterozygous_threshold=0.2)
    d.update(kw)
    return types.SimpleNamespace(**d)


def fresh_context():
    from svim_b200 import runtime
    ctx = runtime.context()
    ctx.resident = None; ctx.collect_batch = None
    return ctx


@pytest.mark.parametrize("name", GOLDEN_GENOTYPE)
def test_genotype_matches_reference_golden(golden, name):
    from svim_b200.SVIM_genotyping import genotype
    batch, genome, exp = golden(name)
    fresh_context()
    opts = options_of(**exp["params"])
    for t in GENO_TYPES:
        cands = [Candidate(*row[0]) for row in exp["genotype"][t]]
        genotype(cands, batch, t, opts)
        got = [c.result() for c in cands]
        want = [row[1] for row in exp["genotype"][t]]
        assert got == want, (t, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w][:3])


def test_genotype_after_collect_reuses_resident_rows(golden):
    # the rows COLLECT uploaded stay in HBM: genotype() on the same batch must not need a second upload
    from svim_b200.SVIM_COLLECT import analyze_alignment_file_coordsorted
    from svim_b200.SVIM_genotyping import genotype
    from svim_b200 import runtime
    batch, genome, exp = golden("geno_mini_indel")
    ctx = fresh_context()
    analyze_alignment_file_coordsorted(batch, options_of(min_sv_size=40, max_sv_size=100000, segment_gap_tolerance=10,
                                                         segment_overlap_tolerance=5, all_bnds=False))
    assert ctx.resident is batch
    uploads = []
    orig = ctx.upload
    ctx.upload = lambda b: (uploads.append(b), orig(b))
    try:
        for t in GENO_TYPES:
            cands = [Candidate(*row[0]) for row in exp["genotype"][t]]
            genotype(cands, batch, t, options_of())
            assert [c.result() for c in cands] == [row[1] for row in exp["genotype"][t]], t
    finally:
        del ctx.upload
    assert uploads == []


def synthetic_candidates(batch, n, seed, ins_like):
    """Random loci + the names of some reads that really overlap them (so the variant-read skip is exercised)."""
    rng = np.random.default_rng(seed)
    ends = orc.record_reference_ends(batch)
    out = []
    for _ in range(n):
        tid = int(rng.integers(0, len(batch.contig_names)))
        L = int(batch.contig_lengths[tid])
        start = int(rng.integers(0, max(1, L - 10)))
        size = int(np.exp(rng.uniform(np.log(40), np.log(20000))))
        end = start if ins_like and rng.random() < 0.5 else start + size
        idx = np.nonzero((batch.tid == tid) & (batch.pos < start + 50) & (ends > start - 50))[0]
        k = int(rng.integers(0, 12))
        reads = [batch.qname(int(batch.qname_id[i])) for i in rng.choice(idx, size=min(k, len(idx)), replace=False)] if len(idx) else []
        if rng.random() < 0.1:
            reads.append("no_such_read")                   # a member name without a record still counts in alt_reads
        out.append((batch.contig_names[tid], start, end, float(rng.integers(0, 30)), reads))
    return out


@pytest.mark.parametrize("config,scale", [("config1", 1.0), ("config5", 0.01)])
def test_genotype_matches_oracle_on_synthetic(config, scale):
    from svim_b200 import synth
    from svim_b200.SVIM_genotyping import genotype
    batch, genome, _ = synth.make_config(config, scale)
    fresh_context()
    ends = orc.record_reference_ends(batch)
    for t in GENO_TYPES:
        rows = synthetic_candidates(batch, 300, 7 + len(t), t in ("INS", "DUP_INT"))
        for kw in ({}, {"min_mapq": 1, "minimum_depth": 9, "homozygous_threshold": 0.6, "heterozygous_threshold": 0.4, "minimum_score": 10}):
            mine = [Candidate(*r) for r in rows]
            want = [orc.GenoCand(*r) for r in rows]
            genotype(mine, batch, t, options_of(**kw))
            orc.genotype(want, batch, t, orc.GenoParams(**kw), ends)
            assert [c.result() for c in mine] == [c.result() for c in want], (config, t, kw)


def test_genotype_edge_cases():
    from svim_b200.records import BatchBuilder
    from svim_b200.SVIM_genotyping import genotype
    b = BatchBuilder(["c", "empty", "d"], [100000, 5000, 300])
    b.add("long", 0, 0, 100, 60, "30000M", None)          # overlaps every window although it starts far to the left
    for k in range(700):
        b.add("r%d" % (k % 650), 0 if k % 7 else 0x800, 0, 9000 + k, 60 if k % 11 else 3, "1000M500N1500M", None)   # repeated names, low MAPQ
    b.add("v", 0, 0, 9990, 60, "5M100D2995M", None)
    b.add("sec", 0x100, 0, 9991, 60, "3000M", None)
    b.add("unm", 0x4, 0, 9992, 0, "", None)
    b.add("short", 0, 2, 10, 60, "50M", None)
    b.add("unplaced", 0x4, -1, -1, 0, "", None)
    batch = b.finish()
    fresh_context()
    ends = orc.record_reference_ends(batch)
    rows = [("c", 9995, 10095, 10, ["v"]), ("c", 9995, 10095, 10, []), ("c", 0, 40000, 5, ["long"]), ("c", 30099, 30150, 5, []),
            ("c", 30101, 30150, 5, ["x"]), ("empty", 100, 200, 9, ["v"]), ("d", 20, 40, 9, []), ("c", 99990, 100000, 3, ["r1", "r1", "r2"]),
            ("c", 9995, 10095, 2, ["v"])]
    for t in GENO_TYPES:
        mine = [Candidate(*r) for r in rows]
        want = [orc.GenoCand(*r) for r in rows]
        genotype(mine, batch, t, options_of())
        orc.genotype(want, batch, t, orc.GenoParams(), ends)
        assert [c.result() for c in mine] == [c.result() for c in want], t
        assert mine[-1].result() == [".", "./.", None, None]          # below minimum_score: untouched (:39-40)
    genotype([], batch, "DEL", options_of())
    # what the reference raises
    with pytest.raises(ValueError):                                    # fetch: start > stop (locus beyond the contig end)
        genotype([Candidate("d", 5000, 5100, 9, [])], batch, "DEL", options_of())
    with pytest.raises(KeyError):
        genotype([Candidate("nope", 1, 2, 9, [])], batch, "DEL", options_of())
    with pytest.raises(ValueError):
        genotype([Candidate("c", 1, 2, 9, [])], batch.take(np.arange(batch.n)[::-1], "unknown"), "DEL", options_of())
    # a mapped record without CIGAR has reference_end None: the comparison raises TypeError in the reference
    b2 = BatchBuilder(["c"], [100000])
    b2.add("nocigar", 0, 0, 100, 60, "", None)
    b2.add("ok", 0, 0, 120, 60, "3000M", None)
    bad = b2.finish()
    with pytest.raises(TypeError):
        genotype([Candidate("c", 500, 600, 9, [])], bad, "DEL", options_of())
    # records out of coordinate order although the header claims otherwise: the library refuses (fetch needs an index)
    from svim_b200 import _lib
    with pytest.raises(_lib.SvimGpuError):
        genotype([Candidate("c", 1, 2, 9, [])], batch.take(np.arange(batch.n)[::-1], "coordinate"), "DEL", options_of())


# ---- cut&paste search of COMBINE (SVIM_merging.py:12-29), svimgpu_closest_source ------------------------------------------
class UniCluster:
    def __init__(self, contig, start, end):
        self.src = (contig, start, end)

    def get_source(self):
        return self.src


class BiCluster(UniCluster):
    def __init__(self, contig, start, end, dcontig, dstart, k):
        super().__init__(contig, start, end)
        self.dst = (dcontig, dstart, dstart + (end - start))
        self.members, self.score, self.std_span, self.std_pos = ["m%d" % k], 4.0, None, None

    def get_destination(self):
        return self.dst


def test_flag_cutpaste_candidates_matches_reference_golden():
    import gzip, json, os
    from conftest import GOLDEN
    from svim_b200.SVIM_merging import flag_cutpaste_candidates, closest_deletion
    g = json.load(gzip.open(os.path.join(GOLDEN, "cutpaste.golden.json.gz"), "rt"))
    opts = types.SimpleNamespace(position_distance_normalizer=g["position_distance_normalizer"], del_ins_dup_max_distance=g["del_ins_dup_max_distance"])
    for case in g["cases"]:
        dels = [UniCluster(*d) for d in case["dels"]]
        inss = [BiCluster(*row, k) for k, row in enumerate(case["inss"])]
        got = flag_cutpaste_candidates(inss, dels, opts)
        assert [bool(c.cutpaste) for c in got] == case["cutpaste"]
        assert [(c.source_contig, c.source_start, c.source_end, c.dest_contig, c.dest_start, c.dest_end, c.members) for c in got] == \
               [(r[0], max(0, r[1]), r[2], r[3], max(0, r[4]), r[4] + r[2] - r[1], ["m%d" % k]) for k, r in enumerate(case["inss"])]
        idx, dist = closest_deletion(inss, dels, opts)
        assert [[int(i), float(d)] for i, d in zip(idx, dist)] == case["closest"]        # bit-equal FP64, first minimum
    assert flag_cutpaste_candidates([], [], opts) == []
    with pytest.raises(IndexError):
        flag_cutpaste_candidates([BiCluster("c", 1, 5, "c", 9, 0)], [], opts)
    with pytest.raises(ZeroDivisionError):
        flag_cutpaste_candidates([BiCluster("c", 5, 5, "c", 9, 0)], [UniCluster("c", 1, 9), UniCluster("c", 7, 7)], opts)


def test_closest_source_matches_oracle_at_scale():
    from svim_b200 import runtime
    rng = np.random.default_rng(11)
    n_b, n_a = 20000, 3000
    b_s = rng.integers(0, 250_000_000, n_b); b_e = b_s + rng.integers(40, 20000, n_b)
    pick = rng.integers(0, n_b, n_a)
    a_s = np.where(rng.random(n_a) < 0.5, b_s[pick] + rng.integers(-50, 50, n_a), rng.integers(0, 250_000_000, n_a))
    a_e = a_s + np.where(rng.random(n_a) < 0.5, b_e[pick] - b_s[pick], rng.integers(40, 20000, n_a))
    idx, dist = runtime.context().closest_source(a_s, a_e, b_s, b_e, 900)
    # numpy restatement of the same FP64 expression (SVIM_clustering.py:99-107), first minimum
    for k in range(0, n_a, 7):
        span1 = (b_e - b_s); span2 = int(a_e[k] - a_s[k])
        d = np.abs((b_s + b_e) // 2 - (int(a_s[k]) + int(a_e[k])) // 2) / 900 + np.abs(span1 - span2) / np.maximum(span1, span2)
        j = int(np.argmin(d))
        assert int(idx[k]) == j and float(dist[k]) == float(d[j]), k
    want = orc.flag_cutpaste(list(zip(a_s[:40].tolist(), a_e[:40].tolist())), list(zip(b_s.tolist(), b_e.tolist())))
    assert [(int(i), float(x)) for i, x in zip(idx[:40], dist[:40])] == [(m[0], m[1]) for m in want]
