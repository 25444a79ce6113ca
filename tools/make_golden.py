#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference (imported from
/root/reference through tools/ref_shim) on seeded synthetic inputs, and check that
oracle/svim_oracle.py reproduces every output exactly.

Run in the build container only:   python tools/make_golden.py
The committed fixtures (inputs as .npz, reference outputs as .json.gz) are what
tests/ and the GPU box use; /root/reference is never read at test time.
"""
import gzip
import json
import os
import sys
import tempfile
import argparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

refenv.activate()

import numpy as np  # noqa: E402
import pysam  # the shim  # noqa: E402
from svim.SVIM_COLLECT import analyze_alignment_file_coordsorted, analyze_alignment_file_querysorted  # noqa: E402
from svim.SVIM_CLUSTER import cluster_sv_signatures  # noqa: E402
from svim.SVIM_input_parsing import parse_arguments  # noqa: E402

from svim_b200 import synth  # noqa: E402
from svim_b200.records import AlignmentBatch  # noqa: E402
from svim_b200.io import Genome  # noqa: E402
from oracle import svim_oracle as orc  # noqa: E402

GOLDEN = os.path.join(refenv.ROOT, "tests", "golden")
SIG_FIELDS = orc.Sig.__slots__


def ref_sig_tuple(s):
    t = s.type
    d = dict.fromkeys(SIG_FIELDS)
    d.update(type=t, signature=s.signature, read=s.read)
    if t in ("DEL", "INS", "INV", "DUP_TAN"):
        d.update(contig=s.contig, start=s.start, end=s.end)
        if t == "INS":
            d["sequence"] = s.sequence
        if t == "INV":
            d["direction"] = s.direction
        if t == "DUP_TAN":
            d.update(copies=s.copies, fully_covered=s.fully_covered)
    elif t == "DUP_INT":
        d.update(contig=s.contig1, start=s.start, end=s.end, contig2=s.contig2, pos=s.pos)
    elif t == "BND":
        d.update(contig=s.contig1, start=s.pos1, end=s.pos1 + 1, contig2=s.contig2, pos=s.pos2,
                 dir1=s.direction1, dir2=s.direction2)
    return [d[f] for f in SIG_FIELDS]


def ref_cluster_tuple(c, index_of):
    members = [index_of[id(m)] for m in c.members]
    if hasattr(c, "source_contig"):
        return [c.type, c.source_contig, c.source_start, c.source_end, c.dest_contig, c.dest_start, c.dest_end,
                c.score, c.size, c.std_span, c.std_pos, getattr(c, "direction1", None), getattr(c, "direction2", None), members]
    return [c.type, c.contig, c.start, c.end, None, None, None, c.score, c.size, c.std_span, c.std_pos, None, None, members]


def oracle_cluster_tuple(c, index_of):
    return [c.type, c.contig, c.start, c.end, c.dest_contig, c.dest_start, c.dest_end, c.score, c.size,
            c.std_span, c.std_pos, c.dir1, c.dir2, [index_of[id(m)] for m in c.members]]


def run_reference(batch, genome, overrides, querysorted=False):
    with tempfile.TemporaryDirectory() as td:
        gpath = os.path.join(td, "genome.fa")
        pysam.register_genome(gpath, genome)
        argv = ["alignment", td, "in.bam", gpath]
        for k, v in overrides.items():
            if v is True:
                argv.append("--" + k)
            else:
                argv += ["--" + k, str(v)]
        options = parse_arguments("2.0.0", argv)
        bam = pysam.AlignmentFile.from_batch(batch)
        sigs, twins = (analyze_alignment_file_querysorted if querysorted else analyze_alignment_file_coordsorted)(bam, options)
        out = {"signatures": [ref_sig_tuple(s) for s in sigs], "all_bnds_signatures": [ref_sig_tuple(s) for s in twins]}
        for key, lst in (("clusters", sigs), ("all_bnds_clusters", twins)):
            index_of = {id(s): i for i, s in enumerate(lst)}
            res = cluster_sv_signatures(lst, options)
            out[key] = {name: [ref_cluster_tuple(c, index_of) for c in cl]
                        for name, cl in zip(("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND"), res)}
        return out


def run_oracle(batch, genome, overrides, querysorted=False):
    p = orc.Params(**overrides)
    sigs, twins = (orc.collect_querysorted if querysorted else orc.collect)(batch, p)
    out = {"signatures": [list(s.as_tuple()) for s in sigs], "all_bnds_signatures": [list(s.as_tuple()) for s in twins]}
    for key, lst in (("clusters", sigs), ("all_bnds_clusters", twins)):
        index_of = {id(s): i for i, s in enumerate(lst)}
        res = orc.cluster(lst, genome, p)
        out[key] = {name: [oracle_cluster_tuple(c, index_of) for c in cl]
                    for name, cl in zip(("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND"), res)}
    return out


def save_input(path, batch, genome):
    arrays = {name: getattr(batch, name) for name, _ in AlignmentBatch.FIELDS}
    extra = {}
    if batch.qnames is not None:
        extra["qnames"] = np.array(batch.qnames)
    np.savez_compressed(path, contig_names=np.array(batch.contig_names), contig_lengths=batch.contig_lengths if len(genome.blob) else np.zeros(len(batch.contig_names), np.int64),
                        cigar=batch.cigar, seq=batch.seq, sa=batch.sa, genome_blob=genome.blob, **arrays, **extra)


def fixtures():
    """name -> (batch, genome, option overrides)."""
    out = {}
    # 1. DEL/INS only, one contig (BASELINE config 1 in miniature)
    names, L = ["chr1"], [200_000]
    svs, al = synth.plant_svs(L, 11, spacing=8000, mix={"DEL": 0.5, "INS": 0.5}, size_range=(50, 1500))
    out["mini_indel"] = (synth.generate(names, L, 400, 11, svs, al, len_mean=5000, len_sd=800, len_min=1000, len_max=9000),
                         synth.random_genome(names, L, 11), {})
    # 2. every SV class, contig names chosen so string order != numeric order
    names, L = ["chr1", "chr10", "chr2"], [150_000, 120_000, 130_000]
    svs, al = synth.plant_svs(L, 12, spacing=5000, size_range=(50, 1200),
                              mix={"DEL": 0.2, "INS": 0.2, "INV": 0.15, "DUP_TAN": 0.15, "BND": 0.15, "DUP_INT": 0.15})
    b2 = synth.generate(names, L, 700, 12, svs, al, len_mean=6000, len_sd=1500, len_min=1500, len_max=12000, p_ins=0.03, p_del=0.02)
    g2 = synth.random_genome(names, L, 12)
    out["mini_mixed"] = (b2, g2, {})
    out["mini_mixed_allbnds"] = (b2, g2, {"all_bnds": True, "min_mapq": 1, "max_sv_size": 3000})
    # 2b. the same records sorted by read name (query-sorted mode, SVIM_COLLECT.py:96-129): primary not always first
    sys.path.insert(0, os.path.join(refenv.ROOT, "tests"))
    from conftest import querysort_order
    out["mini_mixed_querysorted"] = (b2.take(querysort_order(b2), "queryname"), g2, {"all_bnds": True})
    # 3. insertion heavy (haplotype edit distance)
    names, L = ["chr1"], [120_000]
    svs, al = synth.plant_svs(L, 13, spacing=4000, mix={"INS": 1.0}, ins_size_uniform=(60, 700))
    out["mini_ins"] = (synth.generate(names, L, 500, 13, svs, al, len_mean=4000, len_sd=800, len_min=1000, len_max=8000, p_ins=0.03, p_del=0.02),
                       synth.random_genome(names, L, 13), {})
    # 4. partitions above 100 signatures (host RNG sampling) + dedup
    names, L = ["chr1"], [60_000]
    svs, al = synth.plant_svs(L, 14, spacing=6000, hotspots=2, hotspot_svs=(8, 14), size_range=(50, 400))
    out["mini_hotspot"] = (synth.generate(names, L, 2600, 14, svs, al, len_mean=3000, len_sd=500, len_min=1000, len_max=6000, p_ins=0.02, p_del=0.01),
                           synth.random_genome(names, L, 14), {})
    # 5. known-answer vector derived from the reference's own test data (tests/chimeric_read.sam, used by
    #    tests/test_satag.py): one 9.9 kb read, primary + 3 supplementary records with SA tags.
    from svim_b200.io import read_sam, Genome
    kat = read_sam(os.path.join(refenv.REF, "tests", "chimeric_read.sam"))
    kat.sort_order = "coordinate"
    out["chimeric_kat"] = (kat, Genome(kat.contig_names, [np.zeros(0, np.uint8) for _ in kat.contig_names]), {})
    return out


def candidate_fixture():
    """DUP_INT candidates through the reference's partition_and_cluster_candidates (SVIM_clustering.py:306-372),
    the COMBINE-stage twin of the clustering pipeline; also checks the oracle's restatement."""
    import random
    from svim.SVIM_clustering import partition_and_cluster_candidates as ref_pcc
    from svim.SVCandidate import CandidateDuplicationInterspersed as RefCand
    rng = random.Random(42)
    rows = []
    for locus in range(60):
        contig = rng.choice(["chr1", "chr10", "chr2"]); base = rng.randint(10_000, 900_000); dbase = rng.randint(10_000, 900_000)
        n = rng.choice([1, 1, 2, 3, 5, 8]) if locus else 140          # one partition above 100 -> sampling
        for k in range(n):
            s = base + rng.randint(-300, 300); ln = rng.randint(200, 900) + rng.choice([0, 0, 2000])
            d = dbase + rng.randint(-300, 300) + rng.choice([0, 0, 5000])
            rows.append([contig, s, s + ln, rng.choice(["chr1", "chr2"]), d, d + ln + rng.randint(-5, 5), ["m%d_%d" % (locus, k)],
                         rng.randint(1, 40) + rng.random(), rng.choice([None, rng.random() * 30]), rng.choice([None, rng.random() * 30]),
                         rng.random() < 0.2])
    rng.shuffle(rows)
    options = parse_arguments("2.0.0", ["alignment", "wd", "x.bam", "g.fa"])
    ref = ref_pcc([RefCand(*r[:10], cutpaste=r[10]) for r in rows], options, "interspersed duplication candidates")
    mine = orc.partition_and_cluster_candidates([orc.Cand(*r) for r in rows], orc.Params())
    ref_rows = [[c.source_contig, c.source_start, c.source_end, c.dest_contig, c.dest_start, c.dest_end, c.members, c.score, c.std_span,
                 c.std_pos, c.cutpaste] for c in ref]
    mine_rows = [[c.contig, c.start, c.end, c.dest_contig, c.dest_start, c.dest_end, c.members, c.score, c.std_span, c.std_pos, c.cutpaste]
                 for c in mine]
    if ref_rows != mine_rows:
        raise SystemExit("ORACLE != REFERENCE on the candidate clustering twin")
    with gzip.open(os.path.join(GOLDEN, "candidates.golden.json.gz"), "wt") as fh:
        json.dump({"input": rows, "output": ref_rows}, fh)
    print("candidates", len(rows), "->", len(ref_rows))


GENO_TYPES = ("DEL", "INV", "INS", "DUP_INT")      # call order svim:161-170


def genotype_fixtures():
    """name -> (input fixture name or (batch, genome), option overrides).  GENOTYPE (SVIM_genotyping.py:34-93) through the
    UNMODIFIED reference: COLLECT -> CLUSTER -> COMBINE (--skip_consensus: spoa is absent here) -> genotype() on the shim's
    region fetch; the candidates (locus, score, member read names) become the fixture input, the four attributes
    genotype() writes the expected output."""
    fx = fixtures()
    out = {"geno_mini_indel": (fx["mini_indel"][:2], "mini_indel", {}),
           "geno_mini_mixed": (fx["mini_mixed"][:2], "mini_mixed", {"min_mapq": 1, "minimum_score": 1, "minimum_depth": 6,
                                                                    "homozygous_threshold": 0.7, "heterozygous_threshold": 0.3}),
           "geno_mini_hotspot": (fx["mini_hotspot"][:2], "mini_hotspot", {"minimum_score": 0})}
    # deep coverage: more than 500 countable alignments per window (the aln_no cap, :57)
    names, L = ["chrD"], [24_000]
    svs, al = synth.plant_svs(L, 15, spacing=5000, mix={"DEL": 0.5, "INS": 0.5}, size_range=(50, 900))
    deep = synth.generate(names, L, 4200, 15, svs, al, len_mean=3000, len_sd=600, len_min=800, len_max=6000, p_ins=0.02, p_del=0.01)
    out["geno_deep"] = ((deep, synth.random_genome(names, L, 15)), None, {})
    return out


def run_reference_genotype(batch, genome, overrides):
    from svim.SVIM_COMBINE import combine_clusters
    from svim.SVIM_genotyping import genotype
    with tempfile.TemporaryDirectory() as td:
        gpath = os.path.join(td, "genome.fa")
        pysam.register_genome(gpath, genome)
        argv = ["alignment", td, "in.bam", gpath, "--skip_consensus"]
        for k, v in overrides.items():
            argv += ["--" + k, str(v)]
        options = parse_arguments("2.0.0", argv)
        bam = pysam.AlignmentFile.from_batch(batch)
        sigs, _ = analyze_alignment_file_coordsorted(bam, options)
        clusters = cluster_sv_signatures(sigs, options)
        dels, invs, dup_ints, _tans, inss, _bnds = combine_clusters(clusters, options)
        res = {}
        for t, cands in (("DEL", dels), ("INV", invs), ("INS", inss), ("DUP_INT", dup_ints)):
            genotype(cands, bam, t, options)
            rows = []
            for c in cands:
                contig, start, end = c.get_destination() if t in ("INS", "DUP_INT") else c.get_source()
                rows.append([[contig, start, end, c.score, [m.read for m in c.members]],
                             [c.support_fraction, c.genotype, c.ref_reads, c.alt_reads]])
            res[t] = rows
        return res


def genotype_goldens(only=None):
    for name, (pair, inp, overrides) in genotype_fixtures().items():
        if only and name != only:
            continue
        batch, genome = pair
        ref = run_reference_genotype(batch, genome, overrides)
        gp = orc.GenoParams(**{k: v for k, v in overrides.items()})
        for t in GENO_TYPES:
            cands = [orc.GenoCand(*row[0]) for row in ref[t]]
            orc.genotype(cands, batch, t, gp)
            mine = [c.result() for c in cands]
            want = [row[1] for row in ref[t]]
            if mine != want:
                for i, (x, y) in enumerate(zip(want, mine)):
                    if x != y:
                        print("first diff", name, t, i, ref[t][i][0][:4], "\n ref", x, "\n orc", y); break
                raise SystemExit("ORACLE != REFERENCE on genotype %s/%s" % (name, t))
        if inp is None:
            inp = name
            save_input(os.path.join(GOLDEN, inp + ".input.npz"), batch, genome)
        with gzip.open(os.path.join(GOLDEN, name + ".golden.json.gz"), "wt") as fh:
            json.dump({"input": inp + ".input.npz", "params": overrides, "n_records": batch.n, "genotype": ref}, fh)
        from collections import Counter
        print(name, "records", batch.n, {t: (len(ref[t]), dict(Counter(r[1][1] for r in ref[t]))) for t in GENO_TYPES},
              "max ref_reads", max([r[1][2] or 0 for t in GENO_TYPES for r in ref[t]] + [0]))


def cutpaste_fixture():
    """flag_cutpaste_candidates (SVIM_merging.py:12-29) of the unmodified reference on DUP_INT / DEL clusters: the clusters the
    reference makes from mini_mixed plus seeded random ones (ties, identical intervals, far and near deletions)."""
    import random
    from svim.SVIM_merging import flag_cutpaste_candidates
    from svim.SVSignature import SignatureClusterUniLocal, SignatureClusterBiLocal
    rng = random.Random(7)
    cases = []
    for case in range(4):
        n_del = [1, 40, 700, 3][case]; n_ins = [5, 60, 150, 4][case]
        dels = []
        for _ in range(n_del):
            s = rng.randint(0, 200_000); ln = rng.choice([50, 51, 300, 1000, rng.randint(40, 5000)])
            dels.append([rng.choice(["chr1", "chr2"]), s, s + ln])
        inss = []
        for k in range(n_ins):
            if dels and rng.random() < 0.5:           # sit on / near a deletion (and on duplicates of it: ties)
                d = rng.choice(dels); s = d[1] + rng.choice([0, 0, 1, -3, 40, 900]); ln = d[2] - d[1] + rng.choice([0, 0, 2, -7, 100])
            else:
                s = rng.randint(0, 200_000); ln = rng.randint(40, 5000)
            inss.append([rng.choice(["chr1", "chr2"]), s, s + max(1, ln), rng.choice(["chr1", "chr2"]), rng.randint(0, 200_000)])
        if case == 1:
            dels += [list(d) for d in dels[:10]]      # exact duplicates -> equal distances, first index must win
        cases.append((dels, inss))
    options = parse_arguments("2.0.0", ["alignment", "wd", "x.bam", "g.fa"])
    out = []
    for dels, inss in cases:
        dcl = [SignatureClusterUniLocal(c, s, e, 5.0, 3, ["m"], "DEL", 1.0, 1.0) for c, s, e in dels]
        icl = [SignatureClusterBiLocal(c, s, e, dc, dp, dp + (e - s), 4.0, 2, ["m%d" % k], "DUP_INT", None, None) for k, (c, s, e, dc, dp) in enumerate(inss)]
        ref = flag_cutpaste_candidates(icl, dcl, options)
        flags = [bool(c.cutpaste) for c in ref]
        mine = orc.flag_cutpaste([(s, e) for _, s, e, _, _ in inss], [(s, e) for _, s, e in dels],
                                 options.position_distance_normalizer, options.del_ins_dup_max_distance)
        if flags != [m[2] for m in mine]:
            raise SystemExit("ORACLE != REFERENCE on flag_cutpaste_candidates")
        assert [(c.source_start, c.source_end, c.dest_start, c.dest_end) for c in ref] == [(max(0, s), e, max(0, dp), dp + (e - s)) for _, s, e, _, dp in inss]
        out.append({"dels": dels, "inss": inss, "cutpaste": flags, "closest": [[m[0], m[1]] for m in mine]})
    with gzip.open(os.path.join(GOLDEN, "cutpaste.golden.json.gz"), "wt") as fh:
        json.dump({"position_distance_normalizer": options.position_distance_normalizer,
                   "del_ins_dup_max_distance": options.del_ins_dup_max_distance, "cases": out}, fh)
    print("cutpaste", [(len(c["dels"]), len(c["inss"]), sum(c["cutpaste"])) for c in out])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    if not args.only or args.only == "candidates":
        candidate_fixture()
    if not args.only or args.only == "cutpaste":
        cutpaste_fixture()
        if args.only:
            return
    if not args.only or args.only.startswith("geno"):
        genotype_goldens(None if args.only in (None, "geno") else args.only)
        if args.only:
            return
    for name, (batch, genome, overrides) in fixtures().items():
        if args.only and name != args.only:
            continue
        qsort = name.endswith("_querysorted")
        ref = run_reference(batch, genome, overrides, qsort)
        mine = run_oracle(batch, genome, overrides, qsort)
        for key in ref:
            if ref[key] != mine[key]:
                a, b = ref[key], mine[key]
                if isinstance(a, dict):
                    for t in a:
                        if a[t] != b[t]:
                            for i, (x, y) in enumerate(zip(a[t], b[t])):
                                if x != y:
                                    print("first diff", name, key, t, i, "\n ref", x, "\n orc", y); break
                            print("len", len(a[t]), len(b[t]))
                else:
                    for i, (x, y) in enumerate(zip(a, b)):
                        if x != y:
                            print("first diff", name, key, i, "\n ref", x, "\n orc", y); break
                    print("len", len(a), len(b))
                raise SystemExit("ORACLE != REFERENCE on %s/%s" % (name, key))
        ref["params"] = overrides
        ref["n_records"] = batch.n
        inp = name.replace("_allbnds", "")
        if name.endswith("_querysorted"):       # same records as mini_mixed, permuted by tests/conftest.py::querysort_order
            inp = "mini_mixed"
            ref["derive"] = "querysorted"
        elif inp == name:
            save_input(os.path.join(GOLDEN, inp + ".input.npz"), batch, genome)
        ref["input"] = inp + ".input.npz"
        with gzip.open(os.path.join(GOLDEN, name + ".golden.json.gz"), "wt") as fh:
            json.dump(ref, fh)
        from collections import Counter
        print(name, "records", batch.n, "signatures", dict(Counter(s[0] for s in ref["signatures"])),
              "twins", len(ref["all_bnds_signatures"]),
              "clusters", {t: len(v) for t, v in ref["clusters"].items()})


if __name__ == "__main__":
    main()
