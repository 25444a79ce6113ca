#!/usr/bin/env python3
"""Pin the oracle beyond the committed fixtures: many small seeded inputs with random generator settings and random
option values, each run through the UNMODIFIED reference (COLLECT -> CLUSTER -> COMBINE -> GENOTYPE, imported from
/root/reference through tools/ref_shim) and through oracle/svim_oracle.py; every output must be equal.

    python tools/fuzz_oracle.py --cases 60 [--first 0]          # build container only

Writes one summary line per case and a final JSON line (kept under profiles/ as the evidence of the run).
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import make_golden as mg  # noqa: E402  (activates the reference environment)

import numpy as np  # noqa: E402
from svim_b200 import synth  # noqa: E402
from oracle import svim_oracle as orc  # noqa: E402


def random_case(seed):
    rng = np.random.default_rng(10_000 + seed)
    n_contigs = int(rng.integers(1, 4))
    names = [["chr1"], ["chr2", "chr10"], ["chrX", "chr1", "chr10"]][n_contigs - 1]      # string order != numeric order
    L = [int(rng.integers(60_000, 200_000)) for _ in names]
    kinds = ["DEL", "INS", "INV", "DUP_TAN", "BND", "DUP_INT"]
    w = rng.random(6) ** 2 + 0.02
    mix = {k: float(x) for k, x in zip(kinds, w)}
    spacing = int(rng.choice([2500, 4000, 8000]))
    hot = int(rng.integers(0, 2))
    svs, al = synth.plant_svs(L, 20_000 + seed, spacing=spacing, mix=mix, size_range=(int(rng.choice([30, 50])), int(rng.choice([400, 1500, 4000]))),
                              hotspots=hot, hotspot_svs=(6, 12))
    cov_reads = int(rng.integers(150, 900))
    len_mean = int(rng.choice([2500, 4000, 7000]))
    batch = synth.generate(names, L, cov_reads, 30_000 + seed, svs, al, len_mean=len_mean, len_sd=len_mean // 5, len_min=600, len_max=3 * len_mean,
                           p_ins=float(rng.choice([0.01, 0.03, 0.07])), p_del=float(rng.choice([0.01, 0.02, 0.04])),
                           p_lowmapq=float(rng.choice([0.02, 0.15])), p_secondary=0.02, p_unmapped=0.01, p_split=float(rng.choice([0.2, 0.5, 0.8])))
    genome = synth.random_genome(names, L, 40_000 + seed)
    opts = {}
    if rng.random() < 0.5: opts["min_sv_size"] = int(rng.choice([20, 40, 60, 100]))
    if rng.random() < 0.4: opts["max_sv_size"] = int(rng.choice([1000, 3000, 100000]))
    if rng.random() < 0.4: opts["min_mapq"] = int(rng.choice([0, 1, 20, 40]))
    if rng.random() < 0.3: opts["all_bnds"] = True
    if rng.random() < 0.4: opts["cluster_max_distance"] = float(rng.choice([0.2, 0.3, 0.5, 0.8]))
    if rng.random() < 0.4: opts["partition_max_distance"] = int(rng.choice([100, 1000, 5000]))
    if rng.random() < 0.3: opts["position_distance_normalizer"] = int(rng.choice([300, 900, 2000]))
    if rng.random() < 0.3: opts["edit_distance_normalizer"] = float(rng.choice([0.5, 1.0, 2.0]))
    if rng.random() < 0.3: opts["segment_gap_tolerance"] = int(rng.choice([5, 10, 30]))
    if rng.random() < 0.3: opts["segment_overlap_tolerance"] = int(rng.choice([2, 5, 20]))
    gopts = {}
    if rng.random() < 0.5: gopts["minimum_score"] = int(rng.choice([0, 1, 3, 8]))
    if rng.random() < 0.5: gopts["minimum_depth"] = int(rng.choice([1, 4, 10]))
    if rng.random() < 0.3: gopts.update(homozygous_threshold=0.7, heterozygous_threshold=0.3)
    return batch, genome, opts, gopts


def combine_twins(seed, tot):
    import random
    from svim.SVIM_clustering import partition_and_cluster_candidates as ref_pcc
    from svim.SVCandidate import CandidateDuplicationInterspersed as RefCand
    from svim.SVIM_merging import flag_cutpaste_candidates
    from svim.SVSignature import SignatureClusterUniLocal, SignatureClusterBiLocal
    from svim.SVIM_input_parsing import parse_arguments
    rng = random.Random(50_000 + seed)
    bad = []
    pmd = rng.choice([100, 1000, 5000]); cmd = rng.choice([0.2, 0.5, 0.8]); pdn = rng.choice([300, 900])
    options = parse_arguments("2.0.0", ["alignment", "wd", "x.bam", "g.fa", "--partition_max_distance", str(pmd), "--cluster_max_distance", str(cmd),
                                        "--position_distance_normalizer", str(pdn)])
    rows = []
    for locus in range(rng.randint(1, 25)):
        contig = rng.choice(["chr1", "chr10", "chr2"]); base = rng.randint(10_000, 300_000); dbase = rng.randint(10_000, 300_000)
        for k in range(rng.choice([1, 1, 2, 3, 6, 12, 130 if locus == 0 and seed % 5 == 0 else 4])):
            s0 = base + rng.randint(-400, 400); ln = rng.randint(100, 900) + rng.choice([0, 0, 1500]); d = dbase + rng.randint(-400, 400) + rng.choice([0, 0, 4000])
            rows.append([contig, s0, s0 + ln, rng.choice(["chr1", "chr2"]), d, d + ln + rng.randint(-5, 5), ["m%d_%d" % (locus, k)],
                         rng.randint(1, 40) + rng.random(), rng.choice([None, rng.random() * 30]), rng.choice([None, rng.random() * 30]), rng.random() < 0.2])
    rng.shuffle(rows)
    ref = ref_pcc([RefCand(*r[:10], cutpaste=r[10]) for r in rows], options, "interspersed duplication candidates")
    mine = orc.partition_and_cluster_candidates([orc.Cand(*r) for r in rows], orc.Params(partition_max_distance=pmd, cluster_max_distance=cmd, position_distance_normalizer=pdn))
    if [[c.source_contig, c.source_start, c.source_end, c.dest_contig, c.dest_start, c.dest_end, c.members, c.score, c.std_span, c.std_pos, c.cutpaste] for c in ref] != \
       [[c.contig, c.start, c.end, c.dest_contig, c.dest_start, c.dest_end, c.members, c.score, c.std_span, c.std_pos, c.cutpaste] for c in mine]:
        bad.append("candidates")
    tot["candidates"] = tot.get("candidates", 0) + len(rows)
    dels = [(rng.choice(["chr1", "chr2"]), s0, s0 + rng.choice([50, 300, rng.randint(40, 5000)])) for s0 in (rng.randint(0, 100_000) for _ in range(rng.randint(1, 200)))]
    inss = []
    for k in range(rng.randint(1, 60)):
        if rng.random() < 0.5:
            d = rng.choice(dels); s0 = d[1] + rng.choice([0, 1, -3, 40, 900]); ln = d[2] - d[1] + rng.choice([0, 0, 2, -7, 100])
        else:
            s0 = rng.randint(0, 100_000); ln = rng.randint(40, 5000)
        inss.append((rng.choice(["chr1", "chr2"]), s0, s0 + max(1, ln), "chr1", rng.randint(0, 100_000)))
    ddn = rng.choice([0.5, 1.0, 2.0]); options.del_ins_dup_max_distance = ddn
    refc = flag_cutpaste_candidates([SignatureClusterBiLocal(c, s0, e, dc, dp, dp + e - s0, 4.0, 2, ["m"], "DUP_INT", None, None) for c, s0, e, dc, dp in inss],
                                    [SignatureClusterUniLocal(c, s0, e, 5.0, 3, ["m"], "DEL", 1.0, 1.0) for c, s0, e in dels], options)
    minec = orc.flag_cutpaste([(s0, e) for _, s0, e, _, _ in inss], [(s0, e) for _, s0, e in dels], pdn, ddn)
    if [bool(c.cutpaste) for c in refc] != [m[2] for m in minec]:
        bad.append("cutpaste")
    tot["cutpaste_queries"] = tot.get("cutpaste_queries", 0) + len(inss)
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--first", type=int, default=0)
    args = ap.parse_args()
    t_start = time.time()
    tot = {"cases": 0, "records": 0, "signatures": 0, "twin_signatures": 0, "clusters": 0, "genotyped": 0, "mismatches": 0}
    for seed in range(args.first, args.first + args.cases):
        batch, genome, opts, gopts = random_case(seed)
        ref = mg.run_reference(batch, genome, opts)
        mine = mg.run_oracle(batch, genome, opts)
        bad = [k for k in ref if ref[k] != mine[k]]
        # GENOTYPE on the reference's own COMBINE candidates (all_bnds does not reach it)
        gparams = {k: v for k, v in opts.items() if k != "all_bnds"}
        gparams.update(gopts)
        try:
            gref = mg.run_reference_genotype(batch, genome, gparams)
        except IndexError:          # the reference's own COMBINE fails: DUP_INT clusters without any deletion cluster (SVIM_merging.py:20)
            gref = {t: [] for t in mg.GENO_TYPES}
            tot["reference_combine_raised"] = tot.get("reference_combine_raised", 0) + 1
        gp = orc.GenoParams(**{k: v for k, v in gparams.items() if k in ("min_mapq", "minimum_score", "minimum_depth", "homozygous_threshold", "heterozygous_threshold")})
        n_g = 0
        ends = orc.record_reference_ends(batch)
        for t in mg.GENO_TYPES:
            cands = [orc.GenoCand(*row[0]) for row in gref[t]]
            orc.genotype(cands, batch, t, gp, ends)
            if [c.result() for c in cands] != [row[1] for row in gref[t]]:
                bad.append("genotype/" + t)
            n_g += len(cands)
        # the same records sorted by read name through the query-sorted COLLECT (SVIM_COLLECT.py:96-129), every 3rd case
        if seed % 3 == 0:
            from conftest import querysort_order
            qb = batch.take(querysort_order(batch), "queryname")
            qref = mg.run_reference(qb, genome, opts, True)
            qmine = mg.run_oracle(qb, genome, opts, True)
            bad += ["querysorted/" + k for k in qref if qref[k] != qmine[k]]
            tot["querysorted_cases"] = tot.get("querysorted_cases", 0) + 1
        # COMBINE-stage twins on seeded random cluster-like inputs: candidate clustering and the cut&paste search
        bad += combine_twins(seed, tot)
        n_cl = sum(len(v) for v in ref["clusters"].values()) + sum(len(v) for v in ref["all_bnds_clusters"].values())
        tot["cases"] += 1; tot["records"] += batch.n; tot["signatures"] += len(ref["signatures"]); tot["twin_signatures"] += len(ref["all_bnds_signatures"])
        tot["clusters"] += n_cl; tot["genotyped"] += n_g; tot["mismatches"] += len(bad)
        from collections import Counter
        print("case %3d: %4d records, %4d signatures %s, %3d clusters, %3d genotyped, options %s %s -> %s" % (
            seed, batch.n, len(ref["signatures"]), dict(Counter(s[0] for s in ref["signatures"])), n_cl, n_g, opts, gopts, "MISMATCH " + ",".join(bad) if bad else "equal"), flush=True)
    tot["seconds"] = round(time.time() - t_start, 1)
    print(json.dumps(tot))
    return 1 if tot["mismatches"] else 0


if __name__ == "__main__":
    sys.exit(main())
