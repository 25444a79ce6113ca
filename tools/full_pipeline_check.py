#!/usr/bin/env python3
"""Drop-in check of the host mirror under the reference's own CLI (build container only, no GPU):

  run A  the UNMODIFIED `svim alignment` script (reference tree, pysam/edlib/spoa shims) on a BAM + FASTA written to disk;
  run B  the same script after `svim_b200.patch.install()` — COLLECT, CLUSTER, the cut&paste search, the candidate clustering
         and GENOTYPE rebound to svim_b200's host mirror — with the CUDA context replaced by the oracle-backed stand-in of tests/test_host_units.py
         (the CUDA entries themselves are compared with the same goldens in the -m gpu tests).

Everything downstream of the rebound functions (COMBINE, candidate clustering, VCF / BED writers) is the reference's code and
consumes svim_b200's objects in run B; the working directories must come out byte-identical.
"""
import filecmp
import os
import runpy
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

refenv.activate()
sys.path.insert(0, os.path.join(refenv.ROOT, "tests"))

import numpy as np  # noqa: E402

SCRIPT = os.path.join(refenv.REF, "svim", "svim")
CALLS = {}          # how often run B went through each rebound entry (the check is void if they are not reached)


def run_cli(argv):
    import logging
    old = sys.argv
    sys.argv = [SCRIPT] + argv
    sys.modules.pop("svim.SVIM_input_parsing", None)      # its parse_arguments() binds sys.argv[1:] as a default at import time
    try:
        runpy.run_path(SCRIPT, run_name="__main__")
    except SystemExit as e:
        if e.code not in (None, 0):
            raise
    finally:
        sys.argv = old
        root = logging.getLogger()
        for h in list(root.handlers):
            root.removeHandler(h); h.close()


def compare_dirs(a, b):
    """Every file except the time-stamped log must be byte-identical (VCFs: except their ##fileDate line)."""
    diffs, n = [], 0
    for base, _dirs, files in os.walk(a):
        for f in files:
            if f.startswith("SVIM_") and f.endswith(".log"):
                continue
            pa = os.path.join(base, f); pb = os.path.join(b, os.path.relpath(pa, a))
            n += 1
            if not os.path.exists(pb):
                diffs.append(os.path.relpath(pa, a))
            elif f.endswith(".vcf"):          # ##fileDate carries the wall-clock time of the run
                strip = lambda p: [l for l in open(p) if not l.startswith("##fileDate")]
                if strip(pa) != strip(pb):
                    diffs.append(os.path.relpath(pa, a))
            elif not filecmp.cmp(pa, pb, shallow=False):
                diffs.append(os.path.relpath(pa, a))
    for base, _dirs, files in os.walk(b):
        for f in files:
            pb = os.path.join(base, f)
            if not os.path.exists(os.path.join(a, os.path.relpath(pb, b))) and not f.endswith(".log"):
                diffs.append("only in B: " + os.path.relpath(pb, b))
    return n, diffs


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--fuzz", type=int, default=0, help="additionally run this many random inputs / option sets of tools/fuzz_oracle.py")
    args = ap.parse_args()
    from conftest import load_golden
    from svim_b200 import io as sio, runtime
    import test_host_units as thu
    results = []
    cases = [(name, extra, None) for name, extra in (("mini_mixed", []), ("mini_indel", ["--minimum_depth", "2"]), ("mini_mixed", ["--all_bnds", "--min_mapq", "1", "--max_sv_size", "3000"]),
                        ("mini_ins", ["--minimum_score", "1"]), ("mini_hotspot", ["--cluster_max_distance", "0.3"]), ("geno_deep", []))]
    if args.fuzz:
        import fuzz_oracle
        for seed in range(args.fuzz):
            batch, genome, opts, gopts = fuzz_oracle.random_case(1000 + seed)
            extra = []
            for k, v in list(opts.items()) + list(gopts.items()):
                extra += ["--" + k] if v is True else ["--" + k, str(v)]
            cases.append(("fuzz%d" % (1000 + seed), extra, (batch, genome)))
    for name, extra, data in cases:
        batch, genome = data if data is not None else load_golden(name)[:2]
        with tempfile.TemporaryDirectory() as td:
            bam = os.path.join(td, "in.bam"); fa = os.path.join(td, "genome.fa")
            sio.write_bam(bam, batch)
            genome.write_fasta(fa)
            argv = lambda wd: ["alignment", wd, bam, fa, "--skip_consensus"] + extra      # spoa is absent here
            wa, wb = os.path.join(td, "A"), os.path.join(td, "B")
            run_cli(argv(wa))
            # ---- run B: rebind, with the oracle-backed context ----
            import svim_b200.patch as patch
            import svim.SVIM_COLLECT as rc, svim.SVIM_CLUSTER as rcl, svim.SVIM_genotyping as rg, svim.SVIM_merging as rm, svim.SVIM_clustering as rcg
            saved = (rc.analyze_alignment_file_coordsorted, rcl.cluster_sv_signatures, rg.genotype, rm.flag_cutpaste_candidates, rcg.partition_and_cluster_candidates)
            decoded = sio.read_alignments(bam)
            fake = thu._oracle_backed_context(decoded, sio.Genome.from_fasta(fa))
            add_genotype_and_cutpaste(fake, decoded)
            real_context = runtime.context
            runtime.context = lambda device=None: fake
            try:
                patch.install()
                run_cli(argv(wb))
            finally:
                runtime.context = real_context
                rc.analyze_alignment_file_coordsorted, rcl.cluster_sv_signatures, rg.genotype, rm.flag_cutpaste_candidates, rcg.partition_and_cluster_candidates = saved
                cm = sys.modules.get("svim.SVIM_COMBINE")
                if cm is not None:
                    cm.flag_cutpaste_candidates = saved[3]; cm.partition_and_cluster_candidates = saved[4]
            n, diffs = compare_dirs(wa, wb)
            if diffs and os.environ.get("SVIM_CHECK_KEEP"):
                import shutil
                shutil.copytree(wa, os.path.join(os.environ["SVIM_CHECK_KEEP"], name + "_A"), dirs_exist_ok=True)
                shutil.copytree(wb, os.path.join(os.environ["SVIM_CHECK_KEEP"], name + "_B"), dirs_exist_ok=True)
            vp = os.path.join(wa, "variants.vcf")
            vcf = open(vp).read().count("\n") if os.path.exists(vp) else -1      # -1: the reference itself stopped before writing it
            results.append((name, extra, n, vcf, diffs))
            print("%s %s: %d output files compared, variants.vcf %d lines -> %s" % (name, " ".join(extra), n, vcf, "IDENTICAL" if not diffs else "DIFFER: %s" % diffs), flush=True)
    print("rebound entries reached in the B runs:", CALLS)
    return 1 if any(r[4] for r in results) or not all(CALLS.get(k) for k in ("genotype", "closest_source", "candidate_clustering")) else 0


def add_genotype_and_cutpaste(fake, batch):
    """GENOTYPE and closest-deletion entries of the stand-in context (same as in tests/test_host_units.py)."""
    from svim_b200 import _lib
    from oracle import svim_oracle as orc
    ends = orc.record_reference_ends(batch)

    def genotype(type_code, gp, cands, variant_ids, contig_lengths):
        CALLS["genotype"] = CALLS.get("genotype", 0) + 1
        t = _lib.TYPE_NAMES[type_code]
        res = np.zeros(len(cands), dtype=_lib.GENO_RESULT_DTYPE)
        p = orc.GenoParams(min_mapq=gp.min_mapq, minimum_score=-10**9, minimum_depth=gp.minimum_depth,
                           homozygous_threshold=gp.homozygous_threshold, heterozygous_threshold=gp.heterozygous_threshold)
        for k, c in enumerate(cands):
            ids = variant_ids[int(c["variant_off"]):int(c["variant_off"]) + int(c["n_variant_reads"])]
            oc = orc.GenoCand(batch.contig_names[int(c["tid"])], int(c["start"]), int(c["end"]), 0,
                              [batch.qname(int(q)) if q != 0xFFFFFFFF else "\x00none%d" % i for i, q in enumerate(ids)])
            orc.genotype([oc], batch, t, p, ends)
            res[k]["support_fraction"] = float("nan") if oc.support_fraction == "." else oc.support_fraction
            res[k]["genotype"] = _lib.GENOTYPES.index(oc.genotype); res[k]["ref_reads"] = oc.ref_reads; res[k]["alt_reads"] = oc.alt_reads
        return res

    def closest_source(a_s, a_e, b_s, b_e, N):
        CALLS["closest_source"] = CALLS.get("closest_source", 0) + 1
        got = orc.flag_cutpaste(list(zip(a_s, a_e)), list(zip(b_s, b_e)), N, 0.0)
        return np.array([g[0] for g in got], dtype=np.int64), np.array([g[1] for g in got], dtype=np.float64)

    # candidate clustering twin (svimgpu_set_signatures with SVIM_DUP_INT_CAND records + svimgpu_cluster), answered by the oracle
    state = {"cand": None}
    collected_cluster = fake.cluster

    def set_signatures(cs, blob=None, rank_to_tid=None):
        state["cand"] = np.array(cs, copy=True)

    def cluster(sharded=False, view=False):
        cs = state["cand"]
        if cs is None:
            return collected_cluster(sharded)
        state["cand"] = None
        CALLS["candidate_clustering"] = CALLS.get("candidate_clustering", 0) + 1
        dest_end = cs["seq_off"].view(np.float64)
        cands = [orc.Cand(int(r["contig_a"]), int(r["start"]), int(r["end"]), int(r["contig_b"]), int(r["dpos"]), int(dest_end[k]), [k], 1.0, None, None)
                 for k, r in enumerate(cs)]
        merged = orc.partition_and_cluster_candidates(cands, fake.p)
        rows, mem = [], []
        for c in merged:
            rows.append((c.start, c.end, c.dest_start, c.dest_end, 0.0, float("nan"), float("nan"), len(mem), len(c.members), 5, 0, 0, 0, 0))
            mem += c.members
        st = _lib.ClusterStats()
        st.n_clusters_total, st.n_members = len(rows), len(mem)
        return st, np.array(rows, dtype=_lib.CLUSTER_DTYPE) if rows else np.zeros(0, _lib.CLUSTER_DTYPE), np.array(mem, dtype=np.uint32)

    use_collected = fake.use_collected

    def use(which=0):
        state["cand"] = None
        use_collected(which)

    fake.set_signatures = set_signatures
    fake.cluster = cluster
    fake.use_collected = use
    fake.genotype = genotype
    fake.closest_source = closest_source
    fake.upload = lambda b: None


if __name__ == "__main__":
    sys.exit(main())
