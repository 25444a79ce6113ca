#!/usr/bin/env python3
"""Full-size validation of the CUDA path on a BASELINE config (run on the GPU box):

    python tools/fullsize_check.py --workload config5 [--scale 1.0] [--spot 300]

Size-independent properties checked on the whole output, plus an oracle spot-check on a random sample of
partitions (the oracle is far too slow for the whole input):
  1. signatures come back in emission order (record index, ordinal) and COLLECT is deterministic (two runs, same bytes)
  2. every cluster's members share type and partition; clusters of a partition are disjoint; sizes add up
  3. cluster list order: unilocal types sorted by (contig, (start+end)/2)
  4. for `--spot` random partitions of <= 100 signatures per type: membership, order, coordinates and score equal the oracle's
  5. for up to `--spot-large` partitions ABOVE 100 signatures: the sampling stream is replayed with Python's own `random`
     (seed(1524) per type, one sample(range(size), 100) per large partition in partition order, SVIM_clustering.py:129-134), the
     oracle clusters exactly those 100 signatures in sample order, and the CUDA clusters of the partition must be the same
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--spot", type=int, default=300)
    ap.add_argument("--spot-large", type=int, default=60)
    args = ap.parse_args()
    from svim_b200 import _lib, synth, runtime
    from gpu_common import sig_rows
    from oracle import svim_oracle as orc
    t0 = time.time()
    batch, genome, _ = synth.make_config(args.workload, args.scale)
    t_gen = time.time() - t0
    ctx = _lib.Context()
    ctx.set_contigs(batch.contig_names)
    runtime.ensure_genome(ctx, genome, batch.contig_names)
    ctx.upload(batch)
    st = ctx.collect()
    sigs, ins = ctx.fetch_signatures(0, st)
    st2 = ctx.collect()
    sigs2, ins2 = ctx.fetch_signatures(0, st2)
    assert sigs.tobytes() == sigs2.tobytes() and ins.tobytes() == ins2.tobytes(), "COLLECT is not deterministic"
    key = (sigs["aln_idx"].astype(np.uint64) << np.uint64(32)) | sigs["ordinal"].astype(np.uint64)
    assert (np.diff(key.astype(np.int64)) > 0).all() if len(key) > 1 else True, "signatures not in emission order"
    ctx.use_collected(0)
    cst, clusters, members = ctx.cluster()
    tm = ctx.timings()
    order, part_off = ctx.fetch_partitions(len(sigs))
    n = len(sigs)
    part_of = np.empty(n, dtype=np.int64)
    sizes = np.diff(part_off.astype(np.int64))
    part_of[order] = np.repeat(np.arange(len(sizes)), sizes)
    # 2. structural properties
    mo = clusters["member_off"].astype(np.int64); sz = clusters["size"].astype(np.int64)
    assert (sz > 0).all() and int(sz.sum()) == len(members)
    first = members[mo]
    cl_part = part_of[first]
    cl_of_member = np.repeat(np.arange(len(clusters)), sz)
    assert (part_of[members] == cl_part[cl_of_member]).all(), "cluster spans partitions"
    assert (sigs["type"][members] == clusters["type"][cl_of_member]).all(), "cluster mixes types"
    assert len(np.unique(members)) == len(members), "a signature is in two clusters"
    kept_per_part = np.bincount(part_of[members], minlength=len(sizes))
    assert (kept_per_part <= np.minimum(sizes, 100)).all(), "more members than sampled"
    # 3. order of unilocal lists
    rank = np.argsort(np.argsort(np.array(batch.contig_names)))
    for t in (0, 1, 2):
        sel = np.nonzero(clusters["type"] == t)[0]
        if len(sel) < 2:
            continue
        c = rank[sigs["contig1"][first[sel]]]
        mid = (clusters["start"][sel] + clusters["end"][sel]).astype(np.int64)
        k = c.astype(np.int64) * (1 << 40) + mid
        assert (np.diff(k) >= 0).all(), "unilocal clusters not sorted"
    # 4. oracle spot check
    rng = np.random.default_rng(0)
    cand = np.nonzero((sizes >= 2) & (sizes <= 100))[0]
    pick = rng.choice(cand, size=min(args.spot, len(cand)), replace=False) if len(cand) else []
    by_part = {}
    for ci in np.nonzero(np.isin(cl_part, pick))[0]:
        by_part.setdefault(int(cl_part[ci]), []).append(int(ci))
    names = batch.contig_names
    p = orc.Params()
    blob = ins.tobytes()

    def oracle_sig(i):
        s = sigs[i]; t = _lib.TYPE_NAMES[s["type"]]; fl = int(s["flags"])
        o = orc.Sig(t, names[s["contig1"]], int(s["start"]), int(s["end"]), "x", int(s["qname_id"]))
        if t == "INS":
            o.sequence = blob[int(s["seq_off"]):int(s["seq_off"]) + int(s["seq_len"])].decode()
        elif t == "INV":
            o.direction = _lib.INV_DIRECTIONS[(fl >> 4) & 7]
        elif t == "DUP_TAN":
            o.copies = int(s["copies"]); o.fully_covered = bool(fl & 2)
        elif t in ("DUP_INT", "BND"):
            o.contig2 = names[s["contig2"]]; o.pos = int(s["pos"]); o.dir1 = "rev" if fl & 4 else "fwd"; o.dir2 = "rev" if fl & 8 else "fwd"
        return o

    def check_partition(pi, idx):
        """idx: signature indices the reference clusters for partition pi, in its member order"""
        osigs = [oracle_sig(i) for i in idx]
        index_of = {id(o): int(i) for o, i in zip(osigs, idx)}
        want = orc.consolidate(orc.clusters_from_partitions([osigs], genome, p), osigs[0].type in ("DUP_TAN", "BND", "DUP_INT"))
        got = by_part.get(int(pi), [])
        # within one partition the GPU list order is partition order for bilocal types; unilocal lists are re-sorted globally,
        # so compare as sets keyed by member tuple
        want_map = {tuple(index_of[id(m)] for m in c.members): c for c in want}
        got_map = {tuple(int(x) for x in members[mo[ci]:mo[ci] + sz[ci]]): ci for ci in got}
        assert set(want_map) == set(got_map), ("membership differs", int(pi))
        for k, ci in got_map.items():
            w = want_map[k]
            assert (int(clusters["start"][ci]), int(clusters["end"][ci])) == (w.start, w.end), ("coords", int(pi))
            assert abs(float(clusters["score"][ci]) - w.score) <= 1e-9 * max(1, abs(w.score)), ("score", int(pi))

    # 5. partitions above 100: replay the sampling stream with the stdlib, then the oracle on exactly the sampled signatures
    import random
    ptype = sigs["type"][order[part_off[:-1]]]
    large_checked = 0; large_total = 0
    large_pick = {}
    for t in range(6):
        large = np.nonzero((ptype == t) & (sizes > 100))[0]          # partition order = sorted-key order inside a type
        large_total += len(large)
        if len(large) == 0:
            continue
        want_n = max(1, args.spot_large // 3)
        chosen = set(large[:want_n // 2].tolist()) | set(rng.choice(large, size=min(want_n, len(large)), replace=False).tolist())
        random.seed(1524)
        for pi in large.tolist():
            picks = random.sample(range(int(sizes[pi])), 100)
            if pi in chosen:
                large_pick[pi] = picks
    by_part = {}
    wanted = set(int(x) for x in pick) | set(large_pick)
    for ci in np.nonzero(np.isin(cl_part, np.fromiter(wanted, dtype=np.int64, count=len(wanted))))[0]:
        by_part.setdefault(int(cl_part[ci]), []).append(int(ci))
    checked = 0
    for pi in pick:
        check_partition(pi, order[part_off[pi]:part_off[pi + 1]])
        checked += 1
    for pi, picks in large_pick.items():
        base = int(part_off[pi])
        check_partition(pi, [int(order[base + k]) for k in picks])
        large_checked += 1
    out = {"workload": args.workload, "scale": args.scale, "records": batch.n, "signatures": int(n), "partitions": int(len(sizes)),
           "largest_partition": int(sizes.max()) if len(sizes) else 0, "partitions_over_100": int((sizes > 100).sum()),
           "clusters": int(len(clusters)), "myers_pairs": int(cst.myers_pairs), "myers_cells": int(cst.myers_cells),
           "oracle_spot_checked_partitions": checked, "oracle_checked_sampled_partitions": large_checked, "sampled_partitions": int(large_total),
           "input_generation_s": round(t_gen, 1),
           "stage_ms": {k: round(v, 3) for k, v in tm.items() if v}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
