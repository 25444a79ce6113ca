"""-m gpu: the parity block bench.py prints (svim_b200.rows.prefix_parity): CUDA results of a whole input against the oracle run
on its first records only — signatures of those records in order, and every cluster of every partition made only of their
signatures, sampled partitions (above 100, SVIM_clustering.py:129-134) included.  Also checks that the check itself detects a
corrupted cluster, and that a rank-local failure is reported instead of leaving peers in a collective (single-rank half)."""
import numpy as np
import pytest

from svim_b200 import _lib, rows, synth, runtime

pytestmark = pytest.mark.gpu


def _oracle_prefix(batch, genome, n):
    from oracle import svim_oracle as orc
    p = orc.Params()
    sigs, _ = orc.collect(batch.slice(0, n), p)
    res = orc.cluster(sigs, genome, p)
    index_of = {id(x): i for i, x in enumerate(sigs)}
    want = {name: [[c.type, c.contig, c.start, c.end, c.dest_contig, c.dest_start, c.dest_end, c.score, c.size, c.std_span, c.std_pos, c.dir1, c.dir2,
                    [index_of[id(m)] for m in c.members]] for c in cl] for name, cl in zip(rows.TYPES_RETURN_ORDER, res)}
    return [list(x.as_tuple()) for x in sigs], want


@pytest.mark.parametrize("workload,scale,frac", [("config2", 0.004, 0.5), ("config5", 0.004, 0.6)])
def test_prefix_parity_of_a_larger_run(workload, scale, frac):
    batch, genome, _ = synth.make_config(workload, scale)
    ctx = _lib.Context(device=0)
    ctx.set_contigs(batch.contig_names)
    ctx.genome_key = None
    runtime.ensure_genome(ctx, genome, batch.contig_names)
    st = ctx.collect_host(batch)
    sigs, ins = ctx.fetch_signatures(0, st)
    sigs = np.array(sigs); ins = np.array(ins)
    ctx.use_collected(0)
    cst, clusters, members = ctx.cluster()
    order, part_off = ctx.fetch_partitions(len(sigs))
    n = int(batch.n * frac)
    want_sigs, want_clusters = _oracle_prefix(batch, genome, n)
    got = rows.prefix_parity(batch, n, sigs, ins, clusters, members, order, part_off, want_sigs, want_clusters)
    assert got["mismatches"] == 0, got
    assert got["signatures_compared"] > 20 and got["clusters_compared"] > 5, got
    if workload == "config5":
        assert got["sampled_partitions_compared"] > 0, got            # the host-replayed sampling stream is covered
    # the check must see a wrong cluster: shift one compared cluster's start
    first = members[clusters["member_off"].astype(np.int64)]
    pre = np.nonzero(np.asarray(sigs["aln_idx"])[first] < n // 4)[0]
    bad = clusters.copy(); bad["start"][pre[0]] += 1
    assert rows.prefix_parity(batch, n, sigs, ins, bad, members, order, part_off, want_sigs, want_clusters)["mismatches"] >= 1
    # ... and a wrong signature
    bad_s = sigs.copy(); bad_s["end"][0] += 1
    assert rows.prefix_parity(batch, n, bad_s, ins, clusters, members, order, part_off, want_sigs, want_clusters)["mismatches"] >= 1
    # digest: equal bytes <=> equal digest
    assert rows.result_digest(clusters, members, sigs) == rows.result_digest(clusters.copy(), members.copy(), sigs.copy())
    assert rows.result_digest(bad, members, sigs) != rows.result_digest(clusters, members, sigs)
    ctx.close()


def test_cluster_stats_counters_match_host_counts(golden):
    """n_partitions / large_partitions / duplicate_signatures / n_clusters per type come from device counters now."""
    from gpu_common import run_gpu
    from oracle import svim_oracle as orc
    batch, genome, exp = golden("mini_hotspot")
    ctx = _lib.Context(device=0)
    rows_, _t, clusters, st, cst = run_gpu(ctx, batch, genome, exp["params"])
    stats = {}
    sigs, _ = orc.collect(batch, orc.Params(**exp["params"]))
    orc.cluster(sigs, genome, orc.Params(**exp["params"]), stats=stats)
    for i, t in enumerate(("DEL", "INS", "INV", "DUP_TAN", "BND", "DUP_INT")):
        code = _lib.TYPE_CODE[t]
        if t in stats and "partitions" in stats[t]:
            assert cst.n_partitions[code] == stats[t]["partitions"], (t, cst.n_partitions[code], stats[t])
            assert cst.n_clusters[code] == stats[t]["clusters"], t
            assert cst.large_partitions[code] == stats[t]["large_partitions"], t
            assert cst.duplicate_signatures[code] == stats[t]["duplicate_signatures"], t
    ctx.close()
