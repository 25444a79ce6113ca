"""Launched by torchrun (one rank per GPU): sharded COLLECT -> NCCL exchange -> sharded CLUSTER ->
NCCL gather must reproduce the reference's golden outputs on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    from conftest import load_golden
    from gpu_common import sig_rows, cluster_rows, assert_clusters_equal
    from svim_b200 import _lib, parallel, runtime
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    ctx = _lib.Context(device=local)
    parallel.init_comm(ctx)
    if rank == 0:
        print("inserted sequences:", "peer memory" if ctx.lib.svimgpu_peer_ins_active(ctx.h) else "gathered", flush=True)
    for name in ("mini_mixed", "mini_ins", "mini_hotspot", "mini_mixed_allbnds"):
        batch, genome, exp = load_golden(name)
        ctx.set_params(_lib.Params.from_options(None, **exp["params"]))
        ctx.set_contigs(batch.contig_names)
        ctx.genome_key = None
        runtime.ensure_genome(ctx, genome, batch.contig_names)
        shard = parallel.shard_batch(batch, rank, world)
        base, total, sizes = parallel.exchange_layout(shard.n)
        ctx.upload(shard)
        cst, xst, clst, clusters, members = parallel.collect_and_cluster(ctx, base)
        sigs, ins = ctx.fetch_signatures(0, xst)
        rows = sig_rows(sigs, ins, batch)
        assert rows == exp["signatures"], (name, rank, "signatures differ")
        assert_clusters_equal(cluster_rows(clusters, members, rows), exp["clusters"])
        assert [int(x) for x in sigs["aln_idx"]] == sorted(int(x) for x in sigs["aln_idx"])
        if rank == 0:
            print("multi-gpu ok:", name, "world", world, "signatures", len(rows), "clusters", int(clst.n_clusters_total), flush=True)
    # a rank that fails before the cluster exchange must not leave the others waiting in it, and nobody may keep a result:
    # the failing rank reports its own error, every other rank SVIMGPU_ERR_PEER (-7); the next call works again
    if world > 1:
        dist.barrier()          # the other ranks may still be fetching this rank's insertion bytes (include/svimgpu.h, exchange)
        os.environ["SVIM_TEST_FAIL_RANK"] = str(world - 1)
        ctx.collect()
        xst = _lib.CollectStats()
        ctx._check(ctx.lib.svimgpu_exchange_signatures(ctx.h, base, __import__("ctypes").byref(xst)))
        ctx.use_collected(0)
        try:
            ctx.cluster(sharded=True)
            raise AssertionError("sharded cluster returned a result although rank %d failed" % (world - 1))
        except _lib.SvimGpuError as e:
            assert e.code == (-5 if rank == world - 1 else -7), (rank, e.code, str(e))
        del os.environ["SVIM_TEST_FAIL_RANK"]
        cst, xst, clst, clusters, members = parallel.collect_and_cluster(ctx, base)
        assert int(clst.n_clusters_total) > 0
        if rank == 0:
            print("multi-gpu ok: rank failure is collective", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
