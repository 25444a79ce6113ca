"""-m gpu: the on-GPU BAM decoder (svimgpu_decode_bam: BGZF inflate, record boundaries, rows, blobs, read-name ids on the device)
against the host decoder, byte for byte, and COLLECT + CLUSTER straight from the resident buffer against the host-buffer path."""
import numpy as np
import pytest

from svim_b200 import _lib, io as sio, runtime

pytestmark = pytest.mark.gpu


def _assert_same(a, b):
    assert a.n == b.n and a.contig_names == b.contig_names and a.sort_order == b.sort_order
    for name, _ in a.FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for blob in ("cigar", "seq", "sa"):
        assert np.array_equal(getattr(a, blob), getattr(b, blob)), blob
    assert [a.qname(int(i)) for i in a.qname_id] == [b.qname(int(i)) for i in b.qname_id]
    assert a.qnames == b.qnames                                   # first-appearance numbering, like the host decoder


def test_gpu_bam_decoder_equals_host_decoder(tmp_path):
    from svim_b200 import synth
    batch, _genome, _ = synth.make_config("config2", 0.02, with_genome=False)
    p = str(tmp_path / "c2.bam")
    sio.write_bam_native(p, batch, threads=8)
    stats = {}
    _assert_same(sio.read_bam_gpu(p, stats=stats), sio.read_bam_native(p))
    assert sio.BAM_DECODE_COUNTS["gpu"] >= 1
    print("gpu decode stages:", {k: round(v, 3) if isinstance(v, float) else v for k, v in stats.items()})


def test_gpu_bam_decoder_records_spanning_chunks(tmp_path):
    from svim_b200.records import BatchBuilder
    rng = np.random.default_rng(9)
    b = BatchBuilder(["c1", "c2"], [50_000_000, 1000], "coordinate")
    pos = 0
    for k in range(40):
        n = int(rng.choice([50, 3000, 70_000, 400_000]) if k % 3 else 20_000)
        pos += int(rng.integers(1, 1000))
        ops, left = [], n
        while left > 0:
            ln = int(min(left, rng.integers(1, 40 + n // 500))); ops.append((int(rng.choice([0, 0, 0, 1, 7, 8])), ln)); left -= ln
        b.add("q%d" % (k % 30), 0, 0, pos, 60, ops, "".join(rng.choice(list("ACGTN"), size=n)), "c1,%d,+,%dM,60,0;" % (k + 1, n) if k % 4 == 0 else None)
    b.add("tail", 4, -1, -1, 0, "", None, None)
    p = str(tmp_path / "big.bam")
    sio.write_bam_native(p, b.finish(), threads=4)
    _assert_same(sio.read_bam_gpu(p), sio.read_bam_native(p))


def test_collect_and_cluster_from_the_resident_bam_buffer(tmp_path, golden):
    """file -> svimgpu_decode_bam -> svimgpu_collect -> svimgpu_cluster with nothing crossing back to the host in between must give the
    reference's golden signatures and clusters; a truncated file is declined and the host decoder's error surfaces instead."""
    from gpu_common import sig_rows, cluster_rows, assert_clusters_equal
    for name in ("mini_mixed", "mini_hotspot"):
        batch, genome, exp = golden(name)
        p = str(tmp_path / (name + ".bam"))
        sio.write_bam_native(p, batch, threads=2)
        ctx = _lib.Context(device=0)
        ctx.set_params(_lib.Params.from_options(None, **exp["params"]))
        rb = sio.decode_bam_resident(p, ctx)
        assert isinstance(rb, sio.ResidentBatch) and rb.n == batch.n
        ctx.set_contigs(rb.contig_names)
        st = ctx.collect()
        sigs, ins = ctx.fetch_signatures(0, st)
        rows = sig_rows(sigs, ins, rb)
        assert rows == exp["signatures"], name
        ctx.genome_key = None
        runtime.ensure_genome(ctx, genome, rb.contig_names)
        ctx.use_collected(0)
        cst, clusters, members = ctx.cluster()
        assert_clusters_equal(cluster_rows(clusters, members, rows), exp["clusters"])
        ctx.close()
    # a file cut in the middle of a block: the device decoder declines, the fallback is counted, the host decoder reports the damage
    raw = open(p, "rb").read()
    bad = str(tmp_path / "cut.bam")
    open(bad, "wb").write(raw[:len(raw) * 2 // 3])
    before = dict(sio.BAM_DECODE_COUNTS)
    ctx = _lib.Context(device=0)
    with pytest.raises(Exception):
        sio.decode_bam_resident(bad, ctx)
    ctx.close()
    assert sio.BAM_DECODE_COUNTS["gpu"] == before["gpu"]
