"""EXPERIMENTAL GPU code that has not run on a GPU yet (written after the round's GPU budget was spent): kept out of both
`-m gpu` and `-m "not gpu"` runs unless SVIM_RUN_NEXT=1.

    SVIM_RUN_NEXT=1 python -m pytest tests/test_gpu_next.py -q          # on a GPU box

On-GPU BAM decoder (svim_b200/csrc_next/bamgpu.cu, svim_b200.io.read_bam_gpu) against the host decoder."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu_next, pytest.mark.skipif(os.environ.get("SVIM_RUN_NEXT") != "1", reason="experimental: set SVIM_RUN_NEXT=1 on a GPU box")]


def _assert_same(a, b):
    assert a.n == b.n and a.contig_names == b.contig_names and a.sort_order == b.sort_order
    for name, _ in a.FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for blob in ("cigar", "seq", "sa"):
        assert np.array_equal(getattr(a, blob), getattr(b, blob)), blob
    assert [a.qname(int(i)) for i in a.qname_id] == [b.qname(int(i)) for i in b.qname_id]


def test_gpu_bam_decoder_equals_host_decoder(tmp_path):
    from svim_b200 import synth, io as sio
    batch, _genome, _ = synth.make_config("config2", 0.02, with_genome=False)
    p = str(tmp_path / "c2.bam")
    sio.write_bam_native(p, batch, threads=8)
    stats = {}
    _assert_same(sio.read_bam_gpu(p, stats=stats), sio.read_bam_native(p))
    print("gpu decode stages (ms):", stats)


def test_gpu_bam_decoder_records_spanning_chunks(tmp_path):
    from svim_b200.records import BatchBuilder
    from svim_b200 import io as sio
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
    b.add("tail", 4, -1, -1, 0, "", None)
    p = str(tmp_path / "big.bam")
    sio.write_bam_native(p, b.finish(), threads=4)
    _assert_same(sio.read_bam_gpu(p), sio.read_bam_native(p))
