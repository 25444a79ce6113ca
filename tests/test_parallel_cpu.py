"""Host-side multi-process logic on CPU (gloo, world_size 2): shard layout, rendezvous bookkeeping
and the property the multi-GPU design rests on — COLLECT over contiguous record shards,
concatenated in rank order with record indices rebased, equals COLLECT over the whole input."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden, ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from svim_b200 import parallel
    from oracle import svim_oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch, genome, exp = load_golden("mini_mixed")
    lo, hi = parallel.shard_ranges(batch.n_cigar, world)[rank]
    shard = parallel.shard_batch(batch, rank, world)          # compact: the rank's own blobs, offsets rebased
    assert shard.n == hi - lo and shard.cigar.size < batch.cigar.size and shard.cigartuples(0) == batch.cigartuples(lo)
    base, total, sizes = parallel.exchange_layout(shard.n)
    sigs, _ = orc.collect(shard, orc.Params())
    rows = [list(s.as_tuple()) for s in sigs]
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    q.put((rank, lo, hi, base, total, sizes, [r for part in gathered for r in part]))
    dist.barrier()
    dist.destroy_process_group()


def test_contiguous_shards_reproduce_emission_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch, genome, exp = load_golden("mini_mixed")
    (r0, lo0, hi0, b0, t0, s0, all0), (r1, lo1, hi1, b1, t1, s1, all1) = res
    assert (lo0, hi1) == (0, batch.n) and hi0 == lo1                 # ranges tile the input
    assert (b0, b1) == (0, hi0) and t0 == t1 == batch.n and s0 == s1 == [hi0 - lo0, hi1 - lo1]
    assert all0 == all1 == exp["signatures"]                         # concatenation in rank order == full COLLECT


def test_shard_ranges_balance_and_cover():
    from svim_b200 import parallel
    rng = np.random.default_rng(0)
    n_cigar = rng.integers(0, 5000, 10000).astype(np.uint32)
    for world in (1, 2, 3, 8):
        r = parallel.shard_ranges(n_cigar, world)
        assert r[0][0] == 0 and r[-1][1] == len(n_cigar) and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        loads = [int(n_cigar[a:b].sum()) for a, b in r]
        assert max(loads) < 1.1 * (sum(loads) / world) + 5000


def test_compact_slice_and_generated_shards_equal_the_whole_input():
    """AlignmentBatch.compact_slice (what parallel.shard_batch uploads) and synth.generate_shard (bench.py --shard records: a rank
    materialises only its record range of ONE coordinate-sorted input) against the full batch, record for record."""
    from svim_b200 import synth
    names, lengths, reads, seed, pkw, gkw = synth.config_layout("config4", 0.002)
    assert names[:3] == ["chr1", "chr2", "chr3"] and names[-2:] == ["chrX", "chrY"] and len(names) == 24
    svs, alleles = synth.plant_svs(lengths, seed, **pkw)
    full = synth.generate(names, lengths, reads, seed, svs, alleles, **gkw)
    at = 0
    for r in range(3):
        sh, lo, total = synth.generate_shard(names, lengths, reads, seed, svs, alleles, r, 3, **gkw)
        cs = full.compact_slice(lo, lo + sh.n)
        assert total == full.n and lo == at
        for f, _ in full.FIELDS:
            assert np.array_equal(getattr(sh, f), getattr(cs, f)), f
        for blob in ("cigar", "seq", "sa"):
            assert np.array_equal(getattr(sh, blob), getattr(cs, blob)), blob
        for i in (0, sh.n // 2, sh.n - 1):
            assert sh.cigartuples(i) == full.cigartuples(lo + i) and sh.sa_tag(i) == full.sa_tag(lo + i) and sh.sequence(i) == full.sequence(lo + i)
        at += sh.n
    assert at == full.n
