"""-m gpu: hand-built records for the corners of COLLECT the synthetic generator does not reach
(empty / ragged input, missing SEQ, hard clips, N operations, malformed SA entries, MAPQ overflow,
Python slice clamping, --all_bnds twins) — CUDA path vs oracle, plus the error codes for inputs the
reference itself raises on."""
import numpy as np
import pytest

from gpu_common import run_gpu, sig_rows
from svim_b200 import _lib, synth
from svim_b200.records import BatchBuilder
from oracle import svim_oracle as orc

pytestmark = pytest.mark.gpu

NAMES = ["chr1", "chr10", "chr2"]


def _collect(ctx, batch, **overrides):
    ctx.set_params(_lib.Params.from_options(None, **overrides))
    ctx.set_contigs(batch.contig_names)
    st = ctx.collect_host(batch)
    sigs, ins = ctx.fetch_signatures(0, st)
    tw = []
    if st.n_twin_signatures:
        t, ti = ctx.fetch_signatures(1, st)
        tw = sig_rows(t, ti, batch)
    return st, sig_rows(sigs, ins, batch), tw


def _oracle(batch, **overrides):
    s, t = orc.collect(batch, orc.Params(**overrides))
    return [list(x.as_tuple()) for x in s], [list(x.as_tuple()) for x in t]


def test_empty_and_fully_filtered_input(gpu_ctx):
    b = BatchBuilder(NAMES, [10**6] * 3).finish()
    st, rows, tw = _collect(gpu_ctx, b)
    assert st.n_signatures == 0 and rows == [] and st.n_primaries == 0
    b = BatchBuilder(NAMES, [10**6] * 3)
    b.add("u", 4, -1, -1, 0, "", "ACGT")                       # unmapped
    b.add("s", 256, 0, 100, 60, "100M50D100M", "A" * 200)      # secondary
    b.add("l", 0, 0, 100, 19, "100M50D100M", "A" * 200)        # MAPQ below the threshold
    st, rows, tw = _collect(gpu_ctx, b.finish())
    assert rows == [] and st.n_primaries == 0
    gpu_ctx.use_collected(0)
    cst, clusters, members = gpu_ctx.cluster()
    assert cst.n_clusters_total == 0 and len(clusters) == 0


def test_ragged_cigars_and_quirks(gpu_ctx):
    b = BatchBuilder(NAMES, [10**6] * 3)
    seq = "ACGT" * 200
    b.add("r0", 0, 0, 1000, 60, "", None)                                        # mapped flag but no CIGAR
    b.add("r1", 0, 0, 1000, 60, "50M", seq[:50])                                 # < one 16-byte group
    b.add("r2", 0, 0, 2000, 60, "10S30M100N20M45D15M60I40M", seq[:175])          # N does not advance pos_ref (SVIM_intra.py)
    b.add("r3", 16, 1, 3000, 60, "5H10S100M41I100M39D100M40D10M7S3H", seq[:268]) # hard+soft clips, reverse strand
    b.add("r4", 0, 2, 4000, 60, "100M50I100M", None)                             # SEQ '*': sequence ""
    b.add("r5", 2048, 0, 5000, 60, "20H100M45I100M", seq[:245])                  # supplementary: CIGAR analysis only
    b.add("r6", 0, 0, 6000, 60, "100M40I", seq[:120])                            # slice clamps at the end of SEQ
    b.add("r7", 0, 0, 7000, 60, "".join("%dM1I%dM1D" % (3 + k % 5, 2 + k % 3) for k in range(700)) + "60D9M", seq[:100])  # > 512 ops, event in the tail
    batch = b.finish()
    for kw in ({}, {"all_bnds": True}, {"min_sv_size": 1}, {"min_sv_size": 39, "min_mapq": 60}):
        st, rows, tw = _collect(gpu_ctx, batch, **kw)
        want, want_t = _oracle(batch, **kw)
        assert rows == want, kw
        assert tw == want_t, kw


def test_sa_tag_corner_cases(gpu_ctx):
    b = BatchBuilder(NAMES, [10**6] * 3)
    seq = "ACGT" * 2500
    sa_ok = "chr1,20001,+,5000S5000M,60,10;"
    b.add("a", 0, 0, 10000, 60, "5000M5000S", seq, sa_ok)                                     # split DEL
    b.add("b", 0, 0, 10000, 60, "5000M5000S", seq, "chr1,20001,+,5000S5000M,19,10;")          # SA MAPQ below threshold
    b.add("c", 0, 0, 10000, 60, "5000M5000S", seq, "chr1,20001,+,5000S5000M,300,10;")         # MAPQ > 255 -> 0 -> filtered
    b.add("d", 0, 0, 10000, 60, "5000M5000S", seq, "chr1,20001,+,5000S5000M,60;" + sa_ok)     # 5 fields: warning + skipped
    b.add("e", 0, 0, 10000, 60, "5000M5000H", seq[:5000], sa_ok)                              # hard-clipped primary: SA ignored
    b.add("f", 0, 0, 10000, 60, "5000M5000S", None, "chr1,10001,+,5000S60I4940M,60,1;")       # split INS without SEQ
    b.add("g", 16, 0, 10000, 60, "5000S5000M", seq, "chr10,777,-,5000M5000S,60,3;chr2,9,+,2500S10M7490S,60,1;;")  # 3 segments, 2 contigs, empty element
    b.add("h", 0, 0, 10000, 60, "4000M6000S", seq, "chr1,10901,-,4000S6000M,60,0;")           # inversion-type junction
    b.add("i", 0, 0, 50000, 60, "5000M5000S", seq, "chr1,49001,+,5000S5000M,60,0;")           # tandem duplication
    b.add("j", 0, 0, 50000, 60, "3000M7000S", seq, "chr1,52001,+,3000S100M6900S,60,0;chr1,52201,+,3100S6900M,60,0;")  # overlapping chain
    batch = b.finish()
    for kw in ({}, {"all_bnds": True}, {"max_sv_size": 3000}, {"segment_gap_tolerance": 0, "segment_overlap_tolerance": 0}):
        st, rows, tw = _collect(gpu_ctx, batch, **kw)
        want, want_t = _oracle(batch, **kw)
        assert rows == want, kw
        assert tw == want_t, kw
        assert st.n_sa_bad_fields == 1 and st.n_data_errors == 0


def test_inputs_the_reference_raises_on_return_errors(gpu_ctx):
    seq = "ACGT" * 2500
    for sa in ("chrUn,20001,+,5000S5000M,60,10;",          # unknown contig: bam.getrname(-1) raises
               "chr1,abc,+,5000S5000M,60,10;",              # int('abc') raises
               "chr1,20001,+,*,60,10;"):                    # no CIGAR: reference_end is None
        b = BatchBuilder(NAMES, [10**6] * 3)
        b.add("x", 0, 0, 10000, 60, "5000M5000S", seq, sa)
        batch = b.finish()
        gpu_ctx.set_params(_lib.Params.from_options(None))
        gpu_ctx.set_contigs(batch.contig_names)
        st = gpu_ctx.collect_host(batch)
        assert st.n_data_errors == 1


def test_reads_with_hundreds_of_segments_take_the_large_read_pass(gpu_ctx):
    """The reference has no limit on the segments of a read (SVIM_inter.py:24-49).  Reads above the 64 the per-thread arrays hold
    (70, 200 and 700 SA entries here: deletions, tandem duplications, junctions to other contigs and strands) go through
    k_segment_chain_big and must still be the oracle's signatures, in emission order, next to ordinary reads."""
    import random
    rng = random.Random(5)
    seq = "ACGT" * 25000
    b = BatchBuilder(NAMES, [10**7] * 3)
    b.add("small", 0, 0, 5000, 60, "5000M5000S", seq[:10000], "chr1,20001,+,5000S5000M,60,10;")
    for name, k_seg in (("y70", 70), ("y200", 200), ("y700", 700)):
        step = 100000 // (k_seg + 1)
        sa = []
        ref = 200000
        for k in range(k_seg):
            lead = step * (k + 1)
            kind = rng.random()
            contig, strand = "chr1", "+"
            if kind < 0.5:
                ref += step + rng.choice([0, 60, 500, 7000])            # next segment downstream: nothing / deletion of various sizes
            elif kind < 0.7:
                ref -= rng.choice([80, 300, 2000])                      # overlap on the reference: tandem duplication
            elif kind < 0.85:
                contig = rng.choice(["chr10", "chr2"]); ref = rng.randrange(1000, 900000)
            else:
                strand = "-"
            sa.append("%s,%d,%s,%dS%dM%dS,60,0;" % (contig, ref + 1, strand, lead, step, 100000 - lead - step))
        b.add(name, 0, 0, 100000, 60, "%dM%dS" % (step, 100000 - step), seq, "".join(sa))
    b.add("small2", 0, 0, 300000, 60, "5000M5000S", seq[:10000], "chr1,320001,+,5000S5000M,60,10;")
    batch = b.finish()
    for kw in ({}, {"all_bnds": True}, {"max_sv_size": 3000}):
        st, rows, tw = _collect(gpu_ctx, batch, **kw)
        want, want_t = _oracle(batch, **kw)
        assert rows == want, kw
        assert tw == want_t, kw
        assert st.n_data_errors == 0 and len(rows) > 300


def test_cluster_state_errors(gpu_ctx):
    ctx = _lib.Context()
    with pytest.raises(_lib.SvimGpuError) as e:
        ctx.collect()
    assert e.value.code == -3
    with pytest.raises(_lib.SvimGpuError):
        ctx.use_collected(0)
    ctx.close()


def test_packed_cigar16_upload_expands_to_the_same_words(gpu_ctx, golden):
    """svim_aln_soa.cigar16 (16-bit packed CIGAR stream, half the PCIe bytes): the device expansion must give back the caller's
    uint32 words record for record — zero padding included — and COLLECT must not notice the difference.  Operations of 4096,
    2^24 and 2^28-1 bases take the extension words; a stream that does not hold n_cigar operations is refused."""
    from svim_b200 import io as sio
    rng = np.random.default_rng(3)
    b = BatchBuilder(NAMES, [10**9] * 3)
    b.add("long", 0, 0, 100, 60, [(0, 5), (2, 4095), (0, 1), (2, 4096), (1, 70000), (3, 20000000), (0, 3), (4, 2**28 - 1)], None, None)
    b.add("nine", 0, 0, 200, 60, [(0, 10)] * 9, None, None)
    b.add("empty", 4, -1, -1, 0, "", None, None)
    for k in range(300):
        n = int(rng.choice([1, 7, 8, 9, 255, 256, 257, 1000, 3000]))
        ops = [(int(rng.choice([0, 0, 0, 1, 2, 7, 8, 4])), int(rng.choice([1, 2, 3, 10, 50, 4095, 4096, 5000]))) for _ in range(n)]
        b.add("r%d" % k, 0, 0, 1000 + k, 60, ops, None, None)
    batch = b.finish()
    packed = batch.pack_cigar16(3)
    un = sio.unpack_cigar16(packed.cigar16, packed.cigar16_off, batch.n_cigar)
    for i in range(batch.n):
        assert np.array_equal(un[i], batch.cigar[int(batch.cigar_off[i]):int(batch.cigar_off[i]) + int(batch.n_cigar[i])])
    gpu_ctx.set_params(_lib.Params.from_options(None)); gpu_ctx.set_contigs(batch.contig_names)
    gpu_ctx.upload(packed)
    assert np.array_equal(gpu_ctx.download_cigar(batch.cigar.size), batch.cigar)
    # a stream that does not decode to n_cigar operations
    broken = b.finish().pack_cigar16(1)
    broken.cigar16[int(broken.cigar16_off[1])] = 0x000F
    with pytest.raises(_lib.SvimGpuError):
        gpu_ctx.upload(broken)
    # COLLECT + CLUSTER from the packed upload == golden (plain upload is what every other test uses)
    from gpu_common import run_gpu, assert_clusters_equal
    for name in ("mini_mixed", "mini_hotspot"):
        gb, genome, exp = golden(name)
        gb.pack_cigar16(2)
        try:
            rows, trows, clusters, st, cst = run_gpu(gpu_ctx, gb, genome, exp["params"])
        finally:
            gb.cigar16 = None; gb.cigar16_off = None
        assert rows == exp["signatures"]
        assert_clusters_equal(clusters, exp["clusters"])


def test_packed_cigar8_upload_expands_to_the_same_words(gpu_ctx, golden):
    """svim_aln_soa.cigar8 (one byte per operation under 16 bases, extension bytes before the longer ones): the device expansion
    must give back the caller's uint32 words record for record — zero padding included — wherever the extension chains fall
    (inside a lane's 16 bytes, across lanes, across the 512-byte rounds), and COLLECT must not notice the difference.
    The same warp loop is replayed on the host in tests/test_host_units.py."""
    from svim_b200 import io as sio
    rng = np.random.default_rng(8)
    b = BatchBuilder(NAMES, [10**9] * 3)
    b.add("long", 0, 0, 100, 60, [(0, 5), (2, 15), (0, 16), (2, 255), (1, 256), (3, 20000000), (0, 3), (4, 2**28 - 1), (0, 0)], None, None)
    b.add("sixext", 0, 0, 150, 60, [(2, 2**28 - 1)] * 100, None, None)
    b.add("nine", 0, 0, 200, 60, [(0, 10)] * 9, None, None)
    b.add("empty", 4, -1, -1, 0, "", None, None)
    for k in range(400):
        n = int(rng.choice([1, 15, 16, 17, 31, 32, 33, 511, 512, 513, 1000, 3000]))
        p_ext = float(rng.choice([0.0, 0.1, 0.5]))
        ops = [(int(rng.integers(0, 9)), int(rng.choice([16, 200, 4096, 70000, 2**27])) if rng.random() < p_ext else int(rng.integers(0, 16))) for _ in range(n)]
        b.add("r%d" % k, 0, 0, 1000 + k, 60, ops, None, None)
    batch = b.finish()
    packed = batch.pack_cigar8(3)
    un = sio.unpack_cigar8(packed.cigar8, packed.cigar8_off, batch.n_cigar)
    for i in range(batch.n):
        assert np.array_equal(un[i], batch.cigar[int(batch.cigar_off[i]):int(batch.cigar_off[i]) + int(batch.n_cigar[i])])
    gpu_ctx.set_params(_lib.Params.from_options(None)); gpu_ctx.set_contigs(batch.contig_names)
    gpu_ctx.upload(packed)
    assert np.array_equal(gpu_ctx.download_cigar(batch.cigar.size), batch.cigar)
    # a stream that does not decode to n_cigar operations
    broken = b.finish().pack_cigar8(1)
    broken.cigar8[int(broken.cigar8_off[2])] = 0x0F
    with pytest.raises(_lib.SvimGpuError):
        gpu_ctx.upload(broken)
    # COLLECT + CLUSTER from the packed upload == golden
    from gpu_common import run_gpu, assert_clusters_equal
    for name in ("mini_mixed", "mini_hotspot"):
        gb, genome, exp = golden(name)
        gb.pack_cigar8(2)
        try:
            rows, trows, clusters, st, cst = run_gpu(gpu_ctx, gb, genome, exp["params"])
        finally:
            gb.cigar8 = None; gb.cigar8_off = None
        assert rows == exp["signatures"]
        assert_clusters_equal(clusters, exp["clusters"])

