"""-m gpu: hand-built records for the corners of COLLECT the synthetic generator does not reach
(empty / ragged input, missing SEQ, hard clips, N operations, malformed SA entries, MAPQ overflow,
Python slice clamping, --all_bnds twins) — CUDA path vs oracle, plus the error codes for inputs the
reference itself raises on."""
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
    # more than SVIM_MAX_SEGMENTS alignment segments in one read is a documented limit
    b = BatchBuilder(NAMES, [10**6] * 3)
    b.add("y", 0, 0, 10000, 60, "100M9900S", seq, "".join("chr1,%d,+,%dS100M%dS,60,0;" % (20000 + 300 * k, 100 + 100 * k, 9800 - 100 * k) for k in range(70)))
    with pytest.raises(_lib.SvimGpuError) as e:
        gpu_ctx.collect_host(b.finish())
    assert e.value.code == -4


def test_cluster_state_errors(gpu_ctx):
    ctx = _lib.Context()
    with pytest.raises(_lib.SvimGpuError) as e:
        ctx.collect()
    assert e.value.code == -3
    with pytest.raises(_lib.SvimGpuError):
        ctx.use_collected(0)
    ctx.close()
