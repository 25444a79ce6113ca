"""-m gpu: the CUDA path (through the C ABI) against the reference's committed outputs and the
oracle.  Bit-exact on signatures, partitions, cluster membership/order and integer
coordinates; score / std within 1e-6 (tolerance stated in BASELINE.json north_star)."""
import random

import numpy as np
import pytest

from conftest import GOLDEN_NAMES
from gpu_common import run_gpu, assert_clusters_equal, sig_rows, cluster_rows
from svim_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_collect_and_cluster_match_reference_golden(gpu_ctx, golden, name):
    batch, genome, exp = golden(name)
    rows, trows, clusters, st, cst = run_gpu(gpu_ctx, batch, genome, exp["params"])
    assert len(rows) == len(exp["signatures"])
    for k, (a, b) in enumerate(zip(rows, exp["signatures"])):
        assert a == b, (k, a, b)
    assert trows == exp["all_bnds_signatures"]
    assert_clusters_equal(clusters, exp["clusters"])


def test_all_bnds_second_pass_matches_reference(gpu_ctx, golden):
    # svim:133-139: the --all_bnds extras are clustered in a second cluster_sv_signatures call
    batch, genome, exp = golden("mini_mixed_allbnds")
    rows, trows, clusters, st, cst = run_gpu(gpu_ctx, batch, genome, exp["params"], which=1)
    assert_clusters_equal(clusters, exp["all_bnds_clusters"])


def test_cigar_indel_kats_through_scan_kernel(gpu_ctx):
    # tests/test_intra.py:8-22 pushed through k_cigar_scan (pos_ref, len, type)
    from svim_b200.records import encode_cigar
    kats = [([(5, 10), (4, 20), (0, 10), (7, 10), (8, 5), (0, 5), (1, 50), (0, 30), (4, 25), (5, 15)], [(30, 50, 1)]),
            ([(5, 10), (4, 20), (0, 30), (2, 50), (0, 30), (4, 25), (5, 15)], [(30, 50, 2)]),
            ([(5, 10), (4, 20), (0, 30), (2, 40), (1, 50), (0, 30), (4, 25), (5, 15)], [(30, 40, 2), (70, 50, 1)]),
            ([(5, 10), (4, 20), (0, 30), (1, 40), (2, 50), (0, 30), (4, 25), (5, 15)], [(30, 40, 1), (30, 50, 2)])]
    for tuples, want in kats:
        got = gpu_ctx.cigar_indel(encode_cigar(tuples), 30)
        assert [(int(r[0]), int(r[2]), int(r[3])) for r in got] == want


def test_cigar_scan_long_and_ragged(gpu_ctx):
    """Random CIGARs of every length class (empty, < one 16-byte group, many 128-op groups, events in
    every lane/slot position) against the oracle's analyze_cigar_indel restatement."""
    from oracle import svim_oracle as orc
    from svim_b200.records import encode_cigar
    rng = random.Random(7)
    for n in [0, 1, 2, 3, 4, 5, 127, 128, 129, 511, 512, 513, 1000, 5000, 70000]:
        for rep in range(3):
            tuples = []
            for _ in range(n):
                op = rng.choice([0, 0, 0, 1, 2, 1, 2, 3, 4, 5, 6, 7, 8])
                ln = rng.choice([1, 2, 3, 10, 39, 40, 41, 300]) if rng.random() < 0.2 else rng.randint(1, 30)
                tuples.append((op, ln))
            want = [(r, l, 1 if t == "INS" else 2) for r, q, l, t in orc.cigar_indels(tuples, 40)]
            got = gpu_ctx.cigar_indel(encode_cigar(tuples), 40)
            assert [(int(r[0]), int(r[2]), int(r[3])) for r in got] == want, (n, rep)


def test_edit_distance_kernel_matches_oracle(gpu_ctx):
    from oracle import editdist
    rng = random.Random(3)
    pairs = [(b"", b""), (b"", b"ACGT"), (b"ACGT", b""), (b"A", b"A"), (b"A", b"C"), (b"kitten", b"sitting")]
    alph = [b"ACGT", b"ACGTN", b"ACGTNRYKM=acgtn*-"]
    for n in [1, 5, 63, 64, 65, 127, 128, 129, 255, 256, 257, 500, 511, 513, 767, 769, 1023, 1025, 1535, 1537, 2047, 2048, 2049,
              3071, 3073, 4095, 4097, 5000, 6143, 6145, 8191, 8192, 8193, 9000, 12289, 20000]:
        for k in range(3):
            al = alph[k % 3]
            a = bytes(rng.choice(al) for _ in range(n))
            if k == 2:
                b = bytes(rng.choice(al) for _ in range(max(1, n + rng.randint(-n // 2, n // 2))))
            else:
                b = bytearray(a)
                for _ in range(rng.randint(0, max(1, n // 6))):
                    r = rng.random()
                    if r < 0.3 and b:
                        del b[rng.randrange(len(b))]
                    elif r < 0.6:
                        b.insert(rng.randint(0, len(b)), rng.choice(al))
                    elif b:
                        b[rng.randrange(len(b))] = rng.choice(al)
                b = bytes(b)
            pairs.append((a, b))
    got = gpu_ctx.edit_distance(pairs)
    for (a, b), g in zip(pairs, got.tolist()):
        # Python's .upper() is applied by the reference before edlib (SVIM_clustering.py:37-43); the kernel folds case
        want = editdist.edit_distance(a.upper(), b.upper())
        assert g == want, (len(a), len(b), g, want)


@pytest.mark.parametrize("band", ["174,24", "60,0", "420,100", "0"])
def test_banded_edit_distance_matches_oracle(band, monkeypatch):
    """The banded first pass (rotating-lane wavefront, myers_band.cuh) plus the unbanded hand-over must be exact for every
    band policy: a tight bound sends most pairs to the second pass, a wide one keeps divergent pairs on the banded kernels."""
    from oracle import editdist
    from svim_b200 import _lib
    monkeypatch.setenv("SVIM_MYERS_BAND", band)
    ctx = _lib.Context(device=0)
    rng = random.Random(11)

    def mutate(s, rate):
        out = bytearray()
        for ch in s:
            r = rng.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out.append(rng.choice(b"ACGT")); continue
            out.append(ch)
            if r < rate:
                out.append(rng.choice(b"ACGT"))
        return bytes(out)
    pairs = []
    for n in [40, 130, 257, 300, 511, 700, 1000, 1025, 1500, 2000, 2600, 3100, 4000, 4100, 5000, 6200, 7000, 9000, 12500, 17000]:
        for rate in [0.0, 0.03, 0.12, 0.2, 0.35, 0.7]:
            a = bytes(rng.choice(b"ACGT") for _ in range(n))
            b = mutate(a, rate)
            if rng.random() < 0.25:
                b = b[rng.randint(0, len(b) // 4):]                      # length difference: the band is off-centre
            if rng.random() < 0.15:
                a = a[:n // 2] + b"N" + a[n // 2:]                       # a symbol outside A/C/G/T leaves the banded path
            if rng.random() < 0.1:
                a = a.lower()
            pairs.append((a, b if b else b"A"))
    rng.shuffle(pairs)
    got = ctx.edit_distance(pairs)
    for (a, b), g in zip(pairs, got.tolist()):
        want = editdist.edit_distance(a.upper(), b.upper())
        assert g == want, (band, len(a), len(b), g, want)


def test_linkage_kernel_matches_scipy(gpu_ctx):
    from scipy.cluster.hierarchy import linkage, fcluster
    rng = np.random.default_rng(5)
    for t in range(120):
        m = int(rng.integers(2, 101)); npair = m * (m - 1) // 2
        d = [rng.random(npair), rng.integers(0, 3, npair) / 2.0, np.where(rng.random(npair) < 0.3, 99999.0, rng.random(npair)),
             np.round(rng.random(npair), 1)][t % 4]
        d = np.asarray(d, dtype=np.float64)
        Z, T = gpu_ctx.linkage_average(d, m, 0.5)
        Zs = linkage(d, method="average")
        assert np.array_equal(Z, Zs), t                      # bit-exact dendrogram, ties included
        assert T.tolist() == list(fcluster(Zs, 0.5, criterion="distance")), t


def test_oracle_parity_on_fresh_synthetic_inputs(gpu_ctx):
    """Inputs generated at test time (not committed): GPU vs oracle on each BASELINE config in miniature."""
    from svim_b200 import synth
    from oracle import svim_oracle as orc
    from test_oracle_golden import oracle_outputs
    for name, scale in (("config1", 0.3), ("config2", 0.0006), ("config3", 0.002), ("config4", 0.00025), ("config5", 0.0004)):
        batch, genome, _ = synth.make_config(name, scale)
        want = oracle_outputs(batch, genome, {})
        rows, trows, clusters, st, cst = run_gpu(gpu_ctx, batch, genome, {})
        assert rows == want["signatures"], name
        assert_clusters_equal(clusters, want["clusters"])


def test_querysorted_collect_matches_reference_golden(gpu_ctx, golden):
    # analyze_alignment_file_querysorted (SVIM_COLLECT.py:96-129): real supplementary records as segments
    batch, genome, exp = golden("mini_mixed_querysorted")
    rows, trows, clusters, st, cst = run_gpu(gpu_ctx, batch, genome, exp["params"], querysorted=True)
    assert rows == exp["signatures"]
    assert trows == exp["all_bnds_signatures"]
    assert_clusters_equal(clusters, exp["clusters"])
    from svim_b200.SVIM_COLLECT import bam_iterator
    from oracle import svim_oracle as orc
    assert list(bam_iterator(batch)) == orc.read_groups(batch)
