"""CPU-side units: C-ABI library loads and exports what include/svimgpu.h declares, the host
RNG restatement equals CPython's, the pysam-free readers round-trip, the oracle's two edit
distance implementations agree, and the host+device logic headers (tests/hostcheck) agree with
the oracle on fuzzed inputs.  No GPU needed."""
import ctypes
import os
import random
import re
import statistics
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import svim_oracle as orc
from oracle import editdist
from svim_b200 import _lib
from svim_b200.records import BatchBuilder, parse_cigar_string, encode_cigar
from svim_b200 import io as sio


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "svimgpu.h")).read()
    declared = set(re.findall(r"\b(svimgpu_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_lib.EXPORTS) == declared - {"svimgpu_ctx"}
    lib.svimgpu_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.svimgpu_version()


def test_struct_sizes_match_header():
    assert _lib.SIG_DTYPE.itemsize == 48 and _lib.CSIG_DTYPE.itemsize == 64 and _lib.CLUSTER_DTYPE.itemsize == 72
    assert ctypes.sizeof(_lib.Params) == 6 * 4 + 4 * 8


def test_host_rng_matches_cpython_sample():
    # SVIM_clustering.py:129-134: seed(1524) once, sample(partition, 100) per partition > 100
    sizes = [5, 101, 100, 150, 1045, 1046, 2000, 50000, 102, 7, 100000, 1044]
    got = _lib.sample_indices(sizes)
    random.seed(1524)
    want = [random.sample(range(n), 100) for n in sizes if n > 100]
    assert got.tolist() == want
    # sampling a list of objects picks the same positions as sampling range(n)
    random.seed(1524)
    pop = [object() for _ in range(333)]
    pos = {id(o): i for i, o in enumerate(pop)}
    assert [pos[id(o)] for o in random.sample(pop, 100)] == _lib.sample_indices([333])[0].tolist()


def test_edit_distance_oracles_agree():
    rng = random.Random(5)
    for _ in range(200):
        a = "".join(rng.choice("ACGTN") for _ in range(rng.randint(0, 400)))
        b = list(a) if rng.random() < 0.6 else [rng.choice("ACGT") for _ in range(rng.randint(0, 400))]
        for _k in range(rng.randint(0, 40)):
            if b and rng.random() < 0.5:
                del b[rng.randrange(len(b))]
            else:
                b.insert(rng.randint(0, len(b)), rng.choice("ACGT"))
        b = "".join(b)
        assert editdist.edit_distance(a, b) == editdist.edit_distance_dp(a, b)
    assert editdist.edit_distance("kitten", "sitting") == 3


def test_sam_bam_roundtrip(tmp_path):
    b = BatchBuilder(["chr1", "chr2"], [1000, 2000])
    b.add("r1", 0, 0, 10, 60, "5S10M2I3D20M", "ACGTN" * 7 + "AC", "chr2,5,-,10S27M,60,1;")
    b.add("r2", 16, 1, 99, 3, "30M", None)
    b.add("r1", 2048, 1, 4, 60, "10H27M", "A" * 27, "chr1,11,+,5S32M3D,60,2;")
    batch = b.finish()
    sam = str(tmp_path / "x.sam"); bam = str(tmp_path / "x.bam")
    sio.write_sam(sam, batch); sio.write_bam(bam, batch)
    for got in (sio.read_sam(sam), sio.read_bam(bam), sio.read_alignments(bam)):
        assert got.n == 3 and got.contig_names == ["chr1", "chr2"]
        for i in range(3):
            assert got.cigartuples(i) == batch.cigartuples(i)
            assert got.sequence(i) == batch.sequence(i)
            assert got.sa_tag(i) == batch.sa_tag(i)
            assert (got.tid[i], got.pos[i], got.flag[i], got.mapq[i]) == (batch.tid[i], batch.pos[i], batch.flag[i], batch.mapq[i])
        assert got.qname_id[0] == got.qname_id[2] != got.qname_id[1]


def test_native_bam_reader_equals_python_reader(tmp_path):
    """csrc_host/bamio.cpp (parallel inflate + SoA fill) against the pure-Python decoder on a synthetic BAM."""
    from svim_b200 import synth
    names, L = ["chr1", "chr2"], [80_000, 60_000]
    svs, al = synth.plant_svs(L, 3, spacing=6000)
    batch = synth.generate(names, L, 300, 3, svs, al, len_mean=3000, len_sd=600, len_min=800, len_max=6000)
    p = str(tmp_path / "s.bam")
    sio.write_bam(p, batch)
    p2 = str(tmp_path / "n.bam")
    sio.write_bam_native(p2, batch, threads=3)
    assert open(p, "rb").read() == open(p2, "rb").read()           # the native writer emits the same bytes as the Python one
    a, b = sio.read_bam_native(p, threads=4), sio.read_bam_python(p)
    assert a.n == b.n == batch.n and a.contig_names == b.contig_names == names and a.sort_order == b.sort_order == "coordinate"
    for name, _ in a.FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for blob in ("cigar", "seq", "sa"):
        assert np.array_equal(getattr(a, blob), getattr(b, blob)) and np.array_equal(getattr(a, blob), getattr(batch, blob))
    assert [a.qname(int(i)) for i in a.qname_id] == [batch.qname(int(i)) for i in batch.qname_id]


def test_fasta_fetch_clamps(tmp_path):
    g = sio.Genome(["c1", "c2"], [np.frombuffer(b"ACGTacgtNN", dtype=np.uint8), np.frombuffer(b"GG", dtype=np.uint8)])
    p = str(tmp_path / "g.fa"); g.write_fasta(p, width=4)
    g2 = sio.Genome.from_fasta(p)
    assert g2.fetch("c1", 2, 6) == "GTac" and g2.fetch("c1", 8, 50) == "NN" and g2.fetch("c1", 20, 30) == "" and g2.fetch("c2", 0, 2) == "GG"


# ---- host build of the SVIM_HD headers -----------------------------------------------------------
@pytest.fixture(scope="module")
def hc():
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    so = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    deps = [src] + [os.path.join(ROOT, "svim_b200", "csrc", f) for f in ("common.cuh", "collect.cuh", "cluster.cuh", "myers_band.cuh", "myers_tpp.cuh")] + \
           [os.path.join(ROOT, "svim_b200", "csrc", "bgzf_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.hc_spd.restype = ctypes.c_double
    lib.hc_stdev.restype = ctypes.c_double
    lib.hc_score.restype = ctypes.c_double
    lib.hc_round.restype = ctypes.c_longlong
    lib.hc_round.argtypes = [ctypes.c_double]
    lib.hc_score.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    return lib


def _random_read(rng, names):
    """A primary + SA segments with coordinates close enough to reach every branch of the split-read tree."""
    L = rng.randint(2000, 9000)
    k = rng.randint(1, 4)
    cuts = sorted(rng.sample(range(100, L - 100), k))
    bounds = [0] + cuts + [L]
    base = rng.randint(50_000, 60_000)
    segs = []
    for i in range(k + 1):
        qs, qe = bounds[i], bounds[i + 1]
        qs2 = max(0, qs + rng.choice([0, 0, 0, -3, -8, 4, 12, 60]))
        tid = rng.choice([0, 0, 0, 1, 2])
        rev = rng.random() < 0.35
        mode = rng.random()
        if mode < 0.5:
            ref = base + qs + rng.choice([0, 0, 50, -50, 300, -300, 5000, -5000, 150000, -150000, 3, -3, 45, -45])
        else:
            ref = rng.randint(1000, 300_000)
        qlen = max(1, qe - qs2)
        rlen = max(1, qlen + rng.choice([0, 0, 0, -20, 20, 100]))
        segs.append(dict(qs=qs2, qe=qe, tid=tid, rev=rev, ref=max(0, ref), qlen=qlen, rlen=rlen, mapq=rng.choice([60, 60, 60, 5, 300])))
    prim = max(range(len(segs)), key=lambda i: segs[i]["qlen"])

    def cigar(s, clip):
        lead, trail = (L - s["qe"], s["qs"]) if s["rev"] else (s["qs"], L - s["qe"])
        mid = "%dM" % min(s["qlen"], s["rlen"])
        if s["qlen"] > s["rlen"]:
            mid += "%dI" % (s["qlen"] - s["rlen"])
        elif s["rlen"] > s["qlen"]:
            mid += "%dD" % (s["rlen"] - s["qlen"])
        return ("%d%s" % (lead, clip) if lead else "") + mid + ("%d%s" % (trail, clip) if trail else "")
    sa = ""
    for i, s in enumerate(segs):
        if i == prim:
            continue
        sa += "%s,%d,%s,%s,%d,%d;" % (names[s["tid"]], s["ref"] + 1, "-" if s["rev"] else "+", cigar(s, "S"), s["mapq"], 3)
    if rng.random() < 0.05:
        sa += "bad,field;"
    p = segs[prim]
    seq = "".join(rng.choice("ACGT") for _ in range(L)) if rng.random() < 0.9 else None
    return dict(flag=16 if p["rev"] else 0, tid=p["tid"], pos=p["ref"], cigar=cigar(p, "S"), seq=seq, sa=sa)


@pytest.mark.parametrize("all_bnds", [False, True])
def test_segment_chain_logic_matches_oracle(hc, all_bnds):
    rng = random.Random(11 + all_bnds)
    names = ["chr1", "chr10", "chr2"]
    rank = np.array([0, 1, 2], dtype=np.int32)        # string order: chr1 < chr10 < chr2
    blob = "".join(names).encode(); off = np.array([0, 4, 9, 13], dtype=np.int32)
    params = orc.Params(all_bnds=all_bnds, max_sv_size=rng.choice([100000, 2000]))
    cap = 64
    out_m = np.zeros(cap, dtype=_lib.SIG_DTYPE); out_t = np.zeros(cap, dtype=_lib.SIG_DTYPE)
    n_checked = 0
    for it in range(3000):
        r = _random_read(rng, names)
        b = BatchBuilder(names, [400_000] * 3)
        b.add("read", r["flag"], r["tid"], r["pos"], 60, r["cigar"], r["seq"], r["sa"])
        batch = b.finish()
        prim, hard = orc.segment_of_record(batch, 0)
        good = [s for s in orc.sa_segments(batch, 0, hard) if s.mapq >= params.min_mapq]
        want, want_t = orc.segment_signatures(batch, prim, good, batch.sequence(0), "read", params)
        nm, nt, err = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_uint32()
        cig = batch.cigar
        sa = batch.sa
        hc.hc_chain(int(batch.tid[0]), int(batch.pos[0]), int(batch.flag[0]), ctypes.c_int64(int(batch.l_seq[0])),
                    cig.ctypes.data_as(ctypes.c_void_p), int(batch.n_cigar[0]), sa.ctypes.data_as(ctypes.c_void_p), int(batch.sa_len[0]),
                    3, blob, off.ctypes.data_as(ctypes.c_void_p), rank.ctypes.data_as(ctypes.c_void_p),
                    ctypes.c_int64(params.min_sv_size), ctypes.c_int64(params.max_sv_size), ctypes.c_int64(params.segment_gap_tolerance),
                    ctypes.c_int64(params.segment_overlap_tolerance), params.min_mapq, int(all_bnds),
                    out_m.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nm), out_t.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nt), cap,
                    ctypes.byref(err))
        from svim_b200.SVIM_COLLECT import materialize_signatures
        seq = batch.sequence(0) or ""

        def conv(arr, n):
            res = []
            for s in arr[:n]:
                t = _lib.TYPE_NAMES[s["type"]]
                fl = int(s["flags"])
                tup = [t, names[s["contig1"]], int(s["start"]), int(s["end"])]
                if t == "INS":
                    tup.append(seq[int(s["seq_off"]):int(s["seq_off"]) + int(s["seq_len"])])
                elif t == "INV":
                    tup.append(_lib.INV_DIRECTIONS[(fl >> 4) & 7])
                elif t == "DUP_TAN":
                    tup += [int(s["copies"]), bool(fl & 2)]
                elif t == "DUP_INT":
                    tup += [names[s["contig2"]], int(s["pos"])]
                elif t == "BND":
                    tup += [names[s["contig2"]], int(s["pos"]), "rev" if fl & 4 else "fwd", "rev" if fl & 8 else "fwd"]
                res.append(tuple(tup))
            return res

        def conv_o(lst):
            res = []
            for s in lst:
                tup = [s.type, s.contig, s.start, s.end]
                if s.type == "INS":
                    tup.append(s.sequence)
                elif s.type == "INV":
                    tup.append(s.direction)
                elif s.type == "DUP_TAN":
                    tup += [s.copies, s.fully_covered]
                elif s.type == "DUP_INT":
                    tup += [s.contig2, s.pos]
                elif s.type == "BND":
                    tup += [s.contig2, s.pos, s.dir1, s.dir2]
                res.append(tuple(tup))
            return res
        assert conv(out_m, nm.value) == conv_o(want), (it, r)
        assert conv(out_t, nt.value) == conv_o(want_t), (it, r)
        n_checked += len(want) + len(want_t)
    assert n_checked > 1500      # the fuzz really produces signatures of every kind


def test_distance_and_consolidation_scalars(hc):
    rng = random.Random(2)
    for _ in range(2000):
        a = orc.Sig("DEL", "c", rng.randint(0, 10**6), 0, "cigar", "r1"); a.end = a.start + rng.randint(40, 5000)
        b = orc.Sig("DEL", "c", a.start + rng.randint(-900, 900), 0, "cigar", "r2"); b.end = b.start + rng.randint(40, 5000)
        err = ctypes.c_int(0)
        av = (ctypes.c_double * 3)(a.start, a.end, 0); bv = (ctypes.c_double * 3)(b.start, b.end, 0)
        got = hc.hc_spd(0, av, bv, 0, 0, ctypes.c_double(900), ctypes.c_double(1.0), ctypes.c_double(0.5), ctypes.c_double(0), ctypes.byref(err))
        assert got == orc.span_position_distance(a, b, None, orc.Params())
        # INS with a given edit distance (gate open and closed)
        a.type = b.type = "INS"
        ed = rng.randint(0, 3000)
        pd = abs(a.start - b.start) / 900
        s1, s2 = a.end - a.start, b.end - b.start
        want = pd + abs(s1 - s2) / max(s1, s2) if pd > 1.0 else pd + ed / max(s1, s2) / 1.0
        got = hc.hc_spd(1, av, bv, 0, 0, ctypes.c_double(900), ctypes.c_double(1.0), ctypes.c_double(0.5), ctypes.c_double(ed), ctypes.byref(err))
        assert got == want
    for _ in range(500):
        n = rng.randint(2, 100)
        base = rng.randint(0, 2 * 10**8)
        vals = [base + rng.randint(0, 3000) for _ in range(n)]
        if rng.random() < 0.5:
            vals = [(v + rng.randint(0, 3000) + v) / 2 for v in vals]       # half-integral positions
        arr = (ctypes.c_double * n)(*vals)
        got = hc.hc_stdev(arr, n); want = statistics.stdev(vals)
        assert got == want          # correctly rounded, like statistics.stdev (exact rational -> one rounding)
    for x in [0.5, 1.5, 2.5, -0.5, 3.49999, 1e9 + 0.5, 7.0]:
        assert hc.hc_round(x) == int(round(x))
    for _ in range(300):
        L = rng.randint(0, 50); a = rng.randint(-70, 70); n = rng.randint(0, 80)
        lo, hi = ctypes.c_longlong(), ctypes.c_longlong()
        hc.hc_py_slice(ctypes.c_longlong(a), ctypes.c_longlong(n), ctypes.c_longlong(L), ctypes.byref(lo), ctypes.byref(hi))
        s = "x" * L
        assert len(s[a:a + n]) == hi.value - lo.value and (hi.value == lo.value or s[a:a + n] == s[lo.value:hi.value])


def test_fcluster_postprocessing_matches_scipy(hc):
    from scipy.cluster.hierarchy import linkage, fcluster
    rng = np.random.default_rng(9)
    for t in range(300):
        m = int(rng.integers(2, 101)); npair = m * (m - 1) // 2
        d = [rng.random(npair), rng.integers(0, 3, npair) / 2.0, np.where(rng.random(npair) < 0.3, 99999.0, rng.random(npair))][t % 3]
        # unsorted nn-chain rows from the restated algorithm, then the C post-processing
        D = [float(x) for x in d]
        rows = []
        size = [1] * m
        chain = []

        def ix(i, j):
            i, j = (i, j) if i < j else (j, i)
            return m * i - i * (i + 1) // 2 + (j - i - 1)
        for _k in range(m - 1):
            if not chain:
                chain.append(next(i for i in range(m) if size[i] > 0))
            while True:
                x = chain[-1]
                if len(chain) > 1:
                    y = chain[-2]; cur = D[ix(x, y)]
                else:
                    y = -1; cur = float("inf")
                for i in range(m):
                    if size[i] and i != x and D[ix(x, i)] < cur:
                        cur = D[ix(x, i)]; y = i
                if len(chain) > 1 and y == chain[-2]:
                    break
                chain.append(y)
            chain.pop(); chain.pop()
            if x > y:
                x, y = y, x
            nx, ny = size[x], size[y]
            rows.append((x, y, cur)); size[x] = 0; size[y] = nx + ny
            for i in range(m):
                if size[i] and i != y:
                    D[ix(i, y)] = (nx * D[ix(i, x)] + ny * D[ix(i, y)]) / (nx + ny)
        ux = np.array([r[0] for r in rows], dtype=np.int32); uy = np.array([r[1] for r in rows], dtype=np.int32)
        ud = np.array([r[2] for r in rows], dtype=np.float64)
        T = np.zeros(m, dtype=np.int32)
        hc.hc_fcluster(m, ux.ctypes.data_as(ctypes.c_void_p), uy.ctypes.data_as(ctypes.c_void_p), ud.ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_double(0.5), T.ctypes.data_as(ctypes.c_void_p))
        want = fcluster(linkage(np.asarray(d, dtype=float), method="average"), 0.5, criterion="distance")
        assert T.tolist() == list(want)


def test_banded_wavefront_host_replay(hc):
    """myers_band.cuh replayed lane by lane on the host (the kernel k_myers_band runs the same functions, one lane per
    thread): for every shape the band fits,  result <= k  =>  result == distance,  result > k  =>  distance > k."""
    hc.hc_myers_banded.restype = ctypes.c_longlong
    hc.hc_myers_banded.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    hc.hc_myers_band_k.restype = ctypes.c_longlong
    hc.hc_myers_band_k.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    hc.hc_myers_band_bin.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    rng = random.Random(7)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}

    def enc(s):
        return bytes(code[c] for c in s)

    def mutate(s, rate):
        out = []
        for ch in s:
            r = rng.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out.append(rng.choice("ACGT")); continue
            out.append(ch)
            if r < rate:
                out.append(rng.choice("ACGT"))
        return "".join(out)
    exact = bounded = 0
    for it in range(120):
        L = rng.choice([1, 2, 30, 63, 64, 65, 128, 200, 257, 500, 700, 1000, 1500, 2200, 3000])
        a = "".join(rng.choice("ACGT") for _ in range(L))
        b = mutate(a, rng.choice([0, 0.02, 0.1, 0.2, 0.4, 0.8])) or "A"
        if rng.random() < 0.2:
            b = b[rng.randint(0, len(b) // 3):] or "A"
        pat, txt = (a, b) if len(a) >= len(b) else (b, a)
        m, n = len(pat), len(txt)
        d = editdist.edit_distance(pat, txt)
        for k in sorted({m - n, max(m - n, d - 1), max(m - n, d), d + 1, d + 9, max(m - n, d // 2), 2 * d + 3, m + n}):
            for shape in range(10):
                r = hc.hc_myers_banded(enc(pat), m, enc(txt), n, k, shape, it & 1)
                if r == -1:
                    continue            # the band does not fit this shape's rotating window
                assert r >= 0
                if r <= k:
                    assert r == d, (m, n, k, shape, r, d); exact += 1
                else:
                    assert d > k, (m, n, k, shape, r, d); bounded += 1
    assert exact > 1000 and bounded > 500
    # the policy: k = m*num/1024 + add, no banded pass when k < m - n or when no smaller shape holds the band
    assert hc.hc_myers_band_k(1000, 1000, 174, 24) == 193 and hc.hc_myers_band_k(1000, 500, 174, 24) == -1 and hc.hc_myers_band_k(1000, 1000, 0, 24) == -1
    assert hc.hc_myers_band_bin(4000, 3990, 174, 24) == 3 and hc.hc_myers_band_bin(200, 200, 174, 24) == -1


def test_thread_per_pair_band_host_replay(hc):
    """myers_tpp.cuh: the per-thread program of k_myers_tpp (sliding window of B 32-row blocks, D0-form block step, phase-shifted
    columns, virtual rows above and below the pattern) run on the host:  unbanded => exact;  banded: result <= k => exact,
    result > k => distance > k; in the smallest bucket and in wider ones."""
    hc.hc_myers_tpp.restype = ctypes.c_longlong
    hc.hc_myers_tpp.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int]
    hc.hc_tpp_plan.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    rng = random.Random(11)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}

    def enc(s):
        return bytes(code[c] for c in s)

    def mutate(s, rate):
        out = []
        for ch in s:
            r = rng.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out.append(rng.choice("ACGT")); continue
            out.append(ch)
            if r < rate:
                out.append(rng.choice("ACGT"))
        return "".join(out)
    exact = bounded = unbanded = 0
    for it in range(160):
        L = rng.choice([1, 2, 30, 31, 32, 33, 63, 64, 65, 128, 200, 257, 500, 700, 895, 896, 897, 1000, 1500, 2200, 3000])
        a = "".join(rng.choice("ACGT") for _ in range(L))
        b = mutate(a, rng.choice([0, 0.02, 0.1, 0.2, 0.4, 0.8])) or "A"
        if rng.random() < 0.2:
            b = b[rng.randint(0, len(b) // 3):] or "A"
        if rng.random() < 0.15:
            b = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40))) + b     # a shifted diagonal
        pat, txt = (a, b) if len(a) >= len(b) else (b, a)
        m, n = len(pat), len(txt)
        d = editdist.edit_distance(pat, txt)
        r = hc.hc_myers_tpp(enc(pat), m, enc(txt), n, -1, 0)
        if r != -1:
            assert r == d, ("unbanded", m, n, r, d); unbanded += 1
            if m <= 32 * 20:
                assert hc.hc_myers_tpp(enc(pat), m, enc(txt), n, -1, 28) == d
        for k in sorted({m - n, max(m - n, d - 1), max(m - n, d), d + 1, d + 9, max(m - n, d // 2), 2 * d + 3, max(m - n, (m * 156 >> 10) + 20)}):
            for force in (0, 12, 28):
                r = hc.hc_myers_tpp(enc(pat), m, enc(txt), n, k, force)
                if r == -1:
                    continue            # no bucket holds the band
                assert r >= 0
                if r <= k:
                    assert r == d, (m, n, k, force, r, d); exact += 1
                else:
                    assert d > k, (m, n, k, force, r, d); bounded += 1
    assert exact > 1000 and bounded > 300 and unbanded > 60
    a_out = ctypes.c_int(0)
    assert hc.hc_tpp_plan(1000, 1000, 156, 20, ctypes.byref(a_out)) == 7 and a_out.value == 86       # k = 172: 7 blocks instead of 32
    assert hc.hc_tpp_plan(60, 55, 156, 20, ctypes.byref(a_out)) == 2 and a_out.value == -1           # whole pattern in 2 blocks: unbanded
    assert hc.hc_tpp_plan(1000, 500, 156, 20, ctypes.byref(a_out)) == 32 and a_out.value == -1        # k < m - n: no band


def test_candidate_arrays_from_clusters_matches_naive():
    # svim_b200.SVIM_genotyping.candidate_arrays_from_clusters: vectorised marshalling used by bench.py's genotype leg
    from svim_b200 import _lib
    from svim_b200.SVIM_genotyping import candidate_arrays_from_clusters
    rng = np.random.default_rng(5)
    n_sig = 400
    sigs = np.zeros(n_sig, dtype=_lib.SIG_DTYPE)
    sigs["qname_id"] = rng.integers(0, 60, n_sig); sigs["contig1"] = rng.integers(0, 3, n_sig)
    sizes = rng.integers(1, 9, 70)
    members = rng.permutation(n_sig)[: int(sizes.sum())].astype(np.uint32)
    cl = np.zeros(len(sizes), dtype=_lib.CLUSTER_DTYPE)
    cl["size"] = sizes; cl["member_off"] = np.cumsum(sizes) - sizes
    cl["type"] = rng.integers(0, 3, len(sizes)); cl["score"] = rng.integers(0, 8, len(sizes))
    cl["start"] = rng.integers(-5, 1000, len(sizes)); cl["end"] = cl["start"] + rng.integers(0, 500, len(sizes))
    for tcode in (0, 1):
        cands, ids, sel = candidate_arrays_from_clusters(cl, members, sigs, tcode, minimum_score=3)
        want_sel = [k for k in range(len(cl)) if cl["type"][k] == tcode and cl["score"][k] >= 3]
        assert sel.tolist() == want_sel
        off = 0
        for c, k in zip(cands, want_sel):
            m = members[int(cl["member_off"][k]):int(cl["member_off"][k]) + int(cl["size"][k])]
            want_ids = sorted(set(sigs["qname_id"][m].tolist()))
            assert (int(c["start"]), int(c["end"]), int(c["tid"])) == (max(0, int(cl["start"][k])), int(cl["end"][k]), int(sigs["contig1"][m[0]]))
            assert int(c["variant_off"]) == off and int(c["n_variant_reads"]) == len(want_ids)
            assert ids[off:off + len(want_ids)].tolist() == want_ids
            off += len(want_ids)
        assert off == len(ids)


def test_native_bam_reader_records_spanning_units(tmp_path):
    """The streaming decoder cuts the inflated stream into ~1 MiB units; records cut by a unit boundary travel through the
    chain's carry buffer — including records longer than the 256 KiB headroom and longer than a whole unit."""
    from svim_b200.records import BatchBuilder
    rng = np.random.default_rng(9)
    b = BatchBuilder(["c1", "c2"], [50_000_000, 1000], "coordinate")
    pos = 0
    for k in range(60):
        n = int(rng.choice([50, 3000, 70_000, 400_000, 1_600_000]) if k % 3 else 20_000)
        pos += int(rng.integers(1, 1000))
        ops = []
        left = n
        while left > 0:
            ln = int(min(left, rng.integers(1, 40 + n // 500))); ops.append((int(rng.choice([0, 0, 0, 1, 7, 8])), ln)); left -= ln
            if rng.random() < 0.3:
                ops.append((2, int(rng.integers(1, 9))))
        seq = "".join(rng.choice(list("ACGTN"), size=n))
        sa = "c1,%d,+,%dM,60,0;" % (k + 1, n) if k % 4 == 0 else None
        b.add("q%d" % (k % 45), 0 if k % 5 else 0x800, 0, pos, 60, ops, seq, sa)
    b.add("tail", 4, -1, -1, 0, "", None)
    batch = b.finish()
    p = str(tmp_path / "big.bam")
    sio.write_bam_native(p, batch, threads=4)
    for threads in (1, 3, 8):
        a = sio.read_bam_native(p, threads=threads)
        assert a.n == batch.n
        for name, _ in a.FIELDS:
            assert np.array_equal(getattr(a, name), getattr(batch, name)), (threads, name)
        for blob in ("cigar", "seq", "sa"):
            assert np.array_equal(getattr(a, blob), getattr(batch, blob)), (threads, blob)
        assert [a.qname(int(i)) for i in a.qname_id] == [batch.qname(int(i)) for i in batch.qname_id]
        # the decode can emit the 8-bit packed CIGAR stream itself: same bytes as the stand-alone packer
        a8 = sio.read_bam_native(p, threads=threads, pack_cigar=8)
        want8, want_off = sio.pack_cigar8(batch, 2)
        assert np.array_equal(a8.cigar8_off, want_off) and np.array_equal(a8.cigar8, want8), threads
        assert np.array_equal(a8.cigar, batch.cigar)
    # truncated file: the reader reports it instead of returning short data
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(ValueError):
        sio.read_bam_native(p, threads=2)


def test_genotype_marshalling_of_candidate_objects():
    # svim_b200.SVIM_genotyping.genotype_arrays: locus per type, distinct member read NAMES -> sorted qname ids (+ sentinels for
    # names no record carries: they still count in alt_reads, SVIM_genotyping.py:51,93)
    from svim_b200.records import BatchBuilder
    from svim_b200.SVIM_genotyping import genotype_arrays
    b = BatchBuilder(["c1", "c2"], [1000, 2000])
    for k, nm in enumerate(["zed", "amy", "bob", "amy"]):
        b.add(nm, 0, 0, 10 * k, 60, "50M", None)
    batch = b.finish()          # ids by first appearance: zed 0, amy 1, bob 2

    class M:
        def __init__(self, r): self.read = r

    class Cand:
        def __init__(self, src, dst, reads): self.src, self.dst, self.members = src, dst, [M(r) for r in reads]
        def get_source(self): return self.src
        def get_destination(self): return self.dst

    cands = [Cand(("c1", 5, 9), ("c2", 100, 104), ["bob", "zed", "bob"]), Cand(("c2", 7, 8), ("c1", 1, 2), ["ghost", "amy"]), Cand(("c1", 0, 1), ("c1", 0, 1), [])]
    arr, ids = genotype_arrays(cands, batch, "DEL")
    assert arr["tid"].tolist() == [0, 1, 0] and arr["start"].tolist() == [5, 7, 0] and arr["end"].tolist() == [9, 8, 1]
    assert arr["n_variant_reads"].tolist() == [2, 2, 0] and arr["variant_off"].tolist() == [0, 2, 4]
    assert ids.tolist() == [0, 2, 1, 0xFFFFFFFF]
    arr, ids = genotype_arrays(cands, batch, "INS")
    assert arr["tid"].tolist() == [1, 0, 0] and arr["start"].tolist() == [100, 1, 0]
    with pytest.raises(KeyError):
        genotype_arrays([Cand(("nope", 1, 2), ("nope", 1, 2), [])], batch, "DEL")


def test_genotype_host_mirror_against_golden_with_oracle_backed_context(golden, monkeypatch):
    """Host logic of svim_b200.SVIM_genotyping.genotype (score filter, marshalling, result/exception mapping) without a GPU: the
    context's `genotype` entry is replaced by one that answers from the oracle in the C ABI's array format, and the outcome must
    equal the reference's golden outputs.  (The CUDA entry itself is checked in tests/test_gpu_genotype.py.)"""
    import types
    from svim_b200 import _lib, runtime
    from svim_b200.SVIM_genotyping import genotype
    from oracle import svim_oracle as orc
    batch, genome, exp = golden("geno_mini_mixed")
    ends = orc.record_reference_ends(batch)
    calls = {"upload": 0}

    class FakeCtx:
        resident = None

        def upload(self, b):
            calls["upload"] += 1

        def genotype(self, type_code, gp, cands, variant_ids, contig_lengths):
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

    fake = FakeCtx()
    monkeypatch.setattr(runtime, "context", lambda device=None: fake)

    class M:
        def __init__(self, r): self.read = r

    class Cand:
        def __init__(self, contig, start, end, score, reads):
            self.locus, self.score, self.members = (contig, start, end), score, [M(r) for r in reads]
            self.support_fraction, self.genotype, self.ref_reads, self.alt_reads = ".", "./.", None, None
        def get_source(self): return self.locus
        def get_destination(self): return self.locus

    opts = types.SimpleNamespace(min_mapq=20, minimum_score=3, minimum_depth=4, homozygous_threshold=0.8, heterozygous_threshold=0.2)
    for k, v in exp["params"].items():
        setattr(opts, k, v)
    n = 0
    for t in ("DEL", "INV", "INS", "DUP_INT"):
        cands = [Cand(*row[0]) for row in exp["genotype"][t]]
        genotype(cands, batch, t, opts)
        assert [[c.support_fraction, c.genotype, c.ref_reads, c.alt_reads] for c in cands] == [row[1] for row in exp["genotype"][t]], t
        n += len(cands)
    assert n > 30 and calls["upload"] == 1 and fake.resident is batch          # uploaded once, then reused
    with pytest.raises(ValueError):
        genotype([Cand("chr1", 1, 2, 9, [])], batch, "BND", opts)
    with pytest.raises(ValueError):
        genotype([Cand("chr1", 1, 2, 9, [])], batch.take(np.arange(batch.n), "queryname"), "DEL", opts)


def _oracle_backed_context(batch, genome):
    """A stand-in for _lib.Context whose COLLECT / CLUSTER entries answer from the oracle in the C ABI's array formats, so the host
    mirror (marshalling, materialisation, object surface) can be checked against the golden vectors without a GPU."""
    from svim_b200 import _lib
    from oracle import svim_oracle as orc
    tid_of = {n: i for i, n in enumerate(batch.contig_names)}
    qid_of = {batch.qname(i): i for i in set(batch.qname_id.tolist())}
    inv_code = {d: i for i, d in enumerate(_lib.INV_DIRECTIONS)}

    def to_arrays(sigs):
        arr = np.zeros(len(sigs), dtype=_lib.SIG_DTYPE); blob = bytearray()
        for k, s in enumerate(sigs):
            r = arr[k]
            r["type"] = _lib.TYPE_CODE[s.type]; r["contig1"] = tid_of[s.contig]; r["start"] = s.start; r["end"] = s.end
            r["qname_id"] = qid_of[s.read]; r["contig2"] = -1
            fl = 1 if s.signature == "suppl" else 0
            if s.type == "INS":
                r["seq_off"], r["seq_len"] = len(blob), len(s.sequence); blob += s.sequence.encode()
            elif s.type == "INV":
                fl |= inv_code[s.direction] << 4
            elif s.type == "DUP_TAN":
                r["copies"] = s.copies; fl |= 2 if s.fully_covered else 0
            elif s.type == "DUP_INT":
                r["contig2"] = tid_of[s.contig2]; r["pos"] = s.pos
            elif s.type == "BND":
                r["contig2"] = tid_of[s.contig2]; r["pos"] = s.pos; fl |= (4 if s.dir1 == "rev" else 0) | (8 if s.dir2 == "rev" else 0)
            r["flags"] = fl
        return arr, np.frombuffer(bytes(blob), dtype=np.uint8)

    class Ctx:
        resident = None; contigs_key = None; genome_key = None; collect_token = None; collect_batch = None

        def set_params(self, p):
            self.p = orc.Params(min_mapq=p.min_mapq, min_sv_size=p.min_sv_size, max_sv_size=p.max_sv_size, segment_gap_tolerance=p.segment_gap_tolerance,
                                segment_overlap_tolerance=p.segment_overlap_tolerance, all_bnds=bool(p.all_bnds),
                                partition_max_distance=p.partition_max_distance, position_distance_normalizer=p.position_distance_normalizer,
                                edit_distance_normalizer=p.edit_distance_normalizer, cluster_max_distance=p.cluster_max_distance)

        def set_contigs(self, names): pass
        def set_genome(self, g): pass

        def collect_host(self, b):
            self.lists = orc.collect(b, self.p)
            self.arrays = [to_arrays(l) for l in self.lists]
            st = _lib.CollectStats()
            st.n_signatures, st.n_twin_signatures = len(self.lists[0]), len(self.lists[1])
            st.ins_bytes, st.twin_ins_bytes = self.arrays[0][1].size, self.arrays[1][1].size
            return st

        def fetch_signatures(self, which, stats): return self.arrays[which]
        def use_collected(self, which=0): self.which = which

        def cluster(self, sharded=False, view=False):
            lst = self.lists[self.which]
            index_of = {id(s): i for i, s in enumerate(lst)}
            stats = {}
            res = orc.cluster(lst, genome, self.p, stats)
            rows, mem = [], []
            by_type = dict(zip(("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND"), res))
            for t in _lib.TYPE_NAMES:                    # enum order of the ABI: DEL, INS, INV, DUP_TAN, BND, DUP_INT
                for c in by_type[t]:
                    nan = float("nan")
                    rows.append((c.start, c.end, c.dest_start or 0, c.dest_end or 0, c.score, nan if c.std_span is None else c.std_span,
                                 nan if c.std_pos is None else c.std_pos, len(mem), len(c.members), _lib.TYPE_CODE[t],
                                 1 if c.dir1 == "rev" else 0, 1 if c.dir2 == "rev" else 0, 0, 0))
                    mem += [index_of[id(m)] for m in c.members]
            st = _lib.ClusterStats()
            st.n_clusters_total, st.n_members = len(rows), len(mem)
            return st, np.array(rows, dtype=_lib.CLUSTER_DTYPE) if rows else np.zeros(0, _lib.CLUSTER_DTYPE), np.array(mem, dtype=np.uint32)
    return Ctx()


@pytest.mark.parametrize("name", ["mini_mixed", "mini_mixed_allbnds", "mini_ins"])
def test_collect_cluster_host_mirror_against_golden_with_oracle_backed_context(golden, monkeypatch, name):
    """svim_b200.SVIM_COLLECT.analyze_alignment_file_coordsorted + SVIM_CLUSTER.cluster_sv_signatures (the two call sites svim:102 and
    svim:132): records -> SVSignature objects -> cluster objects, compared attribute by attribute with the reference's golden outputs."""
    import types
    from svim_b200 import runtime
    from svim_b200.SVIM_COLLECT import analyze_alignment_file_coordsorted
    from svim_b200.SVIM_CLUSTER import cluster_sv_signatures
    batch, genome, exp = golden(name)
    fake = _oracle_backed_context(batch, genome)
    monkeypatch.setattr(runtime, "context", lambda device=None: fake)
    runtime.register_genome("golden.fa", genome)
    opts = types.SimpleNamespace(genome="golden.fa", **exp["params"])
    sigs, twins = analyze_alignment_file_coordsorted(batch, opts)

    def sig_row(s):
        d = dict(type=s.type, signature=s.signature, read=s.read)
        if s.type in ("DEL", "INS", "INV", "DUP_TAN"):
            d.update(contig=s.contig, start=s.start, end=s.end)
            if s.type == "INS": d["sequence"] = s.sequence
            if s.type == "INV": d["direction"] = s.direction
            if s.type == "DUP_TAN": d.update(copies=s.copies, fully_covered=s.fully_covered)
        elif s.type == "DUP_INT":
            d.update(contig=s.contig1, start=s.start, end=s.end, contig2=s.contig2, pos=s.pos)
        else:
            d.update(contig=s.contig1, start=s.pos1, end=s.pos1 + 1, contig2=s.contig2, pos=s.pos2, dir1=s.direction1, dir2=s.direction2)
        fields = ("type", "contig", "start", "end", "contig2", "pos", "dir1", "dir2", "direction", "copies", "fully_covered", "signature", "read", "sequence")
        return [d.get(f) for f in fields]

    assert [sig_row(s) for s in sigs] == exp["signatures"]
    assert [sig_row(s) for s in twins] == exp["all_bnds_signatures"]
    # object surface used downstream (SVSignature.py): keys, sources, destinations, strings
    for s in sigs[:200]:
        assert s.get_key()[0] == s.type and isinstance(s.as_string(), str) and len(s.get_source()) == 3
    for lst, key in ((sigs, "clusters"), (twins, "all_bnds_clusters")):
        res = cluster_sv_signatures(lst, opts)
        index_of = {id(s): i for i, s in enumerate(lst)}
        for t, cl in zip(("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND"), res):
            got = []
            for c in cl:
                m = [index_of[id(x)] for x in c.members]
                if hasattr(c, "source_contig"):
                    got.append([c.type, c.source_contig, c.source_start, c.source_end, c.dest_contig, c.dest_start, c.dest_end, c.score, c.size,
                                c.std_span, c.std_pos, getattr(c, "direction1", None), getattr(c, "direction2", None), m])
                else:
                    got.append([c.type, c.contig, c.start, c.end, None, None, None, c.score, c.size, c.std_span, c.std_pos, None, None, m])
            assert got == exp[key][t], (key, t)
        assert isinstance(res, tuple) and all(type(x) is list for x in res)      # COMBINE mutates these lists in place


def test_flag_cutpaste_host_mirror_against_golden(monkeypatch):
    # svim_b200.SVIM_merging.flag_cutpaste_candidates: candidate construction + flagging; the distance search answered in numpy
    import gzip, json, types
    from conftest import GOLDEN
    from svim_b200 import runtime
    from svim_b200.SVIM_merging import flag_cutpaste_candidates

    class Ctx:
        def closest_source(self, a_s, a_e, b_s, b_e, N):
            a_s, a_e, b_s, b_e = (np.asarray(x, dtype=np.int64) for x in (a_s, a_e, b_s, b_e))
            idx = np.zeros(len(a_s), np.int64); dist = np.zeros(len(a_s))
            for k in range(len(a_s)):
                d = np.abs((b_s + b_e) // 2 - (a_s[k] + a_e[k]) // 2) / N + np.abs((b_e - b_s) - (a_e[k] - a_s[k])) / np.maximum(b_e - b_s, a_e[k] - a_s[k])
                idx[k] = int(np.argmin(d)); dist[k] = d[idx[k]]
            return idx, dist
    monkeypatch.setattr(runtime, "context", lambda device=None: Ctx())
    g = json.load(gzip.open(os.path.join(GOLDEN, "cutpaste.golden.json.gz"), "rt"))
    opts = types.SimpleNamespace(position_distance_normalizer=g["position_distance_normalizer"], del_ins_dup_max_distance=g["del_ins_dup_max_distance"])

    class Uni:
        def __init__(self, c, s, e): self.src = (c, s, e)
        def get_source(self): return self.src

    class Bi(Uni):
        def __init__(self, c, s, e, dc, dp, k):
            super().__init__(c, s, e); self.dst = (dc, dp, dp + e - s); self.members, self.score, self.std_span, self.std_pos = ["m%d" % k], 4.0, None, 1.5
        def get_destination(self): return self.dst
    for case in g["cases"]:
        got = flag_cutpaste_candidates([Bi(*r, k) for k, r in enumerate(case["inss"])], [Uni(*d) for d in case["dels"]], opts)
        assert [bool(c.cutpaste) for c in got] == case["cutpaste"]
        assert all(c.type == "DUP_INT" and c.score == 4.0 and c.std_pos == 1.5 for c in got)
    assert flag_cutpaste_candidates([], [], opts) == []
    with pytest.raises(IndexError):
        flag_cutpaste_candidates([Bi("c", 1, 5, "c", 9, 0)], [], opts)



def test_bgzf_core_inflate_matches_zlib(hc, tmp_path):
    """csrc/bgzf_core.cuh (host+device core of the on-GPU BAM decoder): the SVIM_HD raw-DEFLATE decoder against zlib on synthetic
    streams of every block type and on every BGZF block of a BAM file; malformed input must come back as an error code."""
    import zlib
    hc.hc_bgzf_inflate.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint]
    rng = np.random.default_rng(4)

    def deflate_raw(data, level, strategy):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        return c.compress(data) + c.flush()

    def inflate(comp, n):
        out = np.full(n + 8, 0xAB, dtype=np.uint8)
        rc = hc.hc_bgzf_inflate(comp, len(comp), out.ctypes.data, n)
        assert (out[n:] == 0xAB).all()                  # nothing written past the declared size
        return rc, out[:n].tobytes()

    n_ok = n_rej = 0
    for t in range(600):
        n = int(rng.integers(0, 9)) if t % 7 == 0 else int(rng.integers(0, 65536))
        mode = t % 6
        if mode == 0: data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif mode == 1: data = rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes()
        elif mode == 2: data = b"\xff" * n
        elif mode == 3: data = bytes((120 if i % 7 < 3 else (i // 300) & 255) for i in range(n))
        elif mode == 4: data = (rng.integers(0, 4, n, dtype=np.uint8) * (rng.random(n) < 0.1)).astype(np.uint8).tobytes()
        else:
            base = rng.integers(0, 64, min(n, 33000), dtype=np.uint8).tobytes(); data = (base * 3)[:n]          # far matches (distance ~ 32 k)
        level = int(rng.integers(0, 10))
        strategy = zlib.Z_FIXED if t % 11 == 0 else zlib.Z_HUFFMAN_ONLY if t % 13 == 0 else zlib.Z_RLE if t % 17 == 0 else zlib.Z_DEFAULT_STRATEGY
        comp = deflate_raw(data, level, strategy)
        rc, got = inflate(comp, len(data))
        assert rc == 0 and got == data, (t, n, level, strategy, rc)
        n_ok += 1
        if n:
            assert inflate(comp, len(data) - 1)[0] != 0 and inflate(comp, len(data) + 3)[0] != 0       # wrong ISIZE
        if len(comp) > 4:
            assert inflate(comp[: len(comp) - 1 - int(rng.integers(0, min(len(comp) - 1, 16)))], len(data))[0] != 0 or n == 0      # truncated
            for _ in range(2):                                                                            # bit flips: never crash
                bad = bytearray(comp); bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
                n_rej += inflate(bytes(bad), len(data))[0] != 0
    assert n_ok == 600 and n_rej > 300
    # every BGZF block of a BAM written by the native writer
    from svim_b200 import synth
    names, L = ["chr1", "chr2"], [90_000, 50_000]
    svs, al = synth.plant_svs(L, 5, spacing=6000)
    batch = synth.generate(names, L, 400, 5, svs, al, len_mean=4000, len_sd=900, len_min=800, len_max=9000)
    p = str(tmp_path / "x.bam")
    sio.write_bam_native(p, batch, threads=2)
    raw = open(p, "rb").read()
    o = 0; stream = bytearray(); n_blocks = 0
    while o + 18 <= len(raw):
        xlen = int.from_bytes(raw[o + 10:o + 12], "little"); bsize = int.from_bytes(raw[o + 16:o + 18], "little") + 1
        isize = int.from_bytes(raw[o + bsize - 4:o + bsize], "little")
        payload = raw[o + 12 + xlen:o + bsize - 8]
        rc, got = inflate(payload, isize)
        assert rc == 0 and got == zlib.decompress(payload, -15), n_blocks
        stream += got; o += bsize; n_blocks += 1
    assert n_blocks > 20
    # record starts: from arbitrary offsets the search must land on the next true record start
    hc.hc_bam_find_record_start.restype = ctypes.c_ulonglong
    hc.hc_bam_find_record_start.argtypes = [ctypes.c_char_p, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int]
    data = bytes(stream)
    l_text = int.from_bytes(data[4:8], "little"); q = 8 + l_text; n_ref = int.from_bytes(data[q:q + 4], "little"); q += 4
    for _ in range(n_ref):
        ln = int.from_bytes(data[q:q + 4], "little"); q += 4 + ln + 4
    starts = []
    while q + 4 <= len(data):
        starts.append(q); q += 4 + int.from_bytes(data[q:q + 4], "little")
    assert len(starts) == batch.n
    starts_a = np.array(starts)
    for frm in [starts[0], starts[0] + 1, starts[5] - 3] + rng.integers(starts[0], starts[-3], 300).tolist():
        want = int(starts_a[np.searchsorted(starts_a, frm)])
        assert hc.hc_bam_find_record_start(data, len(data), frm, n_ref, 3) == want, frm
    # the chunk-parallel boundary scheme (speculative start per chunk, chain, "lands on the next guess" check), several chunk sizes
    hc.hc_bam_chunked_starts.restype = ctypes.c_longlong
    hc.hc_bam_chunked_starts.argtypes = [ctypes.c_char_p, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    for chunk in (1500, 4096, 65536, 1 << 20):
        n_chunks = (len(data) - starts[0] + chunk - 1) // chunk
        got_starts = np.zeros(n_chunks, dtype=np.uint64)
        assert hc.hc_bam_chunked_starts(data, len(data), starts[0], chunk, n_ref, 3, got_starts.ctypes.data) == batch.n, chunk
        want = [int(starts_a[np.searchsorted(starts_a, starts[0] + c * chunk)]) if starts[0] + c * chunk <= starts[-1] else len(data) for c in range(n_chunks)]
        assert got_starts.tolist() == want, chunk
    assert hc.hc_bam_chunked_starts(data[:-7], len(data) - 7, starts[0], 4096, n_ref, 3, None) == -2          # truncated stream


def test_bgzf_layout_for_the_gpu_decoder(tmp_path):
    # svim_b200.io.bgzf_layout: block table + first-record offset handed to svimgpu_decode_bam (csrc/bam.cu), checked with zlib on the CPU
    import zlib
    from svim_b200 import synth
    names, L = ["chr1", "chr2"], [90_000, 50_000]
    svs, al = synth.plant_svs(L, 5, spacing=6000)
    batch = synth.generate(names, L, 300, 5, svs, al, len_mean=4000, len_sd=900, len_min=800, len_max=9000)
    p = str(tmp_path / "x.bam")
    sio.write_bam_native(p, batch, threads=2)
    blocks, first, cn, cl, so = sio.bgzf_layout(p)
    assert cn == names and cl.tolist() == L and so == "coordinate"
    raw = open(p, "rb").read()
    stream = b"".join(zlib.decompress(raw[int(b["coff"]):int(b["coff"]) + int(b["clen"])], -15) for b in blocks)
    assert len(stream) == int(blocks["ulen"].sum()) and (np.cumsum(blocks["ulen"]) - blocks["ulen"] == blocks["uoff"]).all()
    n, q = 0, first
    while q + 4 <= len(stream):
        q += 4 + int.from_bytes(stream[q:q + 4], "little"); n += 1
    assert n == batch.n and q == len(stream)


def test_bam_long_cigar_cg_tag_roundtrip(tmp_path):
    """SAMv1 §4.2.2: more than 65535 CIGAR operations do not fit the record core; the real CIGAR travels in the CG:B,I tag behind the
    placeholder <l_seq>S<rlen>N, and readers put it back (htslib sam.c bam_tag2cigar, so pysam — and the reference — never see the
    placeholder).  Both writers emit it, both readers restore it."""
    from svim_b200.records import BatchBuilder
    rng = np.random.default_rng(3)
    b = BatchBuilder(["c1"], [10_000_000], "coordinate")

    def ops_for(n_ops):
        ops = []
        for k in range(n_ops):
            ops.append((0 if k % 2 == 0 else int(rng.choice([1, 2])), int(rng.integers(1, 4))))
        return ops

    for k, n_ops in enumerate([10, 65535, 65536, 100_001, 3]):
        ops = ops_for(n_ops)
        qlen = sum(n for o, n in ops if o in (0, 1))
        b.add("r%d" % k, 0, 0, 100 + k, 60, ops, "".join(rng.choice(list("ACGT"), size=qlen)), "c1,5,+,%dM,60,0;" % qlen if k % 2 else None)
    b.add("clipped", 0, 0, 900, 60, [(4, 50)], "A" * 50, None)           # 50S with l_seq 50 but no CG tag: stays as it is
    batch = b.finish()
    assert int(batch.n_cigar.max()) == 100_001
    p1, p2 = str(tmp_path / "py.bam"), str(tmp_path / "native.bam")
    sio.write_bam(p1, batch); sio.write_bam_native(p2, batch, threads=2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    for reader in (sio.read_bam_python, lambda p: sio.read_bam_native(p, threads=3)):
        got = reader(p1)
        for name, _ in batch.FIELDS:
            assert np.array_equal(getattr(got, name), getattr(batch, name)), name
        for blob in ("cigar", "seq", "sa"):
            assert np.array_equal(getattr(got, blob), getattr(batch, blob)), blob
    got8 = sio.read_bam_native(p1, threads=3, pack_cigar=8)          # CG:B,I records: the packed stream holds the REAL operations
    want8, want_off = sio.pack_cigar8(batch, 2)
    assert np.array_equal(got8.cigar8_off, want_off) and np.array_equal(got8.cigar8, want8)
    # the record core really carries the placeholder (what a reader without CG support would see)
    raw = sio._bgzf_inflate_all(p1)
    assert raw.count(b"CGBI") == 2


# ---- host mirror: object building in C, CLI writers ------------------------------------------------------------------
def _random_sig_records(n, seed=5):
    from svim_b200 import _lib
    rng = np.random.default_rng(seed)
    sigs = np.zeros(n, dtype=_lib.SIG_DTYPE)
    sigs["type"] = rng.integers(0, 6, n); sigs["start"] = rng.integers(0, 1 << 28, n); sigs["end"] = sigs["start"] + rng.integers(0, 5000, n)
    sigs["pos"] = rng.integers(0, 1 << 28, n); sigs["contig1"] = rng.integers(0, 3, n); sigs["contig2"] = rng.integers(0, 3, n)
    sigs["qname_id"] = rng.integers(0, 500, n); sigs["flags"] = rng.integers(0, 256, n) & 0x4f; sigs["copies"] = rng.integers(1, 5, n)
    lens = np.where(sigs["type"] == 1, rng.integers(0, 300, n), 0)
    sigs["seq_len"] = lens; sigs["seq_off"] = np.concatenate([[0], np.cumsum(lens)])[:-1]
    ins = rng.choice(np.frombuffer(b"ACGTNacgt=", dtype=np.uint8), size=int(lens.sum()))
    return sigs, ins, rng


def _slots_of(o):
    d = {k: getattr(o, k) for c in type(o).__mro__ for k in getattr(c, "__slots__", ()) if k != "__dict__" and hasattr(o, k)}
    d.update(getattr(o, "__dict__", {})); d["type"] = o.type; d["class"] = type(o).__name__
    return d


def test_c_object_builder_equals_python_builder():
    """svim_b200/csrc_host/fastobj.c (what analyze_alignment_file_coordsorted / cluster_sv_signatures return) against the same
    marshalling written in Python: every attribute of every object, the per-type cluster lists and their member identity."""
    from svim_b200 import _lib
    from svim_b200.SVIM_COLLECT import materialize_signatures, materialize_signatures_py
    from svim_b200.SVIM_clustering import build_clusters, build_clusters_py

    class Batch:
        contig_names = ["chr1", "chr10", "chr2"]
        qnames = ["q%d" % i for i in range(500)]

        def qname(self, i):
            return self.qnames[i] if self.qnames is not None else "read%d" % i
    sigs, ins, rng = _random_sig_records(4000)
    for named in (True, False):
        b = Batch()
        if not named:
            b.qnames = None
        got = materialize_signatures(sigs, ins, b); want = materialize_signatures_py(sigs, ins, b)
        assert len(got) == len(want) and all(_slots_of(x) == _slots_of(y) for x, y in zip(got, want))
    got[0].anything = 3                       # downstream code may attach attributes (plain classes in the reference)
    assert got[0].anything == 3 and got[0].get_key()[0] == got[0].type
    nc = 700
    cl = np.zeros(nc, dtype=_lib.CLUSTER_DTYPE)
    sizes = rng.integers(1, 12, nc); off = np.concatenate([[0], np.cumsum(sizes)])
    members = rng.integers(0, len(sigs), int(off[-1])).astype(np.uint32)
    cl["size"] = sizes; cl["member_off"] = off[:-1]; cl["type"] = sigs["type"][members[off[:-1]]]
    cl["start"] = rng.integers(0, 1 << 40, nc); cl["end"] = cl["start"] + 7; cl["dest_start"] = rng.integers(0, 1 << 40, nc); cl["dest_end"] = cl["dest_start"] + 3
    cl["score"] = rng.random(nc); cl["std_span"] = np.where(sizes > 1, rng.random(nc), np.nan); cl["std_pos"] = np.where(sizes > 1, rng.random(nc), np.nan)
    cl["dir1_rev"] = rng.integers(0, 2, nc); cl["dir2_rev"] = rng.integers(0, 2, nc)
    a = build_clusters(cl, members, got); p = build_clusters_py(cl, members, got)
    for t in range(6):
        assert len(a[t]) == len(p[t])
        for x, y in zip(a[t], p[t]):
            dx, dy = _slots_of(x), _slots_of(y)
            assert [id(m) for m in dx.pop("members")] == [id(m) for m in dy.pop("members")] and dx == dy
    assert isinstance(a[0], list) and a[4][0].direction1 in ("fwd", "rev")
    with pytest.raises((ValueError, IndexError)):
        bad = cl.copy(); bad["member_off"][0] = len(members); build_clusters(bad, members, got)


def test_cli_writes_the_reference_file_set(tmp_path):
    """`python -m svim_b200 alignment` writes signatures/{del,ins,inv,dup_tan_source,dup_tan_dest,dup_int,trans}.bed and all.vcf;
    tests/golden/cli_signatures/ holds what the reference's own writers (SVIM_CLUSTER.py:29-107) produce for these clusters."""
    from svim_b200.SVSignature import (SignatureDeletion, SignatureInsertion, SignatureInversion, SignatureDuplicationTandem, SignatureInsertionFrom,
                                       SignatureTranslocation, SignatureClusterUniLocal as U, SignatureClusterBiLocal as Bi)
    from svim_b200.__main__ import write_cluster_beds, write_cluster_vcf
    d1 = SignatureDeletion("chr2", 100, 200, "cigar", "r1"); d2 = SignatureDeletion("chr10", 50, 90, "suppl", "r2")
    i1 = SignatureInsertion("chr1", 10, 60, "cigar", "r3", "ACGT" * 12)
    v1 = SignatureInversion("chr1", 500, 900, "suppl", "r4", "left_fwd")
    t1 = SignatureDuplicationTandem("chr1", 1000, 1500, 2, True, "suppl", "r5")
    f1 = SignatureInsertionFrom("chr1", 100, 300, "chr2", 700, "suppl", "r6")
    b1 = SignatureTranslocation("chr1", 100, "fwd", "chr2", 700, "rev", "suppl", "r7")
    clusters = ([U("chr2", 100, 200, 3.5, 2, [d1, d1], "DEL", 1.5, 2.5), U("chr10", 50, 90, 1.0, 1, [d2], "DEL", None, None)],
                [U("chr1", 10, 60, 1.0, 1, [i1], "INS", None, None)], [U("chr1", 500, 900, 1.0, 1, [v1], "INV", None, None)],
                [Bi("chr1", 1000, 1500, "chr1", 1500, 2500, 2.0, 1, [t1], "DUP_TAN", None, None), Bi("chr1", 10, 15, "chr1", 15, 25, 2.0, 1, [t1], "DUP_TAN", 0.5, 0.25)],
                [Bi("chr1", 100, 300, "chr2", 700, 900, 1.0, 1, [f1], "DUP_INT", None, None)],
                [Bi("chr1", 100, 101, "chr2", 700, 701, 1.0, 1, [b1], "BND", None, None)])
    write_cluster_beds(str(tmp_path), clusters); write_cluster_vcf(str(tmp_path), clusters)
    gold = os.path.join(ROOT, "tests", "golden", "cli_signatures")
    assert sorted(os.listdir(tmp_path / "signatures")) == sorted(os.listdir(gold))
    for f in os.listdir(gold):
        assert (tmp_path / "signatures" / f).read_text() == open(os.path.join(gold, f)).read(), f


# ---- svim_aln_soa.cigar8: host packer, Python inverse, and the device expansion's warp loop replayed on the host ----------------
def _cigar8_cases(rng):
    """Records whose packed streams put extension chains everywhere: inside a lane, across 16-byte lane boundaries, across the
    512-byte round boundary, at the very start and the very end; lengths up to 2^28 - 1; empty records."""
    recs = []
    for _ in range(60):
        n = rng.choice([0, 1, 2, 7, 15, 16, 17, 31, 200, 511, 512, 513, 700, 1500])
        ops = []
        for _k in range(n):
            r = rng.random()
            ln = (rng.randint(1, 15) if r < 0.80 else rng.randint(16, 255) if r < 0.92 else rng.randint(256, 70000) if r < 0.98
                  else rng.randint(1 << 20, (1 << 28) - 1))
            ops.append((ln << 4) | rng.randint(0, 8))
        recs.append(ops)
    recs.append([((1 << 28) - 1) << 4 | 2] * 40)             # every operation needs six extension bytes
    recs.append([(0 << 4) | 0, (16 << 4) | 1, (15 << 4) | 8])  # zero length, first length that needs an extension
    return recs


def test_cigar8_pack_roundtrip_and_device_loop_replay(hc):
    from svim_b200.records import BatchBuilder
    rng = random.Random(808)
    recs = _cigar8_cases(rng)
    b = BatchBuilder(["c"], [1 << 30])
    for i, ops in enumerate(recs):
        b.add(tid=0, pos=10 * i, flag=0, mapq=60, cigar=np.asarray(ops, dtype=np.uint32), seq="", sa="", qname="r%d" % i)
    batch = b.finish().pack_cigar8(3)
    assert batch.cigar8.dtype == np.uint8 and (batch.cigar8_off % 16 == 0).all()
    un = sio.unpack_cigar8(batch.cigar8, batch.cigar8_off, batch.n_cigar)
    hc.hc_expand_cigar8.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_void_p]
    hc.hc_expand_cigar8_staged.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_uint]
    for i, ops in enumerate(recs):
        assert un[i].tolist() == ops
        lo, hi = int(batch.cigar8_off[i]), int(batch.cigar8_off[i + 1])
        src = np.ascontiguousarray(batch.cigar8[lo:hi])
        dst = np.full(len(ops) + 4, 0xDEADBEEF, dtype=np.uint32)
        rc = hc.hc_expand_cigar8(src.ctypes.data if src.size else None, hi - lo, len(ops), dst.ctypes.data)
        assert rc == 0, (i, rc)
        assert dst[:len(ops)].tolist() == ops and (dst[len(ops):] == 0xDEADBEEF).all()
        # the staged variant (k_expand_cigar8_staged): every word written exactly once whatever the destination's line phase
        dst2 = np.full(len(ops) + 4, 0xDEADBEEF, dtype=np.uint32)
        rc = hc.hc_expand_cigar8_staged(src.ctypes.data if src.size else None, hi - lo, len(ops), dst2.ctypes.data, (7 * i) % 32)
        assert rc == 0 and np.array_equal(dst2, dst), (i, rc)
    # a stream that does not hold n_cigar operations, and a dangling extension byte, are reported
    lo, hi = int(batch.cigar8_off[5]), int(batch.cigar8_off[6])
    src = np.ascontiguousarray(batch.cigar8[lo:hi]); n5 = len(recs[5])
    dst = np.zeros(n5 + 8, dtype=np.uint32)
    assert hc.hc_expand_cigar8(src.ctypes.data, hi - lo, n5 + 1, dst.ctypes.data) == 1
    bad = src.copy(); bad[-1] = 0x1F
    assert hc.hc_expand_cigar8(bad.ctypes.data, hi - lo, n5, dst.ctypes.data) == 1
    assert batch.cigar8.size < 2 * batch.cigar.size      # bytes against uint32 words: under half the bytes even on this extension-heavy mix

