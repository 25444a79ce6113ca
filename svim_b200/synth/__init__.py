"""Seeded synthetic inputs for the BASELINE.json configs.

`make_config(name, scale)` -> (AlignmentBatch, Genome, truth dict).  The heavy part
(CIGAR / SEQ / SA synthesis) is native (synth.cpp, OpenMP); SV planting and the
reference genome are numpy.  Bench/test infrastructure, not part of the COLLECT /
CLUSTER product path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from ..records import AlignmentBatch
from ..io import Genome

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsvimsynth.so")
_SRC = os.path.join(_HERE, "synth.cpp")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or (os.path.exists(_SRC) and os.path.getmtime(_SO) < os.path.getmtime(_SRC)):
        from ..build import run_atomic
        run_atomic(["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-o", "@OUT@", _SRC], _SO)
    return _SO


class _Sv(ctypes.Structure):
    _fields_ = [("tid", ctypes.c_int32), ("pos", ctypes.c_int32), ("len", ctypes.c_int32), ("type", ctypes.c_int32),
                ("copies", ctypes.c_int32), ("tid2", ctypes.c_int32), ("pos2", ctypes.c_int32), ("strand2", ctypes.c_int32),
                ("vaf", ctypes.c_float), ("pad", ctypes.c_int32), ("allele_off", ctypes.c_int64)]


SV_DTYPE = np.dtype([("tid", "<i4"), ("pos", "<i4"), ("len", "<i4"), ("type", "<i4"), ("copies", "<i4"),
                     ("tid2", "<i4"), ("pos2", "<i4"), ("strand2", "<i4"), ("vaf", "<f4"), ("pad", "<i4"),
                     ("allele_off", "<i8")])
assert SV_DTYPE.itemsize == ctypes.sizeof(_Sv)


class _Config(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("n_contigs", ctypes.c_int32), ("pad0", ctypes.c_int32),
                ("contig_len", ctypes.c_void_p), ("contig_names", ctypes.c_char_p), ("n_reads", ctypes.c_int64),
                ("len_mean", ctypes.c_double), ("len_sd", ctypes.c_double), ("len_min", ctypes.c_int32),
                ("len_max", ctypes.c_int32), ("p_ins", ctypes.c_double), ("p_del", ctypes.c_double),
                ("geo_ins", ctypes.c_double), ("geo_del", ctypes.c_double), ("p_lowmapq", ctypes.c_double),
                ("p_secondary", ctypes.c_double), ("p_unmapped", ctypes.c_double), ("p_split", ctypes.c_double),
                ("n_sv", ctypes.c_int64), ("svs", ctypes.c_void_p), ("alleles", ctypes.c_void_p)]


class _Sizes(ctypes.Structure):
    _fields_ = [("n_records", ctypes.c_int64), ("cigar_words", ctypes.c_int64), ("seq_bytes", ctypes.c_int64),
                ("sa_bytes", ctypes.c_int64)]


class _Out(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off",
                                                "sa_off", "sa_len", "qname_id", "cigar", "seq", "sa")]


_lib = None


def _get():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.synth_plan.restype = ctypes.c_void_p
        lib.synth_plan.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(_Sizes)]
        lib.synth_fill.restype = None
        lib.synth_fill.argtypes = [ctypes.POINTER(_Config), ctypes.c_void_p, ctypes.POINTER(_Out)]
        lib.synth_free.restype = None
        lib.synth_free.argtypes = [ctypes.c_void_p]
        lib.synth_plan_ncigar.restype = None
        lib.synth_plan_ncigar.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.synth_range_sizes.restype = None
        lib.synth_range_sizes.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(_Sizes)]
        lib.synth_fill_range.restype = None
        lib.synth_fill_range.argtypes = [ctypes.POINTER(_Config), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(_Out)]
        _lib = lib
    return _lib


def random_genome(names, lengths, seed=1524) -> Genome:
    """Uniform ACGT reference (numpy default_rng(1524))."""
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [lut[rng.integers(0, 4, size=int(l), dtype=np.uint8)] for l in lengths]
    return Genome(list(names), seqs)


def plant_svs(lengths, seed, spacing=20000, mix=None, size_range=(50, 5000), ins_size_uniform=None,
              hotspots=0, hotspot_svs=(30, 60)):
    """Planted SV table (sorted by contig, position) + allele blob (ASCII)."""
    rng = np.random.default_rng(seed)
    mix = mix or {"DEL": 0.465, "INS": 0.465, "INV": 0.02, "DUP_TAN": 0.02, "BND": 0.02, "DUP_INT": 0.01}
    tcode = {"DEL": 0, "INS": 1, "INV": 2, "DUP_TAN": 3, "BND": 4, "DUP_INT": 5}
    kinds = np.array([tcode[k] for k in mix]); probs = np.array(list(mix.values()), dtype=float); probs /= probs.sum()
    rows = []
    alleles = []
    a_off = 0
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    n_contigs = len(lengths)
    for tid, L in enumerate(lengths):
        n = int(L // spacing)
        if n == 0:
            continue
        pos = (np.arange(n) * spacing + spacing // 2 + rng.integers(-spacing // 4, spacing // 4, n)).astype(np.int64)
        pos = pos[(pos > 5000) & (pos < L - min(30000, L // 8))]
        typ = kinds[rng.choice(len(kinds), size=len(pos), p=probs)]
        lo, hi = size_range
        size = np.exp(rng.uniform(np.log(lo), np.log(hi), len(pos))).astype(np.int64)
        for p, t, s in zip(pos, typ, size):
            row = dict(tid=tid, pos=int(p), len=int(s), type=int(t), copies=0, tid2=0, pos2=0, strand2=0,
                       vaf=float(rng.choice([0.5, 1.0])), allele_off=0)
            if t == 1:
                if ins_size_uniform:
                    row["len"] = int(rng.integers(ins_size_uniform[0], ins_size_uniform[1] + 1))
                seq = lut[rng.integers(0, 4, row["len"], dtype=np.uint8)]
                row["allele_off"] = a_off; alleles.append(seq); a_off += len(seq)
            elif t == 2:
                row["len"] = int(np.exp(rng.uniform(np.log(200), np.log(20000))))
            elif t == 3:
                row["len"] = int(np.exp(rng.uniform(np.log(100), np.log(3000)))); row["copies"] = int(rng.integers(1, 4))
            elif t == 4:
                t2 = int(rng.integers(0, n_contigs))
                row["tid2"] = t2; row["pos2"] = int(rng.integers(20000, max(20001, lengths[t2] - 40000)))
                row["strand2"] = int(rng.integers(0, 2))
            elif t == 5:
                t2 = int(rng.integers(0, n_contigs))
                row["len"] = int(np.exp(rng.uniform(np.log(100), np.log(2000))))
                row["tid2"] = t2; row["pos2"] = int(rng.integers(20000, max(20001, lengths[t2] - 40000)))
            rows.append(row)
        # hotspots: dense runs of small deletions chained into one partition
        for _ in range(hotspots if tid == 0 else 0):
            start = int(rng.integers(min(50000, L // 10), max(min(50000, L // 10) + 1, L - max(200000, 0) if L > 400000 else L // 2)))
            k = int(rng.integers(hotspot_svs[0], hotspot_svs[1] + 1))
            p = start
            for _j in range(k):
                p += int(rng.integers(150, 260))
                rows.append(dict(tid=tid, pos=p, len=int(rng.integers(50, 100)), type=0, copies=0, tid2=0, pos2=0,
                                 strand2=0, vaf=1.0, allele_off=0))
    rows.sort(key=lambda r: (r["tid"], r["pos"]))
    # enforce a minimum spacing so events do not overlap
    kept, last = [], (-1, -10**9)
    for r in rows:
        if (r["tid"], r["pos"]) > (last[0], last[1] + 120):
            kept.append(r); last = (r["tid"], r["pos"] + (r["len"] if r["type"] in (0, 2) else 0))
    svs = np.zeros(len(kept), dtype=SV_DTYPE)
    for i, r in enumerate(kept):
        for k, v in r.items():
            svs[i][k] = v
    blob = np.concatenate(alleles) if alleles else np.zeros(1, np.uint8)
    return svs, blob


def generate(names, lengths, n_reads, seed, svs, alleles, len_mean=15000, len_sd=3000, len_min=1000, len_max=40000,
             p_ins=0.07, p_del=0.04, geo_ins=0.75, geo_del=0.80, p_lowmapq=0.02, p_secondary=0.01, p_unmapped=0.005,
             p_split=0.4) -> AlignmentBatch:
    lib = _get()
    clen = np.asarray(lengths, dtype=np.int64)
    names_b = b"\x00".join(n.encode() for n in names) + b"\x00"
    svs = np.ascontiguousarray(svs); alleles = np.ascontiguousarray(alleles, dtype=np.uint8)
    cfg = _Config(seed, len(names), 0, clen.ctypes.data, names_b, n_reads, len_mean, len_sd, len_min, len_max,
                  p_ins, p_del, geo_ins, geo_del, p_lowmapq, p_secondary, p_unmapped, p_split,
                  len(svs), svs.ctypes.data, alleles.ctypes.data)
    sizes = _Sizes()
    handle = lib.synth_plan(ctypes.byref(cfg), ctypes.byref(sizes))
    n = sizes.n_records
    arrays = {name: np.zeros(n, dtype=dt) for name, dt in AlignmentBatch.FIELDS}
    cigar = np.zeros(sizes.cigar_words, dtype=np.uint32)
    seq = np.zeros(sizes.seq_bytes, dtype=np.uint8)
    sa = np.zeros(max(1, sizes.sa_bytes), dtype=np.uint8)
    out = _Out(*[arrays[k].ctypes.data for k in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off",
                                                  "sa_off", "sa_len", "qname_id")], cigar.ctypes.data, seq.ctypes.data,
               sa.ctypes.data)
    lib.synth_fill(ctypes.byref(cfg), handle, ctypes.byref(out))
    lib.synth_free(handle)
    return AlignmentBatch(names, clen, arrays, cigar, seq, sa[:sizes.sa_bytes], None, "coordinate")


def generate_shard(names, lengths, n_reads, seed, svs, alleles, rank, world, len_mean=15000, len_sd=3000, len_min=1000, len_max=40000,
                   p_ins=0.07, p_del=0.04, geo_ins=0.75, geo_del=0.80, p_lowmapq=0.02, p_secondary=0.01, p_unmapped=0.005, p_split=0.4):
    """Rank `rank`'s contiguous record range of the same coordinate-sorted output `generate` would produce: the whole input is
    planned (every rank plans the same reads), the range is cut by CIGAR volume (svim_b200.parallel.shard_ranges) and only its
    records are materialised.  Returns (batch of the range with range-relative blob offsets, first record index, total records)."""
    from ..parallel import shard_ranges
    lib = _get()
    clen = np.asarray(lengths, dtype=np.int64)
    names_b = b"\x00".join(n.encode() for n in names) + b"\x00"
    svs = np.ascontiguousarray(svs); alleles = np.ascontiguousarray(alleles, dtype=np.uint8)
    cfg = _Config(seed, len(names), 0, clen.ctypes.data, names_b, n_reads, len_mean, len_sd, len_min, len_max,
                  p_ins, p_del, geo_ins, geo_del, p_lowmapq, p_secondary, p_unmapped, p_split,
                  len(svs), svs.ctypes.data, alleles.ctypes.data)
    sizes = _Sizes()
    handle = lib.synth_plan(ctypes.byref(cfg), ctypes.byref(sizes))
    total = sizes.n_records
    n_cigar = np.zeros(total, dtype=np.uint32)
    lib.synth_plan_ncigar(handle, n_cigar.ctypes.data)
    lo, hi = shard_ranges(n_cigar, world)[rank]
    rs = _Sizes()
    lib.synth_range_sizes(handle, lo, hi, ctypes.byref(rs))
    n = hi - lo
    arrays = {name: np.zeros(n, dtype=dt) for name, dt in AlignmentBatch.FIELDS}
    cigar = np.zeros(rs.cigar_words, dtype=np.uint32)
    seq = np.zeros(rs.seq_bytes, dtype=np.uint8)
    sa = np.zeros(max(1, rs.sa_bytes), dtype=np.uint8)
    out = _Out(*[arrays[k].ctypes.data for k in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off",
                                                  "sa_off", "sa_len", "qname_id")], cigar.ctypes.data, seq.ctypes.data, sa.ctypes.data)
    lib.synth_fill_range(ctypes.byref(cfg), handle, lo, hi, ctypes.byref(out))
    lib.synth_free(handle)
    return AlignmentBatch(names, clen, arrays, cigar, seq, sa[:rs.sa_bytes], None, "coordinate"), lo, total


def config_layout(name: str, scale: float = 1.0):
    """(contig names, lengths, reads, seed, planting keywords, generator keywords) of a BASELINE config at `scale`."""
    c = dict(CONFIGS[name])
    n_contigs = c.pop("contigs"); G = int(c.pop("genome") * scale); reads = max(10, int(c.pop("reads") * scale))
    seed = c.pop("seed")
    if n_contigs == 1:
        names, lengths = ["chr1"], [G]
    else:
        tot = sum(_HUMAN)
        names = ["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY"]
        lengths = [max(200_000, int(G * h / tot)) for h in _HUMAN]
    hotspots = int(round(c.pop("hotspots", 0) * scale)) if "hotspots" in c else 0
    plant_kw = {k: c.pop(k) for k in ("spacing", "mix", "ins_size_uniform", "hotspot_svs") if k in c}
    plant_kw["hotspots"] = hotspots
    return names, lengths, reads, seed, plant_kw, c


# human chr1-22,X,Y lengths (Mb, rounded) used only as proportions for config 4
_HUMAN = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]

CONFIGS = {
    # name: (contigs, total genome bp, reads, seed, extra)
    "config1": dict(contigs=1, genome=1_000_000, reads=1000, seed=1, len_mean=10000, len_sd=1000,
                    mix={"DEL": 0.5, "INS": 0.5}),
    "config2": dict(contigs=1, genome=250_000_000, reads=500_000, seed=2),
    "config3": dict(contigs=1, genome=100_000_000, reads=200_000, seed=3, spacing=15000, mix={"INS": 1.0},
                    ins_size_uniform=(200, 5000)),
    "config4": dict(contigs=24, genome=2_500_000_000, reads=5_000_000, seed=4),
    "config5": dict(contigs=1, genome=300_000_000, reads=2_000_000, seed=5, hotspots=200, hotspot_svs=(50, 500)),
}


def make_config(name: str, scale: float = 1.0, with_genome: bool = True):
    """Build a BASELINE config; `scale` shrinks genome and read count together
    (coverage, and therefore partition sizes, are preserved)."""
    c = dict(CONFIGS[name])
    n_contigs = c.pop("contigs"); G = int(c.pop("genome") * scale); reads = max(10, int(c.pop("reads") * scale))
    seed = c.pop("seed")
    if n_contigs == 1:
        names, lengths = ["chr1"], [G]
    else:
        tot = sum(_HUMAN)
        names = ["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY"]
        lengths = [max(200_000, int(G * h / tot)) for h in _HUMAN]
    hotspots = int(round(c.pop("hotspots", 0) * scale)) if "hotspots" in c else 0
    plant_kw = {k: c.pop(k) for k in ("spacing", "mix", "ins_size_uniform", "hotspot_svs") if k in c}
    svs, alleles = plant_svs(lengths, seed, hotspots=hotspots, **plant_kw)
    batch = generate(names, lengths, reads, seed, svs, alleles, **c)
    genome = random_genome(names, lengths, 1524 + seed) if with_genome else None
    return batch, genome, {"svs": svs, "alleles": alleles}
