"""ctypes binding of include/svimgpu.h (libsvimgpu.so, built in-tree by svim_b200.build).

There is no CPU implementation behind this module: if the shared library is missing
it raises, and `Context()` raises when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvimgpu.so")

TYPE_NAMES = ("DEL", "INS", "INV", "DUP_TAN", "BND", "DUP_INT")     # enum order of svimgpu.h
TYPE_CODE = {n: i for i, n in enumerate(TYPE_NAMES)}
TYPE_DUP_INT_CAND = 6                     # cluster-stage only: DUP_INT candidate (svimgpu.h)
INV_DIRECTIONS = ("left_fwd", "left_rev", "right_fwd", "right_rev", "all")

F_SUPPL, F_FULLY_COVERED, F_DIR1_REV, F_DIR2_REV, F_INVDIR_SHIFT = 1, 2, 4, 8, 4


class SvimGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("svimgpu error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("min_sv_size", C.c_int32), ("max_sv_size", C.c_int32),
                ("segment_gap_tolerance", C.c_int32), ("segment_overlap_tolerance", C.c_int32), ("all_bnds", C.c_int32),
                ("partition_max_distance", C.c_double), ("position_distance_normalizer", C.c_double),
                ("edit_distance_normalizer", C.c_double), ("cluster_max_distance", C.c_double)]

    @classmethod
    def from_options(cls, options=None, **kw):
        """Build from the reference's argparse Namespace (SVIM_input_parsing.py) or keywords."""
        d = dict(min_mapq=20, min_sv_size=40, max_sv_size=100000, segment_gap_tolerance=10, segment_overlap_tolerance=5,
                 all_bnds=False, partition_max_distance=1000, position_distance_normalizer=900,
                 edit_distance_normalizer=1.0, cluster_max_distance=0.5)
        if options is not None:
            for k in d:
                if hasattr(options, k):
                    d[k] = getattr(options, k)
        d.update(kw)
        d["all_bnds"] = 1 if d["all_bnds"] else 0
        return cls(**d)


class AlnSoa(C.Structure):
    _fields_ = [("n_aln", C.c_int64)] + [(n, C.c_void_p) for n in
                ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off", "sa_off", "sa_len", "qname_id")] + \
               [("cigar", C.c_void_p), ("cigar_words", C.c_int64), ("seq", C.c_void_p), ("seq_bytes", C.c_int64),
                ("sa", C.c_void_p), ("sa_bytes", C.c_int64),
                ("cigar16", C.c_void_p), ("cigar16_words", C.c_int64), ("cigar16_off", C.c_void_p),
                ("cigar8", C.c_void_p), ("cigar8_bytes", C.c_int64), ("cigar8_off", C.c_void_p)]


SIG_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("pos", "<i4"), ("contig1", "<i4"), ("contig2", "<i4"),
                      ("aln_idx", "<u4"), ("qname_id", "<u4"), ("ordinal", "<u4"), ("seq_off", "<u8"), ("seq_len", "<u4"),
                      ("type", "u1"), ("flags", "u1"), ("copies", "<u2")])
CSIG_DTYPE = np.dtype([("start", "<f8"), ("end", "<f8"), ("dpos", "<f8"), ("contig_a", "<i4"), ("contig_b", "<i4"),
                       ("read_id", "<u4"), ("seq_len", "<u4"), ("seq_off", "<u8"), ("type", "u1"), ("dirs", "u1"),
                       ("copies", "<u2"), ("pad", "<u4", (3,))])
CLUSTER_DTYPE = np.dtype([("start", "<i8"), ("end", "<i8"), ("dest_start", "<i8"), ("dest_end", "<i8"), ("score", "<f8"),
                          ("std_span", "<f8"), ("std_pos", "<f8"), ("member_off", "<u4"), ("size", "<u4"), ("type", "u1"),
                          ("dir1_rev", "u1"), ("dir2_rev", "u1"), ("pad0", "u1"), ("pad1", "<u4")])
GENO_CAND_DTYPE = np.dtype([("start", "<i8"), ("end", "<i8"), ("tid", "<i4"), ("n_variant_reads", "<u4"), ("variant_off", "<u8")])
GENO_RESULT_DTYPE = np.dtype([("support_fraction", "<f8"), ("ref_reads", "<i4"), ("alt_reads", "<i4"), ("genotype", "u1"),
                              ("status", "u1"), ("pad", "<u2"), ("n_fetched", "<u4")])
GENOTYPES = ("1/1", "0/1", "0/0", "./.")
assert SIG_DTYPE.itemsize == 48 and CSIG_DTYPE.itemsize == 64 and CLUSTER_DTYPE.itemsize == 72
assert GENO_CAND_DTYPE.itemsize == 32 and GENO_RESULT_DTYPE.itemsize == 24


class GenoParams(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("minimum_depth", C.c_int32), ("homozygous_threshold", C.c_double),
                ("heterozygous_threshold", C.c_double)]

    @classmethod
    def from_options(cls, options=None, **kw):
        """Genotyping options of the reference's Namespace (SVIM_input_parsing.py:404-437) or keywords."""
        d = dict(min_mapq=20, minimum_depth=4, homozygous_threshold=0.8, heterozygous_threshold=0.2)
        if options is not None:
            for k in d:
                if hasattr(options, k):
                    d[k] = getattr(options, k)
        d.update(kw)
        return cls(**d)


class ClusterStats(C.Structure):
    _fields_ = [("n_partitions", C.c_int64 * 6), ("n_clusters", C.c_int64 * 6), ("large_partitions", C.c_int64 * 6),
                ("duplicate_signatures", C.c_int64 * 6), ("n_members", C.c_int64), ("n_clusters_total", C.c_int64),
                ("myers_pairs", C.c_int64), ("myers_cells", C.c_int64),
                ("myers_banded_pairs", C.c_int64), ("myers_retry_pairs", C.c_int64), ("myers_band_cells", C.c_int64),
                ("myers_tpp_pairs", C.c_int64), ("myers_tpp_cells", C.c_int64), ("myers_unbanded_cells", C.c_int64)]


class BamInfo(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_records", "cigar_words", "seq_bytes", "sa_bytes", "names_bytes", "n_names", "inflated_bytes")]


class CollectStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_signatures", "n_twin_signatures", "ins_bytes", "twin_ins_bytes", "n_sa_bad_fields",
                                         "n_no_read_length", "n_primaries", "n_data_errors")]


EXPORTS = {
    # name: (restype, argtypes)
    "svimgpu_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Params)]),
    "svimgpu_destroy": (None, [C.c_void_p]),
    "svimgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "svimgpu_set_params": (C.c_int, [C.c_void_p, C.POINTER(Params)]),
    "svimgpu_version": (C.c_char_p, []),
    "svimgpu_pci_bus_id": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "svimgpu_set_contigs": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_void_p]),
    "svimgpu_set_genome": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "svimgpu_pin_host": (C.c_int, [C.c_void_p, C.c_int64]),
    "svimgpu_unpin_host": (C.c_int, [C.c_void_p]),
    "svimgpu_upload_alignments": (C.c_int, [C.c_void_p, C.POINTER(AlnSoa)]),
    "svimgpu_download_cigar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "svimgpu_decode_bam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.POINTER(BamInfo)]),
    "svimgpu_fetch_bam_names": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svimgpu_download_alignments": (C.c_int, [C.c_void_p] * 15),
    "svimgpu_collect": (C.c_int, [C.c_void_p, C.POINTER(CollectStats)]),
    "svimgpu_collect_host": (C.c_int, [C.c_void_p, C.POINTER(AlnSoa), C.POINTER(CollectStats)]),
    "svimgpu_collect_host_querysorted": (C.c_int, [C.c_void_p, C.POINTER(AlnSoa), C.POINTER(CollectStats)]),
    "svimgpu_fetch_signatures": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "svimgpu_signatures_host": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "svimgpu_use_collected": (C.c_int, [C.c_void_p, C.c_int]),
    "svimgpu_set_signatures": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "svimgpu_cluster": (C.c_int, [C.c_void_p, C.POINTER(ClusterStats)]),
    "svimgpu_partition": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "svimgpu_fetch_clusters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "svimgpu_clusters_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "svimgpu_mirror_gathered_ins": (C.c_int, [C.c_void_p, C.c_int]),
    "svimgpu_peer_ins_active": (C.c_int, [C.c_void_p]),
    "svimgpu_fetch_partitions": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "svimgpu_genotype": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(GenoParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int32, C.c_void_p]),
    "svimgpu_closest_source": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_void_p, C.c_void_p]),
    "svimgpu_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "svimgpu_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "svimgpu_exchange_signatures": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(CollectStats)]),
    "svimgpu_cluster_sharded": (C.c_int, [C.c_void_p, C.POINTER(ClusterStats)]),
    "svimgpu_barrier_max": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "svimgpu_cigar_indel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "svimgpu_edit_distance": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svimgpu_linkage_average": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    "svimgpu_sample_indices": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "svimgpu_last_timings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "svimgpu_timing_name": (C.c_char_p, [C.c_int32]),
    "svimgpu_timer_start": (C.c_int, [C.c_void_p]),
    "svimgpu_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "svimgpu_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


def _preload_nccl():
    """libsvimgpu.so links libnccl.so.2 by soname.  PyTorch bundles a newer NCCL under the same soname; whichever
    is loaded first wins for the whole process, and torch cannot run on the older system copy.  Load the bundled
    one first (when present) so that `import torch` keeps working in either import order."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            p = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(p):
                C.CDLL(p, mode=C.RTLD_GLOBAL)
                return p
    except Exception:
        pass
    return None


def load():
    """dlopen libsvimgpu.so and bind every export declared in include/svimgpu.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libsvimgpu.so is missing: run `python -m svim_b200.build` (needs nvcc); "
                              "svim_b200 has no CPU fallback")
        _preload_nccl()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class Context:
    """Owns one svimgpu_ctx (one GPU, one stream)."""

    def __init__(self, params: Params = None, device: int = 0):
        self.lib = load()
        self.params = params or Params.from_options()
        h = C.c_void_p()
        rc = self.lib.svimgpu_create(C.byref(h), device, C.byref(self.params))
        if rc != 0:
            raise SvimGpuError(rc, "svimgpu_create failed (no usable CUDA device?) - there is no CPU path")
        self.h = h
        self._keep = []
        self.genome_key = None      # bookkeeping used by svim_b200.runtime / the host mirror
        self.resident = None
        self.contigs_key = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.svimgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SvimGpuError(rc, self.lib.svimgpu_last_error(self.h).decode("utf-8", "replace"))

    def pci_bus_id(self) -> str:
        buf = C.create_string_buffer(64)
        self._check(self.lib.svimgpu_pci_bus_id(self.h, buf, 64))
        return buf.value.decode()

    def set_params(self, params: Params):
        self.params = params
        self._check(self.lib.svimgpu_set_params(self.h, C.byref(params)))

    def set_contigs(self, names):
        blob = b"".join(n.encode("ascii") for n in names)
        off = np.zeros(len(names) + 1, dtype=np.int32)
        np.cumsum([len(n.encode("ascii")) for n in names], out=off[1:])
        self._check(self.lib.svimgpu_set_contigs(self.h, len(names), blob, off.ctypes.data))

    def set_genome(self, genome):
        """genome: svim_b200.io.Genome (contig order must match the alignment header)."""
        off = np.ascontiguousarray(genome.offsets, dtype=np.int64)
        blob = np.ascontiguousarray(genome.blob, dtype=np.uint8)
        self._check(self.lib.svimgpu_set_genome(self.h, len(genome.names), off.ctypes.data, blob.ctypes.data))

    @staticmethod
    def soa_of(batch) -> AlnSoa:
        s = AlnSoa()
        s.n_aln = batch.n
        for n in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off", "sa_off", "sa_len", "qname_id"):
            setattr(s, n, _ptr(getattr(batch, n)))
        s.cigar = _ptr(batch.cigar); s.cigar_words = batch.cigar.size
        s.seq = _ptr(batch.seq); s.seq_bytes = batch.seq.size
        s.sa = _ptr(batch.sa); s.sa_bytes = batch.sa.size
        c16 = getattr(batch, "cigar16", None)
        if c16 is not None:          # packed CIGAR stream (AlignmentBatch.pack_cigar16): uploaded instead of the uint32 words
            s.cigar16 = _ptr(c16); s.cigar16_words = c16.size; s.cigar16_off = _ptr(batch.cigar16_off)
        c8 = getattr(batch, "cigar8", None)
        if c8 is not None:           # 8-bit packed stream (AlignmentBatch.pack_cigar8): preferred over cigar16
            s.cigar8 = _ptr(c8); s.cigar8_bytes = c8.size; s.cigar8_off = _ptr(batch.cigar8_off)
        return s

    def pin(self, arr):
        if arr.size:
            rc = self.lib.svimgpu_pin_host(arr.ctypes.data, arr.nbytes)
            if rc != 0:
                raise SvimGpuError(rc, "cudaHostRegister failed")

    def unpin(self, arr):
        if arr.size:
            self.lib.svimgpu_unpin_host(arr.ctypes.data)

    def upload(self, batch):
        soa = self.soa_of(batch)
        self._check(self.lib.svimgpu_upload_alignments(self.h, C.byref(soa)))

    def decode_bam(self, file_bytes, blocks, first_record, n_ref) -> BamInfo:
        """file_bytes: uint8 array (or memmap) of the whole .bam; blocks: io.BGZF_BLOCK_DTYPE table.  Leaves the records resident."""
        info = BamInfo()
        self._check(self.lib.svimgpu_decode_bam(self.h, file_bytes.ctypes.data, file_bytes.size, blocks.ctypes.data if len(blocks) else None, len(blocks),
                                                first_record, n_ref, C.byref(info)))
        return info

    def fetch_bam_names(self, info: BamInfo):
        names = np.zeros(max(1, info.names_bytes), dtype=np.uint8); off = np.zeros(info.n_records, dtype=np.uint64)
        rec = np.zeros(info.n_names, dtype=np.uint32); qid = np.zeros(info.n_records, dtype=np.uint32)
        self._check(self.lib.svimgpu_fetch_bam_names(self.h, _ptr(names), _ptr(off), _ptr(rec), _ptr(qid)))
        return names[:info.names_bytes], off, rec, qid

    def download_alignments(self, info: BamInfo):
        from .records import AlignmentBatch
        n = info.n_records
        arrays = {name: np.zeros(n, dtype=dt) for name, dt in AlignmentBatch.FIELDS}
        cigar = np.zeros(info.cigar_words, dtype=np.uint32); seq = np.zeros(info.seq_bytes, dtype=np.uint8); sa = np.zeros(info.sa_bytes, dtype=np.uint8)
        p = lambda a: _ptr(a)
        self._check(self.lib.svimgpu_download_alignments(self.h, *[p(arrays[k]) for k in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off",
                                                                                           "sa_off", "sa_len", "qname_id")], p(cigar), p(seq), p(sa)))
        return arrays, cigar, seq, sa

    def download_cigar(self, words: int):
        out = np.zeros(words, dtype=np.uint32)
        self._check(self.lib.svimgpu_download_cigar(self.h, _ptr(out), words))
        return out

    def collect(self) -> CollectStats:
        st = CollectStats()
        self._check(self.lib.svimgpu_collect(self.h, C.byref(st)))
        return st

    def collect_host(self, batch) -> CollectStats:
        soa = self.soa_of(batch)
        st = CollectStats()
        self._check(self.lib.svimgpu_collect_host(self.h, C.byref(soa), C.byref(st)))
        return st

    def collect_host_querysorted(self, batch) -> CollectStats:
        soa = self.soa_of(batch)
        st = CollectStats()
        self._check(self.lib.svimgpu_collect_host_querysorted(self.h, C.byref(soa), C.byref(st)))
        return st

    def fetch_signatures(self, which, stats: CollectStats):
        n = stats.n_signatures if which == 0 else stats.n_twin_signatures
        nb = stats.ins_bytes if which == 0 else stats.twin_ins_bytes
        # collect_host mirrors the lists into pinned host memory while CLUSTER runs: hand out views of that copy
        # (valid until the next collect on this context; callers that keep them longer must .copy())
        ps, pi = C.c_void_p(), C.c_void_p()
        self._check(self.lib.svimgpu_signatures_host(self.h, which, C.byref(ps), C.byref(pi)))
        if ps.value:
            sigs = np.frombuffer((C.c_uint8 * (n * SIG_DTYPE.itemsize)).from_address(ps.value), dtype=SIG_DTYPE) if n else np.zeros(0, dtype=SIG_DTYPE)
            # pi is NULL on a rank that mirrors the gathered records only (svimgpu_mirror_gathered_ins)
            ins = np.frombuffer((C.c_uint8 * nb).from_address(pi.value), dtype=np.uint8) if nb and pi.value else np.zeros(0, dtype=np.uint8)
            return sigs, ins
        sigs = np.zeros(n, dtype=SIG_DTYPE)
        ins = np.zeros(nb, dtype=np.uint8)
        self._check(self.lib.svimgpu_fetch_signatures(self.h, which, _ptr(sigs), _ptr(ins)))
        return sigs, ins

    def use_collected(self, which=0):
        self._check(self.lib.svimgpu_use_collected(self.h, which))

    def set_signatures(self, csig, ins_blob=None, rank_to_tid=None):
        csig = np.ascontiguousarray(csig, dtype=CSIG_DTYPE)
        ins = np.ascontiguousarray(ins_blob, dtype=np.uint8) if ins_blob is not None else np.zeros(0, np.uint8)
        r2t = np.ascontiguousarray(rank_to_tid, dtype=np.int32) if rank_to_tid is not None else np.zeros(0, np.int32)
        self._check(self.lib.svimgpu_set_signatures(self.h, len(csig), _ptr(csig), _ptr(ins), ins.size, _ptr(r2t), r2t.size))

    def cluster(self, sharded=False, view=False):
        """CLUSTER on the selected signatures -> (stats, clusters, members).  `view=True` hands out the context's own pinned
        result arrays instead of copies: valid until the next cluster on this context (callers that keep them must .copy())."""
        st = ClusterStats()
        fn = self.lib.svimgpu_cluster_sharded if sharded else self.lib.svimgpu_cluster
        self._check(fn(self.h, C.byref(st)))
        if view:
            pc, pm = C.c_void_p(), C.c_void_p()
            self._check(self.lib.svimgpu_clusters_host(self.h, C.byref(pc), C.byref(pm)))
            nc, nm = st.n_clusters_total, st.n_members
            clusters = (np.frombuffer((C.c_uint8 * (nc * CLUSTER_DTYPE.itemsize)).from_address(pc.value), dtype=CLUSTER_DTYPE)
                        if nc and pc.value else np.zeros(0, dtype=CLUSTER_DTYPE))
            members = np.frombuffer((C.c_uint32 * nm).from_address(pm.value), dtype=np.uint32) if nm and pm.value else np.zeros(0, dtype=np.uint32)
            return st, clusters, members
        clusters = np.zeros(st.n_clusters_total, dtype=CLUSTER_DTYPE)
        members = np.zeros(st.n_members, dtype=np.uint32)
        self._check(self.lib.svimgpu_fetch_clusters(self.h, _ptr(clusters), _ptr(members)))
        return st, clusters, members

    def partition(self):
        n = C.c_int64()
        self._check(self.lib.svimgpu_partition(self.h, C.byref(n)))
        return n.value

    def fetch_partitions(self, n_sigs):
        npart = C.c_int64()
        self._check(self.lib.svimgpu_fetch_partitions(self.h, C.byref(npart), None, None))
        order = np.zeros(n_sigs, dtype=np.uint32)
        off = np.zeros(npart.value + 1, dtype=np.uint32)
        self._check(self.lib.svimgpu_fetch_partitions(self.h, C.byref(npart), _ptr(order), off.ctypes.data))
        return order, off

    def genotype(self, type_code, gparams: "GenoParams", cands, variant_ids, contig_lengths):
        """svimgpu_genotype on the resident records -> GENO_RESULT_DTYPE[len(cands)]."""
        cands = np.ascontiguousarray(cands, dtype=GENO_CAND_DTYPE)
        variant_ids = np.ascontiguousarray(variant_ids, dtype=np.uint32)
        clen = np.ascontiguousarray(contig_lengths, dtype=np.int64)
        out = np.zeros(len(cands), dtype=GENO_RESULT_DTYPE)
        self._check(self.lib.svimgpu_genotype(self.h, type_code, C.byref(gparams), len(cands), _ptr(cands), _ptr(variant_ids), variant_ids.size,
                                              clen.ctypes.data, clen.size, _ptr(out)))
        return out

    def closest_source(self, a_start, a_end, b_start, b_end, normalizer):
        """svimgpu_closest_source -> (index int64[n_a], distance float64[n_a])."""
        a_s = np.ascontiguousarray(a_start, dtype=np.int64); a_e = np.ascontiguousarray(a_end, dtype=np.int64)
        b_s = np.ascontiguousarray(b_start, dtype=np.int64); b_e = np.ascontiguousarray(b_end, dtype=np.int64)
        idx = np.zeros(len(a_s), dtype=np.int64); dist = np.zeros(len(a_s), dtype=np.float64)
        self._check(self.lib.svimgpu_closest_source(self.h, len(a_s), _ptr(a_s), _ptr(a_e), len(b_s), _ptr(b_s), _ptr(b_e), float(normalizer),
                                                    _ptr(idx), _ptr(dist)))
        return idx, dist

    def timings(self):
        ms = np.zeros(32, dtype=np.float64)
        n = C.c_int32()
        self._check(self.lib.svimgpu_last_timings(self.h, ms.ctypes.data, 32, C.byref(n)))
        return {self.lib.svimgpu_timing_name(i).decode(): float(ms[i]) for i in range(n.value)}

    def timer_start(self):
        self._check(self.lib.svimgpu_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.lib.svimgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.svimgpu_launch_count(self.h))

    # ---- micro entry points (unit tests) ----
    def cigar_indel(self, cigar_u32, min_len):
        cigar = np.ascontiguousarray(cigar_u32, dtype=np.uint32)
        out = np.zeros((max(1, len(cigar)), 4), dtype=np.int64)
        n = C.c_int64()
        self._check(self.lib.svimgpu_cigar_indel(self.h, _ptr(cigar), len(cigar), min_len, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def edit_distance(self, pairs):
        """pairs: list of (bytes a, bytes b) -> np.int32[len]"""
        blob = bytearray()
        ao, al, bo, bl = [], [], [], []
        for a, b in pairs:
            ao.append(len(blob)); al.append(len(a)); blob += a
            bo.append(len(blob)); bl.append(len(b)); blob += b
        blob_a = np.frombuffer(bytes(blob) + b"\x00", dtype=np.uint8)
        ao = np.asarray(ao, np.int64); al = np.asarray(al, np.int32); bo = np.asarray(bo, np.int64); bl = np.asarray(bl, np.int32)
        out = np.zeros(len(pairs), dtype=np.int32)
        self._check(self.lib.svimgpu_edit_distance(self.h, len(pairs), blob_a.ctypes.data, _ptr(ao), _ptr(al), _ptr(bo), _ptr(bl), _ptr(out)))
        return out

    def linkage_average(self, condensed, m, t):
        d = np.ascontiguousarray(condensed, dtype=np.float64)
        Z = np.zeros((m - 1, 4), dtype=np.float64)
        T = np.zeros(m, dtype=np.int32)
        self._check(self.lib.svimgpu_linkage_average(self.h, d.ctypes.data, m, t, Z.ctypes.data, T.ctypes.data))
        return Z, T


def sample_indices(sizes):
    """Host RNG stream of SVIM_clustering.py:129-134 for a list of partition sizes (no GPU needed)."""
    lib = load()
    sizes = np.ascontiguousarray(sizes, dtype=np.int64)
    k = int((sizes > 100).sum())
    out = np.zeros((k, 100), dtype=np.int32)
    rc = lib.svimgpu_sample_indices(_ptr(sizes), len(sizes), out.ctypes.data if k else None)
    if rc != 0 and k:
        raise SvimGpuError(rc, "svimgpu_sample_indices")
    return out
