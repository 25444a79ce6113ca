"""pysam-free readers/writers: SAM text, BAM (BGZF via zlib), FASTA (+ .fai).

The reference reads its input through pysam/htslib (SVIM_COLLECT.py:133,
SVIM_clustering.py:377), which is not available here.  These readers decode
straight into the flattened record buffer (`records.AlignmentBatch`) – there is
no per-record Python object on the product path.

File formats follow the SAM/BAM specification (SAMv1 §1.4, §4.2).
"""
from __future__ import annotations

import os
import sys
import struct
import zlib
from typing import List, Optional, Tuple

import numpy as np

from .records import AlignmentBatch, BatchBuilder, CIGAR_OPS, SEQ_NT16, parse_cigar_string


# --------------------------------------------------------------------------
# SAM text
# --------------------------------------------------------------------------
def _parse_sam_header(lines) -> Tuple[List[str], List[int], str]:
    names, lengths, so = [], [], "unknown"
    for ln in lines:
        f = ln.rstrip("\n").split("\t")
        if f[0] == "@SQ":
            d = dict(x.split(":", 1) for x in f[1:] if ":" in x)
            names.append(d["SN"]); lengths.append(int(d.get("LN", 0)))
        elif f[0] == "@HD":
            d = dict(x.split(":", 1) for x in f[1:] if ":" in x)
            so = d.get("SO", "unknown")
    return names, lengths, so


def read_sam(path: str) -> AlignmentBatch:
    with open(path, "r") as fh:
        text = fh.read().split("\n")
    header = [l for l in text if l.startswith("@")]
    names, lengths, so = _parse_sam_header(header)
    tid_of = {n: i for i, n in enumerate(names)}
    b = BatchBuilder(names, lengths, so)
    for ln in text:
        if not ln or ln.startswith("@"):
            continue
        f = ln.split("\t")
        sa = None
        for tag in f[11:]:
            if tag.startswith("SA:Z:"):
                sa = tag[5:]
        tid = -1 if f[2] == "*" else tid_of.get(f[2], -1)
        b.add(f[0], int(f[1]), tid, int(f[3]) - 1, int(f[4]), parse_cigar_string(f[5]),
              None if f[9] == "*" else f[9], sa)
    return b.finish()


def write_sam(path: str, batch: AlignmentBatch):
    with open(path, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:%s\n" % batch.sort_order)
        for n, l in zip(batch.contig_names, batch.contig_lengths):
            fh.write("@SQ\tSN:%s\tLN:%d\n" % (n, l))
        for i in range(batch.n):
            tid = int(batch.tid[i])
            cig = batch.cigartuples(i)
            seq = batch.sequence(i)
            cols = [batch.qname(int(batch.qname_id[i])), str(int(batch.flag[i])),
                    batch.contig_names[tid] if tid >= 0 else "*", str(int(batch.pos[i]) + 1),
                    str(int(batch.mapq[i])),
                    "".join("%d%s" % (n, CIGAR_OPS[o]) for o, n in cig) or "*",
                    "*", "0", "0", seq if seq else "*", "*"]
            sa = batch.sa_tag(i)
            if sa:
                cols.append("SA:Z:" + sa)
            fh.write("\t".join(cols) + "\n")


# --------------------------------------------------------------------------
# BGZF / BAM
# --------------------------------------------------------------------------
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_blocks(fh):
    while True:
        head = fh.read(12)
        if len(head) < 12:
            return
        if head[:4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF file")
        xlen = struct.unpack("<H", head[10:12])[0]
        extra = fh.read(xlen)
        bsize = None
        o = 0
        while o + 4 <= xlen:
            si1, si2, slen = extra[o], extra[o + 1], struct.unpack("<H", extra[o + 2:o + 4])[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack("<H", extra[o + 4:o + 6])[0]
            o += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC field")
        cdata = fh.read(bsize - xlen - 19)
        fh.read(8)  # crc32, isize
        yield zlib.decompress(cdata, -15)


def _bgzf_inflate_all(path: str) -> bytes:
    with open(path, "rb") as fh:
        return b"".join(_bgzf_blocks(fh))


def _aux_find(aux: bytes):
    """Walk BAM aux fields: (SA:Z payload without NUL or None, CG:B,I payload as uint32 array or None)."""
    o, n = 0, len(aux)
    sizes = {ord("A"): 1, ord("c"): 1, ord("C"): 1, ord("s"): 2, ord("S"): 2,
             ord("i"): 4, ord("I"): 4, ord("f"): 4}
    sa = cg = None
    while o + 3 <= n:
        tag, typ = aux[o:o + 2], aux[o + 2]
        o += 3
        if typ in sizes:
            o += sizes[typ]
        elif typ in (ord("Z"), ord("H")):
            e = aux.index(b"\x00", o)
            if tag == b"SA" and typ == ord("Z") and sa is None:
                sa = aux[o:e]
            o = e + 1
        elif typ == ord("B"):
            sub = aux[o]; cnt = struct.unpack("<I", aux[o + 1:o + 5])[0]
            if tag == b"CG" and sub == ord("I") and cg is None:
                cg = np.frombuffer(aux, dtype="<u4", count=cnt, offset=o + 5).copy()
            o += 5 + cnt * sizes[sub]
        else:
            raise ValueError("bad aux type %r" % typ)
    return sa, cg


def _aux_find_sa(aux: bytes) -> Optional[bytes]:
    return _aux_find(aux)[0]


def _long_cigar_core(l_seq: int, rlen: int) -> np.ndarray:
    """SAMv1 §4.2.2: a CIGAR of more than 65535 operations is stored in the CG:B,I tag and the record carries the
    placeholder `<l_seq>S<reference length>N` (what htslib writes and, on reading, replaces by the tag: sam.c bam_tag2cigar)."""
    return np.array([(l_seq << 4) | 4, (rlen << 4) | 3], dtype="<u4")


# ---- native multi-threaded reader (csrc_host/bamio.cpp) -----------------------------------------------------
_BAMIO_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc_host", "bamio.cpp")
_BAMIO_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsvimio.so")
_bamio = None


def build_bamio(force: bool = False) -> str:
    import subprocess
    if force or not os.path.exists(_BAMIO_SO) or (os.path.exists(_BAMIO_SRC) and os.path.getmtime(_BAMIO_SO) < os.path.getmtime(_BAMIO_SRC)):
        from .build import run_atomic
        run_atomic(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", "@OUT@", _BAMIO_SRC, "-lz"], _BAMIO_SO)
    return _BAMIO_SO


def _bamio_lib():
    global _bamio
    if _bamio is None:
        import ctypes as C
        lib = C.CDLL(build_bamio())
        lib.bamio_open.restype = C.c_void_p
        lib.bamio_open.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        lib.bamio_header.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_char_p]
        lib.bamio_qnames.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        lib.bamio_fill.argtypes = [C.c_void_p, C.c_void_p]
        lib.bamio_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        lib.bamio_close.argtypes = [C.c_void_p]
        lib.bamio_close.restype = None
        lib.bamio_layout.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.bamio_blocks.argtypes = [C.c_void_p, C.c_void_p]
        lib.bamio_pack_cigar16.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.bamio_pack_cigar8.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.bamio_set_pack.argtypes = [C.c_void_p, C.c_int]
        lib.bamio_cigar8_bytes.argtypes = [C.c_void_p]; lib.bamio_cigar8_bytes.restype = C.c_int64
        lib.bamio_cigar8_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _bamio = lib
    return _bamio


def _reserve(n: int, dtype) -> np.ndarray:
    """Uninitialised array of `n` items whose pages are committed only when written: an anonymous MAP_NORESERVE mapping, so an
    upper-bound allocation is not refused by the kernel's overcommit heuristic; falls back to np.empty."""
    nbytes = max(1, int(n)) * np.dtype(dtype).itemsize
    try:
        import mmap
        flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | getattr(mmap, "MAP_NORESERVE", 0x4000 if sys.platform.startswith("linux") else 0)
        return np.frombuffer(mmap.mmap(-1, nbytes, flags=flags), dtype=dtype)[:max(0, int(n))]
    except (OSError, ValueError, AttributeError):
        return np.empty(int(n), dtype=dtype)


def read_bam_native(path: str, threads: int = 0, pack_cigar: int = 0) -> AlignmentBatch:
    """BAM -> AlignmentBatch with the streaming multi-threaded decoder (csrc_host/bamio.cpp): blocks are
    inflated into cache-warm per-thread buffers and the CIGAR / SEQ bytes go straight into the arrays allocated here (to an
    upper bound computed from the BGZF index; only the pages actually written are ever committed).
    pack_cigar=8: the decode also emits the 8-bit packed CIGAR stream (svim_aln_soa.cigar8) from the same warm bytes, so the
    batch crosses PCIe at ~1.1 bytes per operation without a separate packing pass."""
    import ctypes as C
    lib = _bamio_lib()
    class Info(C.Structure):
        _fields_ = [(n, C.c_int64) for n in ("n_records", "cigar_words", "seq_bytes", "sa_bytes", "n_qnames", "names_bytes",
                                             "cigar_bound_words", "seq_bound_bytes")] + \
                   [("n_contigs", C.c_int32), ("sorted_coordinate", C.c_int32)]
    inf = Info()
    err = C.create_string_buffer(256)
    threads = threads or min(128, os.cpu_count() or 1)
    h = lib.bamio_open(path.encode(), threads, C.byref(inf), err, 256)
    if not h:
        raise ValueError("read_bam_native(%s): %s" % (path, err.value.decode()))
    try:
        names_buf = C.create_string_buffer(max(1, 256 * inf.n_contigs + 64))
        lengths = np.zeros(max(1, inf.n_contigs), dtype=np.int64)
        so = C.create_string_buffer(16)
        if lib.bamio_header(h, names_buf, len(names_buf), lengths.ctypes.data, so) != 0:
            raise ValueError("contig names too long")
        names = names_buf.raw.split(b"\x00")[:inf.n_contigs]
        names = [x.decode("ascii") for x in names]
        cigar = _reserve(inf.cigar_bound_words, np.uint32)
        seq = _reserve(inf.seq_bound_bytes, np.uint8)
        if pack_cigar:
            if lib.bamio_set_pack(h, int(pack_cigar)) != 0:
                raise ValueError("read_bam_native: pack_cigar must be 0 or 8")
        if lib.bamio_decode(h, cigar.ctypes.data, seq.ctypes.data, C.byref(inf), err, 256) != 0:
            raise ValueError("read_bam_native(%s): %s" % (path, err.value.decode()))
        n = inf.n_records
        c8 = off8 = None
        if pack_cigar == 8:
            off8 = np.zeros(n + 1, dtype=np.uint64)
            c8 = np.empty(int(lib.bamio_cigar8_bytes(h)), dtype=np.uint8)
            if lib.bamio_cigar8_fill(h, off8.ctypes.data, c8.ctypes.data if c8.size else None) != 0:
                raise ValueError("read_bam_native: packed CIGAR stream not available")
        arrays = {name: np.empty(n, dtype=dt) for name, dt in AlignmentBatch.FIELDS}
        sa = np.empty(max(1, inf.sa_bytes), dtype=np.uint8)
        ptrs = (C.c_void_p * 14)(*[arrays[k].ctypes.data for k in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off",
                                                                 "sa_off", "sa_len", "qname_id")], cigar.ctypes.data, seq.ctypes.data, sa.ctypes.data)
        lib.bamio_fill(h, ptrs)
        qbuf = C.create_string_buffer(max(1, inf.names_bytes))
        lib.bamio_qnames(h, qbuf, len(qbuf))
        qnames = [x.decode("ascii") for x in qbuf.raw.split(b"\x00")[:inf.n_qnames]]
    finally:
        lib.bamio_close(h)
    batch = AlignmentBatch(names, lengths[:inf.n_contigs], arrays, cigar[:inf.cigar_words], seq[:inf.seq_bytes], sa[:inf.sa_bytes], qnames,
                           so.value.decode() or "unknown")
    if c8 is not None:
        batch.cigar8, batch.cigar8_off = c8, off8
    return batch


BGZF_BLOCK_DTYPE = np.dtype([("coff", "<u8"), ("uoff", "<u8"), ("clen", "<u4"), ("ulen", "<u4")])


def pack_cigar16(batch, threads: int = 0):
    """(cigar16, cigar16_off) of a batch: the 16-bit packed CIGAR stream of include/svimgpu.h, built by csrc_host/bamio.cpp."""
    lib = _bamio_lib()
    threads = threads or (os.cpu_count() or 8)
    n = batch.n
    off = np.zeros(n + 1, dtype=np.uint64)
    args = (n, batch.n_cigar.ctypes.data, batch.cigar_off.ctypes.data, batch.cigar.ctypes.data if batch.cigar.size else None)
    if lib.bamio_pack_cigar16(*args, off.ctypes.data, None, threads) != 0:
        raise ValueError("pack_cigar16: bad arguments")
    out = np.empty(int(off[n]), dtype=np.uint16)
    if lib.bamio_pack_cigar16(*args, off.ctypes.data, out.ctypes.data if out.size else None, threads) != 0:
        raise ValueError("pack_cigar16: bad arguments")
    return out, off


def unpack_cigar16(cigar16, off16, n_cigar):
    """Inverse of pack_cigar16 in numpy/Python (tests): list of uint32 arrays, one per record."""
    out = []
    for i in range(len(n_cigar)):
        w = cigar16[int(off16[i]):int(off16[i + 1])].astype(np.uint64)
        ops = []; acc = 0
        for x in w.tolist():
            if x & 15 == 15:
                acc = (acc << 12) | (x >> 4)
            else:
                ops.append((((acc << 12) | (x >> 4)) << 4) | (x & 15)); acc = 0
        assert len(ops) == int(n_cigar[i])
        out.append(np.asarray(ops, dtype=np.uint32))
    return out


def pack_cigar8(batch, threads: int = 0):
    """(cigar8, cigar8_off) of a batch: the 8-bit packed CIGAR stream of include/svimgpu.h, built by csrc_host/bamio.cpp."""
    lib = _bamio_lib()
    threads = threads or (os.cpu_count() or 8)
    n = batch.n
    off = np.zeros(n + 1, dtype=np.uint64)
    args = (n, batch.n_cigar.ctypes.data, batch.cigar_off.ctypes.data, batch.cigar.ctypes.data if batch.cigar.size else None)
    if lib.bamio_pack_cigar8(*args, off.ctypes.data, None, threads) != 0:
        raise ValueError("pack_cigar8: bad arguments")
    out = np.empty(int(off[n]), dtype=np.uint8)
    if lib.bamio_pack_cigar8(*args, off.ctypes.data, out.ctypes.data if out.size else None, threads) != 0:
        raise ValueError("pack_cigar8: bad arguments")
    return out, off


def unpack_cigar8(cigar8, off8, n_cigar):
    """Inverse of pack_cigar8 in numpy/Python (tests): list of uint32 arrays, one per record."""
    out = []
    for i in range(len(n_cigar)):
        ops = []; acc = 0
        for x in cigar8[int(off8[i]):int(off8[i + 1])].tolist():
            if x & 15 == 15:
                acc = (acc << 4) | (x >> 4)
            else:
                ops.append((((acc << 4) | (x >> 4)) << 4) | (x & 15)); acc = 0
        assert len(ops) == int(n_cigar[i]) and acc == 0
        out.append(np.asarray(ops, dtype=np.uint32))
    return out


def bgzf_layout(path: str):
    """(block table as BGZF_BLOCK_DTYPE[], offset of the first alignment record in the inflated stream, contig names, contig
    lengths, sort order) from the host-side index pass of csrc_host/bamio.cpp."""
    import ctypes as C
    lib = _bamio_lib()
    inf = (C.c_int64 * 10)()
    err = C.create_string_buffer(256)
    h = lib.bamio_open(path.encode(), 1, inf, err, 256)
    if not h:
        raise ValueError("bgzf_layout(%s): %s" % (path, err.value.decode()))
    try:
        fp, fb, nb, first = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
        lib.bamio_layout(h, C.byref(fp), C.byref(fb), C.byref(nb), C.byref(first))
        blocks = np.zeros(nb.value, dtype=BGZF_BLOCK_DTYPE)
        if nb.value:
            lib.bamio_blocks(h, blocks.ctypes.data)
        n_contigs = C.cast(inf, C.POINTER(C.c_int32))[16]          # bamio_info: 8 int64, then n_contigs, sorted_coordinate
        names_buf = C.create_string_buffer(max(1, 256 * n_contigs + 64))
        lengths = np.zeros(max(1, n_contigs), dtype=np.int64)
        so = C.create_string_buffer(16)
        lib.bamio_header(h, names_buf, len(names_buf), lengths.ctypes.data, so)
        names = [x.decode("ascii") for x in names_buf.raw.split(b"\x00")[:n_contigs]]
        return blocks, first.value, names, lengths[:n_contigs], so.value.decode() or "unknown"
    finally:
        lib.bamio_close(h)


#: how often the GPU BAM decoder was used / declined a file and the host decoder took over (never silent: see decode_bam_resident)
BAM_DECODE_COUNTS = {"gpu": 0, "host_fallback": 0}


class ResidentBatch:
    """What `decode_bam_resident` returns: the record buffer lives in HBM (filled by svimgpu_decode_bam), the host keeps the header
    and fetches read names only when somebody asks for them.  Quacks like AlignmentBatch where the host mirror needs it
    (contig_names, contig_lengths, n, sort_order, qnames / qname)."""

    def __init__(self, ctx, info, contig_names, contig_lengths, sort_order):
        self.ctx, self.info = ctx, info
        self.contig_names = list(contig_names); self.contig_lengths = np.asarray(contig_lengths, dtype=np.int64)
        self.n = int(info.n_records); self.sort_order = sort_order
        self._qnames = None

    @property
    def qnames(self):
        """list[str] indexed by qname_id (first-appearance numbering, like the host decoders)"""
        if self._qnames is None:
            names, off, rec, _qid = self.ctx.fetch_bam_names(self.info)
            blob = names.tobytes()
            starts = off[rec].tolist()
            self._qnames = [blob[s:blob.index(b"\x00", s)].decode("ascii") for s in starts]
        return self._qnames

    def qname(self, qid: int) -> str:
        return self.qnames[qid]

    def to_batch(self) -> AlignmentBatch:
        """D2H of everything (tests; callers that want the flattened buffer on the host)."""
        arrays, cigar, seq, sa = self.ctx.download_alignments(self.info)
        return AlignmentBatch(self.contig_names, self.contig_lengths, arrays, cigar, seq, sa, self.qnames, self.sort_order)


def decode_bam_resident(path: str, ctx=None, stats: dict = None, fallback: bool = True):
    """BAM file -> record buffer resident in HBM, decoded on the GPU (svimgpu_decode_bam: the compressed file crosses PCIe, BGZF inflate,
    record boundaries, rows, blobs and read-name ids run on the device).  Returns a ResidentBatch; svimgpu_collect runs on it directly.
    When the device decoder declines the file (malformed stream, a speculative record boundary or a name hash that does not verify)
    the host decoder takes over — counted in BAM_DECODE_COUNTS and logged, never silent — and an AlignmentBatch is returned."""
    import logging
    import time
    from . import _lib, runtime
    ctx = ctx or runtime.context()
    t0 = time.perf_counter()
    blocks, first, names, lengths, so = bgzf_layout(path)
    t1 = time.perf_counter()
    raw = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else np.zeros(0, np.uint8)
    try:
        info = ctx.decode_bam(raw, blocks, first, len(names))
    except _lib.SvimGpuError as e:
        if not fallback or e.code != -5:
            raise
        BAM_DECODE_COUNTS["host_fallback"] += 1
        logging.warning("GPU BAM decoder declined %s (%s): decoding on the host", path, e)
        return read_bam_native(path, pack_cigar=8)
    BAM_DECODE_COUNTS["gpu"] += 1
    if stats is not None:
        stats.update({k: v for k, v in ctx.timings().items() if k.startswith("bam_")})
        stats.update(index_s=t1 - t0, decode_s=time.perf_counter() - t1, bam_bytes=int(raw.size), inflated_bytes=int(info.inflated_bytes))
    return ResidentBatch(ctx, info, names, lengths, so)


def read_bam_gpu(path: str, device: int = 0, stats: dict = None) -> AlignmentBatch:
    """BAM -> AlignmentBatch with the BGZF blocks inflated and the records parsed on the GPU, then copied back.
    Same result as read_bam_native (tests/test_gpu_bam.py); raises when the GPU path declines the file."""
    from . import _lib
    ctx = _lib.Context(device=device)
    try:
        return decode_bam_resident(path, ctx, stats, fallback=False).to_batch()
    finally:
        ctx.close()


def write_bam_native(path: str, batch: AlignmentBatch, level: int = 1, threads: int = 0):
    """SoA -> BAM with parallel BGZF compression (csrc_host/bamio.cpp::bamio_write)."""
    import ctypes as C
    lib = _bamio_lib()

    class In(C.Structure):
        _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off", "sa_off",
                                                                   "sa_len", "qname_id", "cigar", "seq", "sa", "qnames", "qname_off")] + \
                   [("n_contigs", C.c_int32), ("contig_names", C.c_char_p), ("contig_len", C.c_void_p), ("sort_order", C.c_char_p)]
    qblob = qoff = None
    if batch.qnames is not None:
        enc = [q.encode("ascii") for q in batch.qnames]
        qoff = np.zeros(len(enc) + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=qoff[1:])
        qblob = np.frombuffer(b"".join(enc) + b"\x00", dtype=np.uint8)
    names = b"\x00".join(n.encode("ascii") for n in batch.contig_names) + b"\x00"
    clen = np.ascontiguousarray(batch.contig_lengths, dtype=np.int64)
    sa = batch.sa if batch.sa.size else np.zeros(1, np.uint8)
    arg = In(batch.n, *[getattr(batch, k).ctypes.data for k in ("tid", "pos", "flag", "mapq", "n_cigar", "cigar_off", "l_seq", "seq_off", "sa_off",
                                                               "sa_len", "qname_id")], batch.cigar.ctypes.data, batch.seq.ctypes.data, sa.ctypes.data,
             qblob.ctypes.data if qblob is not None else None, qoff.ctypes.data if qoff is not None else None,
             len(batch.contig_names), names, clen.ctypes.data, batch.sort_order.encode("ascii"))
    lib.bamio_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    rc = lib.bamio_write(path.encode(), C.byref(arg), level, threads or min(32, os.cpu_count() or 1))
    if rc != 0:
        raise OSError("bamio_write(%s) failed with %d" % (path, rc))


def read_bam(path: str) -> AlignmentBatch:
    """Native multi-threaded reader when g++/zlib are available, else the pure-Python one below."""
    import subprocess
    try:
        return read_bam_native(path, pack_cigar=8)      # the batch is headed for the GPU: its CIGAR crosses PCIe as the 8-bit stream
    except (OSError, ImportError, FileNotFoundError, subprocess.SubprocessError) as e:     # toolchain missing or the build failed: fall back to the Python decoder
        if isinstance(e, FileNotFoundError) and not os.path.exists(path):
            raise
        return read_bam_python(path)


def read_bam_python(path: str) -> AlignmentBatch:
    data = _bgzf_inflate_all(path)
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack("<i", data[4:8])[0]
    text = data[8:8 + l_text].split(b"\x00")[0].decode("ascii", "replace")
    _, _, so = _parse_sam_header(text.split("\n"))
    o = 8 + l_text
    n_ref = struct.unpack("<i", data[o:o + 4])[0]; o += 4
    names, lengths = [], []
    for _ in range(n_ref):
        l_name = struct.unpack("<i", data[o:o + 4])[0]; o += 4
        names.append(data[o:o + l_name - 1].decode("ascii")); o += l_name
        lengths.append(struct.unpack("<i", data[o:o + 4])[0]); o += 4
    b = BatchBuilder(names, lengths, so)
    n = len(data)
    while o + 4 <= n:
        bs = struct.unpack("<i", data[o:o + 4])[0]; o += 4
        rec = data[o:o + bs]; o += bs
        (tid, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nt, _np, _tl) = struct.unpack("<iiBBHHHiiii", rec[:32])
        p = 32
        qname = rec[p:p + l_rn - 1].decode("ascii"); p += l_rn
        cigar = np.frombuffer(rec, dtype="<u4", count=n_cig, offset=p).copy(); p += 4 * n_cig
        nb = (l_seq + 1) // 2
        packed = np.frombuffer(rec, dtype=np.uint8, count=nb, offset=p).copy(); p += nb
        p += l_seq  # qual
        sa, cg = _aux_find(rec[p:])
        if cg is not None and n_cig > 0 and tid >= 0 and pos >= 0 and (int(cigar[0]) & 15) == 4 and (int(cigar[0]) >> 4) == l_seq:
            cigar = cg                          # the real CIGAR of a record with more than 65535 operations (htslib bam_tag2cigar)
        b.add(qname, flag, tid, pos, mapq, cigar, None, sa.decode("ascii") if sa else None,
              packed_seq=packed, l_seq=l_seq)
    return b.finish()


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_bam(path: str, batch: AlignmentBatch, level: int = 1):
    """Write the batch as a BAM file (one BGZF block per <=64 KiB of payload)."""
    out = bytearray()
    text = "@HD\tVN:1.6\tSO:%s\n" % batch.sort_order
    for nm, ln in zip(batch.contig_names, batch.contig_lengths):
        text += "@SQ\tSN:%s\tLN:%d\n" % (nm, ln)
    tb = text.encode("ascii")
    out += b"BAM\x01" + struct.pack("<i", len(tb)) + tb + struct.pack("<i", len(batch.contig_names))
    for nm, ln in zip(batch.contig_names, batch.contig_lengths):
        nb = nm.encode("ascii") + b"\x00"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(ln))
    ref_adv = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], dtype=np.int64)
    for i in range(batch.n):
        o = int(batch.cigar_off[i]); nc = int(batch.n_cigar[i])
        cig = batch.cigar[o:o + nc]
        rlen = int(((cig >> 4).astype(np.int64) * ref_adv[cig & 15]).sum()) if nc else 0
        pos = int(batch.pos[i])
        qn = batch.qname(int(batch.qname_id[i])).encode("ascii") + b"\x00"
        l_seq = int(batch.l_seq[i]); so = int(batch.seq_off[i])
        core = cig if nc <= 65535 else _long_cigar_core(l_seq, rlen)
        rec = struct.pack("<iiBBHHHiiii", int(batch.tid[i]), pos, len(qn), int(batch.mapq[i]),
                          _reg2bin(max(pos, 0), max(pos, 0) + max(rlen, 1)), len(core), int(batch.flag[i]), l_seq, -1, -1, 0)
        rec += qn + core.astype("<u4").tobytes() + batch.seq[so:so + (l_seq + 1) // 2].tobytes() + b"\xff" * l_seq
        sa = batch.sa_tag(i)
        if sa:
            rec += b"SAZ" + sa.encode("ascii") + b"\x00"
        if nc > 65535:
            rec += b"CGBI" + struct.pack("<I", nc) + cig.astype("<u4").tobytes()
        out += struct.pack("<i", len(rec)) + rec
    with open(path, "wb") as fh:
        view = memoryview(out)
        for s in range(0, len(out), 0xff00):
            chunk = bytes(view[s:s + 0xff00])
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            cdata = co.compress(chunk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
                     + struct.pack("<H", len(cdata) + 25) + cdata
                     + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        fh.write(_BGZF_EOF)


def read_alignments(path: str) -> AlignmentBatch:
    with open(path, "rb") as fh:
        magic = fh.read(4)
    if magic[:2] == b"\x1f\x8b":
        return read_bam(path)
    return read_sam(path)


# --------------------------------------------------------------------------
# FASTA
# --------------------------------------------------------------------------
class Genome:
    """Whole reference in memory: one uint8 blob + per-contig offsets.

    Stands in for `pysam.FastaFile` (SVIM_clustering.py:377): `fetch(contig,
    start, end)` returns `seq[start:end]`, clamped at the contig end, case
    preserved (htslib faidx semantics; parity unpinned – no reference test
    fetches from a FASTA)."""

    def __init__(self, names: List[str], seqs: List[np.ndarray]):
        self.names = list(names)
        self.lengths = np.array([len(s) for s in seqs], dtype=np.int64)
        self.offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(self.lengths, out=self.offsets[1:])
        self.blob = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)
        self._idx = {n: i for i, n in enumerate(self.names)}

    @classmethod
    def from_fasta(cls, path: str) -> "Genome":
        names, seqs, cur = [], [], []
        with open(path, "rb") as fh:
            for ln in fh:
                if ln.startswith(b">"):
                    if names:
                        seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
                    names.append(ln[1:].split()[0].decode("ascii")); cur = []
                else:
                    cur.append(ln.strip())
        if names:
            seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
        return cls(names, seqs)

    def fetch(self, contig: str, start: int, end: int) -> str:
        i = self._idx[contig]
        L = int(self.lengths[i])
        start = max(0, min(start, L)); end = max(start, min(end, L))
        o = int(self.offsets[i])
        return self.blob[o + start:o + end].tobytes().decode("ascii")

    def write_fasta(self, path: str, width: int = 60):
        with open(path, "wb") as fh, open(path + ".fai", "w") as fai:
            for i, nm in enumerate(self.names):
                fh.write(b">" + nm.encode() + b"\n")
                off = fh.tell()
                s = self.blob[self.offsets[i]:self.offsets[i + 1]].tobytes()
                for k in range(0, len(s), width):
                    fh.write(s[k:k + width] + b"\n")
                fai.write("%s\t%d\t%d\t%d\t%d\n" % (nm, len(s), off, width, width + 1))
