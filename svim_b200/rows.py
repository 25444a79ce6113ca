"""Device records -> the reference's value tuples, and the parity bookkeeping built on them.

`sig_rows` / `cluster_rows` put svim_sig / svim_cluster records into the row shape of the committed golden vectors
(tests/golden, tools/make_golden.py) — what the reference's Signature* / SignatureCluster* objects hold (SVSignature.py:3-310).
`prefix_parity` and `result_digest` are what bench.py prints under "parity": the CUDA path against the CPU restatement on the
genomic prefix the CPU baseline walked anyway, and a digest that every rank of a multi-GPU run must agree on.  This module
never computes a result itself; the CPU side of the comparison is passed in by the caller (bench.py / tests)."""
import hashlib
import math

import numpy as np

from . import _lib

SIG_FIELDS = ("type", "contig", "start", "end", "contig2", "pos", "dir1", "dir2", "direction", "copies", "fully_covered",
              "signature", "read", "sequence")
TYPES_RETURN_ORDER = ("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND")


def sig_rows(sigs, ins, batch):
    """svim_sig records -> golden-style rows (list per signature, field order of oracle.Sig.__slots__)."""
    names = batch.contig_names
    blob = ins.tobytes() if hasattr(ins, "tobytes") else bytes(ins)
    rows = []
    for s in sigs:
        t = _lib.TYPE_NAMES[s["type"]]
        fl = int(s["flags"])
        d = dict.fromkeys(SIG_FIELDS)
        d.update(type=t, contig=names[s["contig1"]], start=int(s["start"]), end=int(s["end"]),
                 signature="suppl" if fl & 1 else "cigar", read=batch.qname(int(s["qname_id"])))
        if t == "INS":
            d["sequence"] = blob[int(s["seq_off"]):int(s["seq_off"]) + int(s["seq_len"])].decode("ascii")
        elif t == "INV":
            d["direction"] = _lib.INV_DIRECTIONS[(fl >> 4) & 7]
        elif t == "DUP_TAN":
            d.update(copies=int(s["copies"]), fully_covered=bool(fl & 2))
        elif t == "DUP_INT":
            d.update(contig2=names[s["contig2"]], pos=int(s["pos"]))
        elif t == "BND":
            d.update(contig2=names[s["contig2"]], pos=int(s["pos"]), dir1="rev" if fl & 4 else "fwd", dir2="rev" if fl & 8 else "fwd")
        rows.append([d[f] for f in SIG_FIELDS])
    return rows


def cluster_row(c, mem, first_row):
    t = _lib.TYPE_NAMES[c["type"]]
    sd_span = None if math.isnan(c["std_span"]) else float(c["std_span"])
    sd_pos = None if math.isnan(c["std_pos"]) else float(c["std_pos"])
    row = [t, first_row[1], int(c["start"]), int(c["end"]), None, None, None, float(c["score"]), int(c["size"]), sd_span, sd_pos, None, None, mem]
    if t == "DUP_TAN":
        row[4:7] = [first_row[1], int(c["dest_start"]), int(c["dest_end"])]
    elif t in ("DUP_INT", "BND"):
        row[4:7] = [first_row[4], int(c["dest_start"]), int(c["dest_end"])]
        if t == "BND":
            row[11:13] = ["rev" if c["dir1_rev"] else "fwd", "rev" if c["dir2_rev"] else "fwd"]
    return row


def cluster_rows(clusters, members, sig_rows_list):
    """svim_cluster records -> {type: [golden-style cluster rows]}"""
    out = {t: [] for t in TYPES_RETURN_ORDER}
    mem = members.tolist()
    for c in clusters:
        ms = mem[int(c["member_off"]):int(c["member_off"]) + int(c["size"])]
        row = cluster_row(c, ms, sig_rows_list[ms[0]])
        out[row[0]].append(row)
    return out


def cluster_rows_differ(a, b, float_tol=1e-6):
    """Bit-exact on membership, order and integer coordinates; score / std within `float_tol` (BASELINE.json north_star)."""
    if a[:7] != b[:7] or a[8] != b[8] or a[11:] != b[11:]:
        return True
    for i in (7, 9, 10):
        if a[i] is None or b[i] is None:
            if not (a[i] is None and b[i] is None):
                return True
        elif abs(a[i] - b[i]) > float_tol * max(1.0, abs(b[i])):
            return True
    return False


def result_digest(clusters, members, sigs=None):
    """64-bit digest of the result bytes (cluster records in list order + member indices [+ signature records])."""
    h = hashlib.blake2b(digest_size=8)
    h.update(np.ascontiguousarray(clusters).tobytes()); h.update(np.ascontiguousarray(members).tobytes())
    if sigs is not None:
        h.update(np.ascontiguousarray(sigs).tobytes())
    return h.hexdigest()


def prefix_parity(batch, n_prefix, sigs, ins, clusters, members, order, part_off, want_sig_rows, want_clusters, float_tol=1e-6):
    """CUDA results of the WHOLE input against CPU results of its first `n_prefix` records.

    Signatures: every record emits on its own (SVIM_COLLECT.py:142-161), so the signatures of records < n_prefix must be the
    CPU list, in order.  Clusters: a partition (SVIM_clustering.py:17-29) is comparable when all of its signatures come from the
    prefix and the sampling stream (seed(1524) per type, consumed by partitions above 100, :129-134) has seen the same
    partitions before it; its clusters must be the CPU clusters with the same first member — members, coordinates, score,
    std — and appear in the same relative order in the per-type output lists.
    `order` / `part_off`: the device's sorted order and partition offsets (svimgpu_fetch_partitions).
    `want_clusters`: {type: [golden-style rows with member indices into the CPU signature list]}."""
    aln = np.asarray(sigs["aln_idx"])
    is_pre = aln < n_prefix
    pre_idx = np.nonzero(is_pre)[0]
    out = {"prefix_records": int(n_prefix), "signatures_compared": int(len(want_sig_rows)), "clusters_compared": 0, "partitions_compared": 0,
           "sampled_partitions_compared": 0, "mismatches": 0}
    got_rows = sig_rows(sigs[pre_idx], ins, batch)
    sig_bad = int(len(got_rows) != len(want_sig_rows)) + sum(1 for a, b in zip(got_rows, want_sig_rows) if a != b)
    out["signature_mismatches"] = sig_bad
    out["mismatches"] += sig_bad
    if sig_bad:
        return out
    # device partitions: which are pure prefix, and up to where the sampling stream is the CPU's
    n_part = len(part_off) - 1
    sizes = np.diff(part_off.astype(np.int64))
    part_of_sorted = np.repeat(np.arange(n_part), sizes)
    part_of_sig = np.empty(len(sigs), dtype=np.int64); part_of_sig[order] = part_of_sorted
    foreign_in_part = np.zeros(n_part, dtype=np.int64)
    np.add.at(foreign_in_part, part_of_sig[~is_pre], 1)
    pure = foreign_in_part == 0
    ptype = np.asarray(sigs["type"])[order[part_off[:-1]]]
    comparable = pure.copy()
    for t in range(6):
        sel = np.nonzero(ptype == t)[0]
        if len(sel) == 0:
            continue
        impure = sel[~pure[sel]]
        if len(impure):                      # past the first partition that differs, sampled partitions see another stream
            late = sel[sel > impure[0]]
            comparable[late[sizes[late] > 100]] = False
    out["partitions_compared"] = int(comparable.sum())
    out["sampled_partitions_compared"] = int((comparable & (sizes > 100)).sum())
    # device clusters of comparable partitions, keyed by first member (a global signature index)
    cpu_to_dev = pre_idx                      # CPU signature i is device signature pre_idx[i]
    mem = np.asarray(members)
    first = mem[np.asarray(clusters["member_off"], dtype=np.int64)]
    dev_part = part_of_sig[first]
    dev_by_first = {}
    pos_in_type = {}
    counters = {}
    for ci in range(len(clusters)):
        t = int(clusters["type"][ci]); k = counters.get(t, 0); counters[t] = k + 1
        if comparable[dev_part[ci]]:
            dev_by_first[int(first[ci])] = ci; pos_in_type[ci] = k
    seen = 0
    bad = 0
    for tname in TYPES_RETURN_ORDER:
        last_pos = -1
        for row in want_clusters.get(tname, []):
            gm = [int(cpu_to_dev[i]) for i in row[13]]
            if not comparable[part_of_sig[gm[0]]]:
                continue
            seen += 1
            ci = dev_by_first.get(gm[0])
            if ci is None:
                bad += 1; continue
            c = clusters[ci]
            ms = mem[int(c["member_off"]):int(c["member_off"]) + int(c["size"])].tolist()
            got = cluster_row(c, ms, got_rows_by_global(ms[0], pre_idx, got_rows))
            want = list(row); want[13] = gm
            if cluster_rows_differ(got, want, float_tol) or pos_in_type[ci] <= last_pos:
                bad += 1
            last_pos = max(last_pos, pos_in_type[ci])
    # a comparable device cluster the CPU did not produce is a mismatch too
    bad += max(0, len(dev_by_first) - seen)
    out["clusters_compared"] = seen
    out["cluster_mismatches"] = bad
    out["mismatches"] += bad
    return out


def got_rows_by_global(g, pre_idx, got_rows):
    return got_rows[int(np.searchsorted(pre_idx, g))]
