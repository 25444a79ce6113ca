"""Helpers shared by the -m gpu tests: run the CUDA path through the C ABI and put its
outputs in the same shape as the committed golden vectors / the oracle."""
import math

import numpy as np

from svim_b200 import _lib

SIG_FIELDS = ("type", "contig", "start", "end", "contig2", "pos", "dir1", "dir2", "direction", "copies", "fully_covered",
              "signature", "read", "sequence")
TYPES_RETURN_ORDER = ("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND")


def sig_rows(sigs, ins, batch):
    """svim_sig records -> golden-style rows (list per signature, field order of oracle.Sig.__slots__)."""
    names = batch.contig_names
    blob = ins.tobytes()
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


def cluster_rows(clusters, members, sig_rows_list):
    """svim_cluster records -> {type: [golden-style cluster rows]}"""
    out = {t: [] for t in TYPES_RETURN_ORDER}
    mem = members.tolist()
    for c in clusters:
        t = _lib.TYPE_NAMES[c["type"]]
        ms = mem[int(c["member_off"]):int(c["member_off"]) + int(c["size"])]
        first = sig_rows_list[ms[0]]
        sd_span = None if math.isnan(c["std_span"]) else float(c["std_span"])
        sd_pos = None if math.isnan(c["std_pos"]) else float(c["std_pos"])
        row = [t, first[1], int(c["start"]), int(c["end"]), None, None, None, float(c["score"]), int(c["size"]), sd_span, sd_pos, None, None, ms]
        if t == "DUP_TAN":
            row[4:7] = [first[1], int(c["dest_start"]), int(c["dest_end"])]
        elif t in ("DUP_INT", "BND"):
            row[4:7] = [first[4], int(c["dest_start"]), int(c["dest_end"])]
            if t == "BND":
                row[11:13] = ["rev" if c["dir1_rev"] else "fwd", "rev" if c["dir2_rev"] else "fwd"]
        out[t].append(row)
    return out


def assert_clusters_equal(got, want, float_tol=1e-6):
    """Bit-exact on membership, order and integer coordinates; score/std within `float_tol`
    (BASELINE.json north_star: "float distances within 1e-6")."""
    for t in TYPES_RETURN_ORDER:
        g, w = got[t], want[t]
        assert len(g) == len(w), (t, len(g), len(w))
        for k, (a, b) in enumerate(zip(g, w)):
            assert a[:7] == b[:7], (t, k, a[:7], b[:7])
            assert a[8] == b[8] and a[11:] == b[11:], (t, k, a, b)
            for i in (7, 9, 10):
                if a[i] is None or b[i] is None:
                    assert a[i] is None and b[i] is None, (t, k, i, a, b)
                else:
                    assert abs(a[i] - b[i]) <= float_tol * max(1.0, abs(b[i])), (t, k, i, a[i], b[i])


def run_gpu(ctx, batch, genome, overrides, which=0, querysorted=False):
    """COLLECT + CLUSTER through the C ABI -> (sig rows, twin rows, clusters dict for `which`)."""
    from svim_b200 import runtime
    ctx.set_params(_lib.Params.from_options(None, **overrides))
    ctx.set_contigs(batch.contig_names)
    st = ctx.collect_host_querysorted(batch) if querysorted else ctx.collect_host(batch)
    assert st.n_data_errors == 0
    sigs, ins = ctx.fetch_signatures(0, st)
    rows = sig_rows(sigs, ins, batch)
    trows = []
    if st.n_twin_signatures:
        tsigs, tins = ctx.fetch_signatures(1, st)
        trows = sig_rows(tsigs, tins, batch)
    ctx.genome_key = None
    runtime.ensure_genome(ctx, genome, batch.contig_names)
    ctx.use_collected(which)
    cst, clusters, members = ctx.cluster()
    return rows, trows, cluster_rows(clusters, members, rows if which == 0 else trows), st, cst
