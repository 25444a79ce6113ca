"""Helpers shared by the -m gpu tests: run the CUDA path through the C ABI and put its
outputs in the same shape as the committed golden vectors / the oracle."""
import math

import numpy as np

from svim_b200 import _lib

from svim_b200.rows import SIG_FIELDS, TYPES_RETURN_ORDER, sig_rows, cluster_rows  # noqa: F401,E402


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
