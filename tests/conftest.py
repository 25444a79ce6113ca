import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ("mini_indel", "mini_mixed", "mini_mixed_allbnds", "mini_ins", "mini_hotspot", "chimeric_kat")
GOLDEN_QUERYSORTED = ("mini_mixed_querysorted",)
GOLDEN_GENOTYPE = ("geno_mini_indel", "geno_mini_mixed", "geno_mini_hotspot", "geno_deep")


def querysort_order(batch):
    """Deterministic read-name order for the query-sorted fixtures: records of a read become adjacent, in a
    pseudo-random order inside the read (the primary is not always first)."""
    idx = np.arange(batch.n, dtype=np.int64)
    return np.argsort(batch.qname_id.astype(np.int64) * 8 + (idx * 2654435761) % 7, kind="stable")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "gpu_next: experimental GPU code that has not run on a GPU yet; only with SVIM_RUN_NEXT=1")


def load_golden(name):
    """-> (AlignmentBatch, Genome, expected dict) from the committed fixtures."""
    from svim_b200.records import AlignmentBatch
    from svim_b200.io import Genome
    with gzip.open(os.path.join(GOLDEN, name + ".golden.json.gz"), "rt") as fh:
        exp = json.load(fh)
    z = np.load(os.path.join(GOLDEN, exp["input"]))
    names = [str(x) for x in z["contig_names"]]
    arrays = {f: z[f] for f, _ in AlignmentBatch.FIELDS}
    qnames = [str(x) for x in z["qnames"]] if "qnames" in z.files else None
    batch = AlignmentBatch(names, z["contig_lengths"], arrays, z["cigar"], z["seq"], z["sa"], qnames, "coordinate")
    if exp.get("derive") == "querysorted":
        batch = batch.take(querysort_order(batch), "queryname")
    blob = z["genome_blob"]
    offs = np.concatenate([[0], np.cumsum(z["contig_lengths"])]).astype(np.int64)
    genome = Genome(names, [blob[offs[i]:offs[i + 1]] for i in range(len(names))])
    return batch, genome, exp


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def gpu_ctx():
    from svim_b200 import _lib
    ctx = _lib.Context()
    yield ctx
    ctx.close()
