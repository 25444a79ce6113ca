#!/usr/bin/env python3
"""Why does the 8-GPU e2e step spend 275 ms in H2D against 112 ms on one GPU?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/h2d_topo_check.py

Every rank uploads the same pinned 2 GiB record buffer (a) alone, one rank after the other, (b) all ranks at once, (c) all at once
after re-allocating the buffer with the process confined to the CPUs local to its GPU (`local_cpulist` of the PCI device), and rank 0
prints the GB/s table next to `nvidia-smi topo -m`.  Shared PCIe uplinks show up as (b) << (a) for pairs of ranks regardless of (c);
NUMA placement shows up as (c) >> (b).  GPU box only; a diagnosis tool, not part of the measured contract."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np


def make_batch(words):
    from svim_b200.records import AlignmentBatch
    arrays = {name: np.zeros(1, dtype=dt) for name, dt in AlignmentBatch.FIELDS}
    arrays["n_cigar"][0] = words
    return AlignmentBatch(["c"], [1], arrays, np.ones(words, dtype=np.uint32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), None, "coordinate")


def main():
    import torch.distributed as dist
    from svim_b200 import _lib
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist.init_process_group("gloo")
    ctx = _lib.Context(device=local)
    words = (2 << 30) // 4

    def timed_upload(batch):
        ctx.upload(batch)
        return batch.cigar.nbytes / (ctx.timings()["h2d_alignments"] * 1e-3) / 1e9

    def run(batch):
        ctx.pin(batch.cigar)
        timed_upload(batch)                                   # warm-up
        alone = None
        for r in range(world):
            dist.barrier()
            if r == rank:
                alone = timed_upload(batch)
        dist.barrier()
        together = timed_upload(batch)
        ctx.unpin(batch.cigar)
        return alone, together

    a, b = run(make_batch(words))
    cpus = None
    try:
        bus = ctx.pci_bus_id().lower()
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        ids = []
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, ids)
    except Exception as e:
        cpus = "unavailable (%s)" % e
    _a2, c = run(make_batch(words))                         # first-touched under the local CPU list
    rows = [None] * world
    dist.all_gather_object(rows, (rank, ctx.pci_bus_id(), cpus, a, b, c))
    if rank == 0:
        print("rank  pci              local cpus            alone GB/s  all-at-once  all-at-once, NUMA-local")
        for r in sorted(rows):
            print("%4d  %-15s  %-20s  %9.1f  %11.1f  %11.1f" % r)
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
        except Exception:
            pass
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
