#!/usr/bin/env python3
"""bench.py — alignments/sec through COLLECT+CLUSTER (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # N>1: launched by torchrun, one rank per GPU
    python bench.py --impl reference ...                    # the CPU path (oracle port of the reference) on a bounded sample

A "step" is one pass of the hot path over one batch of synthetic alignments: BASELINE.json
configs[1] (1 contig, 500 k alignments, 15 kb CLR-like reads) per GPU; with N GPUs every rank
holds one such contig (weak scaling), the ranks exchange signature records once after COLLECT
and cluster records once at the end (NCCL all-gather-v).

Printed JSON (rank 0, one line):
  value     whole-job alignments/s with the record buffer already resident in HBM
            (CUDA events around collect -> [exchange] -> cluster, max over ranks)
  e2e       same metric through the C ABI with HOST buffers: H2D of the record buffer and D2H of
            the signature + cluster records inside the timed region, every step
  roofline  the CIGAR-scan kernel: algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline  oracle port of the reference, 1 host thread, bounded genomic slice of the same input
  clocks    SM clock / throttle reasons sampled through NVML every 2 ms during the timed region (also under e2e)
  genotype  (N=1, outside the timed step) GENOTYPE of the step's DEL+INS candidates through svimgpu_genotype: candidates/s, the
            one-off preparation (k_ref_end streams every CIGAR once more) and the oracle on a bounded sample beside it
  e2e_from_bam  (--with-bam) the same path starting from a BAM file: native streaming decode + pageable H2D + COLLECT + CLUSTER
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun pins OMP_NUM_THREADS=1 in every rank; the synthetic-input generator (host C++/OpenMP) should use its share of the host
_world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // _world))

import numpy as np  # noqa: E402

METRIC = "alignments/sec through COLLECT+CLUSTER"
UNIT = "alignments/s"


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_rank_input(workload, scale, rank, world):
    """Rank r's shard: one contig 'chr<r+1>' of the workload (records are contiguous in coordinate order)."""
    from svim_b200 import synth
    c = dict(synth.CONFIGS[workload])
    G = int(c.pop("genome") * scale); reads = max(10, int(c.pop("reads") * scale)); seed = c.pop("seed") + 100 * rank
    c.pop("contigs")
    hotspots = int(round(c.pop("hotspots", 0) * scale)) if "hotspots" in c else 0
    plant_kw = {k: c.pop(k) for k in ("spacing", "mix", "ins_size_uniform", "hotspot_svs") if k in c}
    names = ["chr%d" % (r + 1) for r in range(world)]
    lengths = [G] * world
    svs, alleles = synth.plant_svs([G], seed, hotspots=hotspots, **plant_kw)
    batch = synth.generate([names[rank]], [G], reads, seed, svs, alleles, **c)
    batch.contig_names = names; batch.contig_lengths = np.asarray(lengths, dtype=np.int64)
    batch._tid_of = {n: i for i, n in enumerate(names)}
    batch.tid[batch.tid >= 0] = rank
    batch.qname_id += np.uint32(rank << 26)
    genomes = [synth.random_genome([names[r]], [G], 1524 + synth.CONFIGS[workload]["seed"] + 100 * r) for r in range(world)]
    from svim_b200.io import Genome
    genome = Genome(names, [g.blob for g in genomes])
    return batch, genome


def make_sharded_input(workload, scale, rank, world):
    """--shard records: ONE coordinate-sorted input of the workload (config4: 24 contigs chr1..chr22, chrX, chrY), cut into `world`
    contiguous record ranges balanced by CIGAR volume (svim_b200.parallel.shard_ranges); this rank materialises only its range.
    Returns (batch, genome, first record index, total records)."""
    from svim_b200 import synth
    names, lengths, reads, seed, plant_kw, gen_kw = synth.config_layout(workload, scale)
    svs, alleles = synth.plant_svs(lengths, seed, **plant_kw)
    batch, lo, total = synth.generate_shard(names, lengths, reads, seed, svs, alleles, rank, world, **gen_kw)
    genome = synth.random_genome(names, lengths, 1524 + seed)
    return batch, genome, lo, total


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line).

    The timed region is a few hundred milliseconds, shorter than an `nvidia-smi` start-up, so the samples come from NVML
    in-process (nvidia-ml-py), every 2 ms on a thread, for the GPU this rank drives (matched by PCI bus id, so
    CUDA_VISIBLE_DEVICES does not matter).  Fallback: an `nvidia-smi -lms` child started before the warm-up."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, device, ctx=None):
        self.device = device; self.proc = None; self.rows = []; self.nvml = None; self.handle = None
        self.samples = []; self.mask = 0; self.max_mhz = None; self._stop = threading.Event(); self.t = None; self.t_start = None
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = ctx.pci_bus_id() if ctx is not None else None
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()) if bus else pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
            try:        # fallback, started early: nvidia-smi needs about a second before its first line
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read_smi, daemon=True); self.t.start()
            except OSError:
                self.proc = None

    def _read_smi(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in ln.split(",")]))

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(reasons(self.handle))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        self.t_start = time.perf_counter()
        if self.nvml:
            self._stop.clear()
            self.t = threading.Thread(target=self._poll, daemon=True); self.t.start()

    def stop(self):
        t_stop = time.perf_counter()
        if self.nvml:
            self._stop.set(); self.t.join(timeout=2)
            sm = self.samples
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": [name for bit, name in self.BITS if self.mask & bit], "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"], "samples": 0}
        time.sleep(0.12)        # let the line that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        t0 = self.t_start if self.t_start is not None else 0.0
        rows = [r for t, r in self.rows if t0 <= t <= t_stop + 0.12 and len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for t, r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(batch, genome, target_seconds=15.0, keep=None):
    """Oracle port of the reference on a contiguous genomic prefix of the same input, 1 host thread.
    keep: dict that receives the prefix length and the oracle's signature / cluster rows (the parity check reuses them)."""
    from oracle import svim_oracle as orc
    p = orc.Params()
    # calibrate on 300 records, then size the slice for ~target_seconds of COLLECT (+ CLUSTER on what it emits)
    n0 = min(300, batch.n)
    t0 = time.perf_counter(); orc.collect(batch.slice(0, n0), p); dt = max(1e-4, time.perf_counter() - t0)
    n = int(min(batch.n, max(n0, 0.6 * target_seconds / dt * n0)))
    sl = batch.slice(0, n)
    t0 = time.perf_counter()
    sigs, _ = orc.collect(sl, p)
    t1 = time.perf_counter()
    res = orc.cluster(sigs, genome, p)
    t2 = time.perf_counter()
    if keep is not None:
        index_of = {id(x): i for i, x in enumerate(sigs)}
        keep["n"] = n
        keep["signatures"] = [list(x.as_tuple()) for x in sigs]
        keep["clusters"] = {name: [[c.type, c.contig, c.start, c.end, c.dest_contig, c.dest_start, c.dest_end, c.score, c.size, c.std_span, c.std_pos,
                                    c.dir1, c.dir2, [index_of[id(m)] for m in c.members]] for c in cl]
                            for name, cl in zip(("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND"), res)}
    return {"value": n / (t2 - t0), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d of %d coordinate-sorted records (%.1f s COLLECT + %.1f s CLUSTER, %d signatures); "
                      "edit distance by oracle/editdist.c Myers (edlib stand-in)" % (n, batch.n, t1 - t0, t2 - t1, len(sigs))}


def genotype_leg(ctx, batch, clusters, members, sigs, args):
    """SURVEY.md §8f rank 2, outside the timed step: GENOTYPE (SVIM_genotyping.py:34-93) of the DEL and INS candidates COMBINE makes
    1:1 from this step's clusters, on the record rows the e2e step left in HBM.  Reports the one-off preparation (k_ref_end streams
    every CIGAR again: HBM-bound), the per-call time through the C ABI from host candidate arrays, and the oracle beside it."""
    from svim_b200 import _lib
    from svim_b200.SVIM_genotyping import candidate_arrays_from_clusters
    gp = _lib.GenoParams.from_options()
    out = {"unit": "candidates/s", "types": {}}
    prep = None
    n_all = 0; ms_all = 0.0
    per_type = {}
    for tname in ("DEL", "INS"):
        cands, var_ids, _sel = candidate_arrays_from_clusters(clusters, members, sigs, _lib.TYPE_CODE[tname])
        if len(cands) == 0:
            continue
        res = ctx.genotype(_lib.TYPE_CODE[tname], gp, cands, var_ids, batch.contig_lengths)      # first call also prepares
        tm = ctx.timings()
        if prep is None:
            prep = tm.get("genotype_prepare", 0.0)
        ms = []
        for _ in range(max(3, args.steps)):
            ctx.timer_start()
            res = ctx.genotype(_lib.TYPE_CODE[tname], gp, cands, var_ids, batch.contig_lengths)
            ms.append(ctx.timer_stop())
        m = float(np.mean(ms))
        per_type[tname] = (cands, var_ids, res)
        calls = {g: int((res["genotype"] == i).sum()) for i, g in enumerate(_lib.GENOTYPES)}
        out["types"][tname] = {"candidates": int(len(cands)), "ms_per_call": round(m, 4), "records_fetched": int(res["n_fetched"].sum()),
                               "mean_ref_reads": float(res["ref_reads"].mean()), "calls": calls}
        n_all += len(cands); ms_all += m
    if not n_all:
        return None
    out["value"] = n_all / (ms_all * 1e-3)
    out["prepare_ms"] = prep
    rows_bytes = batch.n * (4 + 8 + 2 + 4 + 4) + 4 * int(batch.n_cigar.sum(dtype=np.int64))
    peak, _src = measured_peak()
    out["prepare_note"] = ("one-off per resident buffer: k_ref_end (bam_endpos of every record, %.2f GB of CIGAR + rows) + contig row bounds + "
                           "running maximum; %.0f GB/s if all of it were k_ref_end = %.2f of the measured HBM peak"
                           % (rows_bytes / 1e9, rows_bytes / max(prep, 1e-6) / 1e6, rows_bytes / max(prep, 1e-6) / 1e6 / peak)) if prep else None
    if not args.no_cpu_baseline:
        # oracle on a bounded sample of the same candidates, checked against the GPU result on the way: a genomic prefix of the
        # records (every record a window of the sampled candidates can reach is inside it) keeps the numpy passes short
        from oracle import svim_oracle as orc
        n_sl = min(batch.n, 60000)
        sl = batch.slice(0, n_sl)
        reach = int(batch.pos[n_sl]) if n_sl < batch.n else 1 << 62          # windows must end before the first excluded record starts
        tid0 = int(batch.tid[0])
        t0 = time.perf_counter(); ends = orc.record_reference_ends(sl); t_ends = time.perf_counter() - t0
        done = 0; t_cpu = 0.0; mismatches = 0
        for tname, (cands, var_ids, res) in per_type.items():
            keep = np.nonzero((cands["tid"] == tid0) & (np.maximum(cands["start"], cands["end"]) + 1000 < reach))[0][:400]
            ocs = []
            for i in keep.tolist():
                c = cands[i]
                ids = var_ids[int(c["variant_off"]):int(c["variant_off"]) + int(c["n_variant_reads"])]
                ocs.append(orc.GenoCand(batch.contig_names[int(c["tid"])], int(c["start"]), int(c["end"]), 10, [batch.qname(int(q)) for q in ids]))
            t0 = time.perf_counter(); orc.genotype(ocs, sl, tname, orc.GenoParams(), ends); t_cpu += time.perf_counter() - t0
            for i, oc in zip(keep.tolist(), ocs):
                f = float(res["support_fraction"][i])
                got = ["." if f != f else f, _lib.GENOTYPES[int(res["genotype"][i])], int(res["ref_reads"][i]), int(res["alt_reads"][i])]
                mismatches += got != oc.result()
            done += len(ocs)
        if not done:
            return out
        out["cpu_baseline"] = {"value": done / t_cpu, "unit": "candidates/s", "cores": 1, "kind": "port",
                               "sample": "%d candidates on the first %d records (genomic prefix); oracle region fetch is a numpy mask over "
                                         "those records; + %.2f s once for their reference ends" % (done, n_sl, t_ends),
                               "mismatches_vs_gpu": int(mismatches)}
    return out


def run_reference(args):
    rank, local_rank, world = env_rank()
    if rank != 0:
        return
    batch, genome = make_rank_input(args.workload, args.scale, 0, 1)
    vals = []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_baseline(batch, genome, target_seconds=args.ref_seconds)
        if s >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    n_s = int(base["sample"].split()[1])
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1000.0 * n_s / v, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
                      "config": {"workload": workload_name(args), "note": "reference is single-threaded pure Python (README.rst:73); "
                                 "each step is a bounded sample of the workload"},
                      "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


WORKLOAD_TEXT = {"config2": "BASELINE.json configs[1]: 1 contig, 500k alignments, 15 kb CLR-like reads",
                 "config3": "BASELINE.json configs[2]: insertion-heavy, 200-5000 bp inserted sequences",
                 "config4": "BASELINE.json configs[3]: 24 contigs chr1..chr22,chrX,chrY, 5 M alignments, 30x whole-genome scale",
                 "config5": "BASELINE.json configs[4]: 1 contig at 100x, 2 M alignments, hotspot partitions up to 50 k signatures",
                 "config1": "BASELINE.json configs[0]: 1 contig, 1000 reads"}


def workload_name(args):
    if args.shard == "records":
        return "%s x%g, ONE coordinate-sorted input cut into contiguous record ranges, one per GPU (%s)" % (args.workload, args.scale, WORKLOAD_TEXT.get(args.workload, ""))
    return "%s x%g per GPU (%s)" % (args.workload, args.scale, WORKLOAD_TEXT.get(args.workload, ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink genome and read count together (testing only)")
    ap.add_argument("--shard", default="contigs", choices=["contigs", "records"],
                    help="contigs (default): every rank holds its own contig of the workload (weak scaling); records: one input of the "
                         "workload, contiguous record ranges balanced by CIGAR volume (strong scaling)")
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cigar16", action="store_true", help="e2e legs upload BAM's uint32 CIGAR words instead of a packed stream (same as --cigar-pack 32)")
    ap.add_argument("--cigar-pack", type=int, choices=[8, 16, 32], default=8, help="CIGAR stream of the e2e legs: 8 = svim_aln_soa.cigar8, 16 = cigar16, 32 = BAM's uint32 words")
    ap.add_argument("--profile-steps", action="store_true", help="only run resident steps (for ncu)")
    ap.add_argument("--no-bam", action="store_true", help="skip the e2e_from_bam leg (writes a BAM of the workload to a temporary directory)")
    ap.add_argument("--with-bam", action="store_true", help=argparse.SUPPRESS)        # round-1 spelling: the leg is on by default now
    ap.add_argument("--profile-genotype", action="store_true", help="with --profile-steps: also run the GENOTYPE leg (for ncu)")
    ap.add_argument("--scan-variant", type=int, default=None, help="0 = 128-bit LDG streaming scan, 1 = cp.async.bulk ring (default: library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3 and not args.profile_steps:
        args.warmup = 3

    rank, local_rank, world = env_rank()
    if args.scan_variant is not None:
        os.environ["SVIM_SCAN_VARIANT"] = str(args.scan_variant)
    from svim_b200 import _lib
    t_gen = time.perf_counter()
    if args.shard == "records":
        batch, genome, shard_lo, shard_total = make_sharded_input(args.workload, args.scale, rank, world)
    else:
        batch, genome = make_rank_input(args.workload, args.scale, rank, world)
    t_gen = time.perf_counter() - t_gen

    ctx = _lib.Context(device=local_rank)
    ctx.set_contigs(batch.contig_names)
    ctx.set_genome(genome)
    if world > 1:
        import torch.distributed as dist
        from svim_b200 import parallel
        dist.init_process_group("gloo")
        parallel.init_comm(ctx)
        aln_base, total_aln, _sizes = parallel.exchange_layout(batch.n)
        if args.shard == "records":
            assert aln_base == shard_lo and total_aln == shard_total, "record ranges do not tile the input"
    else:
        aln_base = 0; total_aln = batch.n

    def barrier_max(x):
        v = C.c_double(x)
        ctx._check(ctx.lib.svimgpu_barrier_max(ctx.h, C.byref(v)))
        return v.value

    xchg_ms = {}

    def cluster_step():
        if world > 1:
            st = _lib.CollectStats()
            ctx._check(ctx.lib.svimgpu_exchange_signatures(ctx.h, aln_base, C.byref(st)))
            # the timing slots are reset by the next call: keep the signature exchange apart from the cluster exchange
            xchg_ms["nccl_exchange_signatures"] = ctx.timings().get("nccl_exchange", 0.0)
            ctx.use_collected(0)
            return st, ctx.cluster(sharded=True, view=True)
        ctx.use_collected(0)
        return None, ctx.cluster(view=True)       # the context's pinned result arrays, no second host copy

    # ---------------- resident: record buffer already in HBM -------------------------------------
    ctx.upload(batch)
    stage_ms = {}
    launches0 = 0
    res_ms = []
    own_ms = []
    sampler = ClockSampler(local_rank, ctx)
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier_max(0.0); sampler.start(); launches0 = ctx.launch_count()
        ctx.timer_start()
        cst = ctx.collect()
        tm = dict(ctx.timings())
        xst, (clst, clusters, members) = cluster_step()
        ms = ctx.timer_stop()
        for k, v in list(tm.items()) + list(ctx.timings().items()) + list(xchg_ms.items()):
            if v and s >= args.warmup:
                stage_ms.setdefault(k, []).append(v)
        if s >= args.warmup:
            own_ms.append(ms)
            res_ms.append(barrier_max(ms))
    clocks = sampler.stop()
    clusters = np.array(clusters); members = np.array(members)      # own copies of the last step's result
    per_rank = None
    if world > 1:
        # every rank's own step time and stage table (the headline is the max over ranks): where the ranks differ, and how much of
        # a rank's step no stage accounts for
        import torch.distributed as dist
        st_mine = {k: round(float(np.mean(v)), 3) for k, v in stage_ms.items()}
        mine = {"rank": rank, "step_ms": round(float(np.mean(own_ms)), 3), "stages_sum_ms": round(sum(st_mine.values()), 3),
                "myers": st_mine.get("myers_edit_distance", 0.0), "linkage": st_mine.get("linkage", 0.0),
                "exchange_signatures": st_mine.get("nccl_exchange_signatures", 0.0), "exchange_clusters": st_mine.get("nccl_exchange", 0.0),
                "cluster_d2h": st_mine.get("cluster_d2h", 0.0), "myers_pairs": int(clst.myers_pairs), "myers_band_cells": int(clst.myers_band_cells)}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    launches = ctx.launch_count() - launches0
    n_sigs = (xst.n_signatures if xst else cst.n_signatures)
    if args.profile_steps:
        if args.profile_genotype and world == 1:
            sigs, _ins = ctx.fetch_signatures(0, cst)
            args.no_cpu_baseline = True
            print(json.dumps(genotype_leg(ctx, batch, clusters, members, sigs, args)))
        return

    # ---------------- e2e: host buffers through the C ABI, copies inside the timed region --------
    # the record buffer crosses PCIe with its CIGAR blob in a packed form of include/svimgpu.h (svim_aln_soa.cigar8 / cigar16): what the
    # host BAM decoder emits while it decodes (io.read_bam_native(pack_cigar=8)).  The synthetic batch never was a BAM file, so it is
    # re-encoded here ONCE, outside the timed region (csrc_host/bamio.cpp; its time is reported as e2e.cigar_pack_s); the expansion to
    # BAM's uint32 words in HBM is inside it
    t_pack = time.perf_counter()
    pack = 32 if args.no_cigar16 else args.cigar_pack
    if pack == 16:
        batch.pack_cigar16()
    elif pack == 8:
        batch.pack_cigar8()
    t_pack = time.perf_counter() - t_pack
    packed_arrays = [batch.cigar8, batch.cigar8_off] if batch.cigar8 is not None else [batch.cigar16, batch.cigar16_off] if batch.cigar16 is not None else [batch.cigar]
    pinned = packed_arrays + [batch.seq, batch.sa] + [getattr(batch, f) for f, _ in batch.FIELDS]
    for a in pinned:
        ctx.pin(a)
    # collect_host keeps SEQ on the host and uploads only the packed bases of emitted insertions (lazy SEQ)
    h2d_full = int(sum(a.nbytes for a in pinned))
    e2e_ms = []
    e2e_stage = {}
    d2h = 0
    if world > 1:
        # the job's result lands on ONE host: rank 0 mirrors the gathered records and every rank's insertion bytes (what the objects
        # are built from), the other ranks the gathered records and the cluster arrays only
        ctx._check(ctx.lib.svimgpu_mirror_gathered_ins(ctx.h, 1 if rank == 0 else 0))
    sampler_e2e = ClockSampler(local_rank, ctx)
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier_max(0.0); sampler_e2e.start()
        ctx.timer_start()
        cst = ctx.collect_host(batch)
        tm = dict(ctx.timings())
        xst, (clst, clusters, members) = cluster_step()
        tm2 = dict(ctx.timings())
        fst = xst or cst
        sigs, ins = ctx.fetch_signatures(0, fst)
        ms = ctx.timer_stop()
        if s >= args.warmup:
            for k, v in list(tm.items()) + list(tm2.items()) + list(xchg_ms.items()):
                if v:
                    e2e_stage.setdefault(k, []).append(v)
        d2h = 2 * sigs.nbytes + ins.nbytes + clusters.nbytes + members.nbytes     # signature records cross twice (staging + fetch)
        if world > 1:
            d2h = 48 * cst.n_signatures + sigs.nbytes + ins.nbytes + clusters.nbytes + members.nbytes      # collect's staging + the gathered lists (this is rank 0)
        h2d = h2d_full - batch.seq.nbytes + (cst.ins_bytes + 1) // 2 + 8 * cst.n_signatures
        if s >= args.warmup:
            e2e_ms.append(barrier_max(ms))
    clocks_e2e = sampler_e2e.stop()
    if world > 1:
        ctx._check(ctx.lib.svimgpu_mirror_gathered_ins(ctx.h, 1))
        if rank != 0:       # the parity block below compares full lists on every rank (no rank has collected again: the peers' bytes are intact)
            sigs, ins = (np.zeros(xst.n_signatures, dtype=_lib.SIG_DTYPE), np.zeros(xst.ins_bytes, dtype=np.uint8))
            ctx._check(ctx.lib.svimgpu_fetch_signatures(ctx.h, 0, sigs.ctypes.data, ins.ctypes.data))
    # ---------------- parity evidence, outside the timed regions --------------------------------------------
    # the device's sorted order + partition offsets of the last step (svimgpu_fetch_partitions), before anything re-clusters
    sigs = np.array(sigs); ins = np.array(ins)        # own copies: the pinned mirrors are reused by later calls
    clusters = np.array(clusters); members = np.array(members)
    part_order, part_off = ctx.fetch_partitions(len(sigs))
    parity = {"mismatches": 0}
    if world > 1:
        from svim_b200 import rows as svrows
        import torch.distributed as dist
        mine = svrows.result_digest(clusters, members, sigs)
        digests = [None] * world
        dist.all_gather_object(digests, mine)
        # the same gathered signature list clustered by ONE GPU (no partition sharding, no cluster exchange) must give the same bytes
        ctx.use_collected(0)
        _st1, cl1, mem1 = ctx.cluster()
        single = svrows.result_digest(cl1, mem1, sigs)
        singles = [None] * world
        dist.all_gather_object(singles, single)
        parity.update({"rank_digests": digests, "ranks_agree": len(set(digests)) == 1, "sharded_equals_single_gpu_cluster": all(x == mine for x in singles),
                       "note": "digest = blake2b-64 of (cluster records, member indices, gathered signature records); single = every rank re-clusters "
                               "the gathered signatures alone (svimgpu_cluster) and compares"})
        parity["mismatches"] += int(len(set(digests)) != 1) + int(any(x != mine for x in singles))
    # ---------------- e2e_python: the reference-shaped call pair, objects included ("... to 6-tuple of cluster lists
    # materialised in Python") — svim:102 analyze_alignment_file_coordsorted(bam, options) + svim:132 cluster_sv_signatures(sigs, options)
    # through svim_b200's host mirror on the same pinned host buffers, wall clock, every step uploads and builds every object ----------
    py_leg = None
    obj_s = None
    if world == 1:
        import gc
        from svim_b200 import runtime
        from svim_b200.SVIM_COLLECT import analyze_alignment_file_coordsorted
        from svim_b200.SVIM_CLUSTER import cluster_sv_signatures
        runtime._CTX[local_rank] = ctx                      # same context (and its device buffers) as the legs above
        os.environ["SVIM_B200_DEVICE"] = str(local_rank)
        runtime.register_genome("bench-genome", genome)
        ctx.genome_key = None
        opts = argparse.Namespace(min_mapq=20, min_sv_size=40, max_sv_size=100000, segment_gap_tolerance=10, segment_overlap_tolerance=5,
                                  partition_max_distance=1000, position_distance_normalizer=900, edit_distance_normalizer=1.0,
                                  cluster_max_distance=0.5, all_bnds=False, genome="bench-genome")
        py_ms = []; n_obj = 0
        for s in range(2 + max(2, min(args.steps, 5))):
            gc.collect()
            t0 = time.perf_counter()
            sv_sigs, _twins = analyze_alignment_file_coordsorted(batch, opts)
            t1 = time.perf_counter()
            six = cluster_sv_signatures(sv_sigs, opts)
            t2 = time.perf_counter()
            if s >= 2:
                py_ms.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
            n_obj = len(sv_sigs) + sum(len(x) for x in six)
            del sv_sigs, _twins, six                        # freeing 270 k objects is the caller's cost, not part of the call pair
        m = np.mean(np.asarray(py_ms), axis=0)
        py_leg = {"value": batch.n / (m[0] * 1e-3), "unit": UNIT, "ms_per_step": float(m[0]), "collect_call_ms": float(m[1]), "cluster_call_ms": float(m[2]),
                  "python_objects": int(n_obj), "note": "wall clock of analyze_alignment_file_coordsorted + cluster_sv_signatures (svim_b200 host mirror) from pinned "
                  "host record buffers to the 6-tuple of SignatureCluster lists; objects built by svim_b200/csrc_host/fastobj.c"}
        obj_s = None
    for a in pinned:
        ctx.unpin(a)
    geno_leg = genotype_leg(ctx, batch, clusters, members, sigs, args) if world == 1 else None
    bam_leg = None
    if not args.no_bam and world == 1:
        # ---------------- e2e_from_bam: the user's real starting point, a BAM file ----------------------------------
        # file (page cache) -> svimgpu_decode_bam (compressed bytes over PCIe; BGZF inflate, record boundaries, rows, blobs, read-name ids on
        # the device) -> svimgpu_collect on the resident buffer -> svimgpu_cluster -> signature + cluster records on the host.  Wall clock.
        import tempfile
        from svim_b200 import io as sio
        try:
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "bench.bam")
                t0 = time.perf_counter(); sio.write_bam_native(path, batch, level=1, threads=os.cpu_count() or 8); t_w = time.perf_counter() - t0
                size = os.path.getsize(path)
                runs = []
                for rep in range(3):
                    st_bam = {}
                    fb0 = sio.BAM_DECODE_COUNTS["host_fallback"]
                    t0 = time.perf_counter()
                    rb = sio.decode_bam_resident(path, ctx, st_bam)
                    t_dec = time.perf_counter() - t0
                    on_gpu = isinstance(rb, sio.ResidentBatch)
                    if on_gpu:
                        cst2 = ctx.collect()
                    else:
                        cst2 = ctx.collect_host(rb)
                    ctx.use_collected(0); clst2, cl2, mem2 = ctx.cluster()
                    sg2, in2 = ctx.fetch_signatures(0, cst2)
                    t_all = time.perf_counter() - t0
                    runs.append((t_all, t_dec, st_bam, on_gpu, sio.BAM_DECODE_COUNTS["host_fallback"] - fb0))
                from svim_b200 import rows as svrows
                same = svrows.result_digest(cl2, mem2) == svrows.result_digest(clusters, members) and int(cst2.n_signatures) == int(cst.n_signatures)
                t_all, t_dec, st_bam, on_gpu, n_fb = min(runs, key=lambda r: r[0])
                t0 = time.perf_counter(); _host = sio.read_bam_native(path, threads=os.cpu_count() or 8); t_host = time.perf_counter() - t0
                n_rec = _host.n; del _host
            bam_leg = {"value": n_rec / t_all, "unit": UNIT, "total_s": t_all, "bam_bytes": size, "decoder": "gpu (svimgpu_decode_bam)" if on_gpu else "host (fallback)",
                       "host_fallbacks": int(n_fb), "decode_s": t_dec, "bam_index_s": st_bam.get("index_s"),
                       "stages_ms": {k: round(v, 3) for k, v in st_bam.items() if k.startswith("bam_") and k != "bam_bytes"},
                       "inflated_bytes": st_bam.get("inflated_bytes"), "host_decoder_s": t_host, "host_threads": os.cpu_count(), "bam_write_s": t_w,
                       "same_result_as_host_buffers": bool(same), "runs_s": [round(r[0], 4) for r in runs],
                       "note": "best of 3; BAM in the page cache -> compressed file H2D + on-device decode -> COLLECT + CLUSTER on the resident buffer -> records D2H; "
                               "host_decoder_s is svim_b200's multi-threaded host BAM decoder alone on the same file"}
        except Exception as e:          # the leg is reported, never fatal to the bench line
            bam_leg = {"error": (type(e).__name__ + ": " + str(e))[:400]}

    if rank != 0:
        return
    ms_step = float(np.mean(res_ms)); e2e_step = float(np.mean(e2e_ms))
    scan_ms = float(np.mean(stage_ms.get("cigar_scan", [float("nan")])))
    alg_bytes = batch.algorithmic_bytes() + int(cst.n_signatures) * 48
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    traffic = None      # dram__bytes_read+write of k_cigar_scan from the committed ncu capture of this same workload
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        if t.get("workload") == args.workload and abs(t.get("scale", 0) - args.scale) < 1e-9:
            traffic = t["dram_bytes_per_launch"]
    out = {
        "metric": METRIC, "value": total_aln / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.shard == "records" else "weak", "vs_baseline": None, "dtype": "u32 CIGAR words / i64 coordinates / f64 distances",
        "data": "synthetic", "gpu_launches": int(launches),
        "config": {"workload": workload_name(args), "alignments_per_gpu": batch.n, "signatures": int(n_sigs), "clusters": int(clst.n_clusters_total),
                   "myers_pairs": int(clst.myers_pairs), "myers_cells": int(clst.myers_cells),
                   "myers_banded_pairs": int(clst.myers_banded_pairs), "myers_handed_over": int(clst.myers_retry_pairs),
                   "myers_band_cells": int(clst.myers_band_cells),
                   "l2": "inputs larger than L2 (%.2f GB CIGAR per GPU vs 126 MB)" % (batch.cigar.nbytes / 1e9),
                   "parallelism": (("contiguous record ranges of one input balanced by CIGAR volume" if args.shard == "records" else "records sharded by contig")
                                   + "; 2 NCCL allgatherv (signature records, cluster records); inserted sequences "
                                   + ("read from the owning rank through NVSwitch peer memory (CUDA IPC)" if ctx.lib.svimgpu_peer_ins_active(ctx.h) else "gathered with the records")
                                   ) if world > 1 else "single GPU",
                   "input_generation_s": round(t_gen, 1)},
        "e2e": {"value": total_aln / (e2e_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_step, "stages_ms": {k: round(float(np.mean(v)), 4) for k, v in e2e_stage.items()},
                "cigar_upload": ("8-bit packed stream (svim_aln_soa.cigar8), %.2f GB; expanded on the device" % (batch.cigar8.nbytes / 1e9)) if batch.cigar8 is not None
                                else ("16-bit packed stream (svim_aln_soa.cigar16), %.2f GB; expanded on the device" % (batch.cigar16.nbytes / 1e9)) if batch.cigar16 is not None
                                else "uint32 BAM words, %.2f GB" % (batch.cigar.nbytes / 1e9),
                "cigar_pack_s": round(t_pack, 3),
                "result_on_host": ("rank 0: gathered signature records + every rank's insertion bytes + cluster arrays; other ranks: gathered records + cluster arrays"
                                   if world > 1 else "signature records + insertion bytes + cluster arrays"),
                "clocks": clocks_e2e},
        "roofline": {"kernel": "k_cigar_scan", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": scan_ms},
        "stages_ms": {k: round(float(np.mean(v)), 4) for k, v in stage_ms.items()},
        "clocks": clocks,
    }
    if per_rank:
        out["per_rank"] = per_rank
    if py_leg:
        out["e2e_python"] = py_leg
    if bam_leg:
        out["e2e_from_bam"] = bam_leg
    if geno_leg:
        out["genotype"] = geno_leg
    if clst.myers_cells and "myers_edit_distance" in out["stages_ms"]:
        # The time-dominant stage.  computed_cells = what the kernels really touched: window cells of the first pass (thread-per-pair
        # windows: columns x 32 x blocks; wavefront bands) + full matrices of the pairs that run unbanded (their own + hand-overs).
        # Peak: the block step of k_myers_tpp is 9 ALU-pipe instructions (7 LOP3 + 2 SHF) per 32 cells per lane, the ALU pipe takes one
        # warp instruction every 2 clocks per scheduler (B300_MICROARCH.md "Pipe rates"): SMs x 4 schedulers x 32 lanes x 32 cells / 18 clk.
        t_my = out["stages_ms"]["myers_edit_distance"] * 1e-3
        computed = int(clst.myers_band_cells) + int(clst.myers_unbanded_cells)
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        peak = 148 * 4 * 32 * 32 / 18.0 * sm_clock / 1e12
        out["roofline_myers"] = {"kernel": "k_myers_tpp<B> (+ k_myers_band / k_myers_fast for what it hands over)", "bound": "int-alu",
                                 "computed_cells": computed, "ms": t_my * 1e3, "tcups": computed / t_my / 1e12, "peak_tcups": peak,
                                 "frac": computed / t_my / 1e12 / peak,
                                 "peak_derivation": "148 SMs x 4 schedulers x 32 lanes x 32 cells per block step / (9 ALU-pipe instr x 2 clk) x %.0f MHz" % (sm_clock / 1e6),
                                 "full_matrix_cells": int(clst.myers_cells), "effective_tcups": clst.myers_cells / t_my / 1e12,
                                 "pairs": int(clst.myers_pairs), "first_pass_pairs": int(clst.myers_banded_pairs), "tpp_pairs": int(clst.myers_tpp_pairs),
                                 "tpp_cells": int(clst.myers_tpp_cells), "handed_over_pairs": int(clst.myers_retry_pairs),
                                 "unbanded_cells": int(clst.myers_unbanded_cells)}
    if not args.no_cpu_baseline:
        # rank 0's own records are the head of the gathered list, so the prefix check holds at every N (shorter prefix at N > 1,
        # where no cpu_baseline is reported)
        from svim_b200 import rows as svrows
        keep = {}
        base = cpu_baseline(batch, genome, target_seconds=args.ref_seconds if world == 1 else min(args.ref_seconds, 4.0), keep=keep)
        if world == 1:
            out["cpu_baseline"] = base
        pp = svrows.prefix_parity(batch, keep["n"], sigs, ins, clusters, members, part_order, part_off, keep["signatures"], keep["clusters"])
        parity.update(pp); parity["mismatches"] = pp["mismatches"] + (parity.get("mismatches", 0) if world > 1 else 0)
        parity["against"] = "oracle port of the reference on the first %d records; clusters of every partition made only of their signatures" % keep["n"]
    out["parity"] = parity
    print(json.dumps(out))


if __name__ == "__main__":
    main()
