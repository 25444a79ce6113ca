"""A/B of kernel variants on one synthetic workload: every setting runs in a fresh context on the same records.

    python tools/ab_bench.py --workload config2 SVIM_MYERS_MODE=0 SVIM_MYERS_MODE=1 SVIM_SCAN_VARIANT=5+SVIM_MYERS_MODE=2

Prints one JSON line per setting: resident ms/step, the scan and Myers stage times, and a checksum of the clusters
(all settings must agree).  Development tool; bench.py is the measured contract.
"""
import argparse, hashlib, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("settings", nargs="*", default=[""])
    args = ap.parse_args()
    from svim_b200 import _lib
    batch, genome = bench.make_rank_input(args.workload, args.scale, 0, 1)
    alg_bytes = batch.algorithmic_bytes()
    for setting in args.settings:
        env = dict(kv.split("=") for kv in setting.split("+") if kv)
        for k in ("SVIM_MYERS_MODE", "SVIM_MYERS_BAND", "SVIM_MYERS_TPP", "SVIM_SCAN_VARIANT", "SVIM_SCAN_CHUNKS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ctx = _lib.Context(device=0)
        ctx.set_contigs(batch.contig_names); ctx.set_genome(genome); ctx.upload(batch)
        ms, stages = [], {}
        for s in range(args.warmup + args.steps):
            ctx.timer_start()
            cst = ctx.collect(); tm = dict(ctx.timings())
            ctx.use_collected(0); clst, clusters, members = ctx.cluster()
            t = ctx.timer_stop()
            if s >= args.warmup:
                ms.append(t)
                for k, v in list(tm.items()) + list(ctx.timings().items()):
                    if v: stages.setdefault(k, []).append(v)
        digest = hashlib.sha1(np.ascontiguousarray(clusters).tobytes() + np.ascontiguousarray(members).tobytes()).hexdigest()[:12]
        scan = float(np.mean(stages.get("cigar_scan", [float("nan")])))
        print(json.dumps({"setting": setting, "ms_per_step": round(float(np.mean(ms)), 3), "cigar_scan_ms": round(scan, 4),
                          "scan_GBps": round((alg_bytes + cst.n_signatures * 48) / scan / 1e6, 1),
                          "myers_ms": round(float(np.mean(stages.get("myers_edit_distance", [float("nan")]))), 3),
                          "pairs": int(clst.myers_pairs), "banded": int(clst.myers_banded_pairs), "retry": int(clst.myers_retry_pairs),
                          "tpp_pairs": int(clst.myers_tpp_pairs), "tpp_cells": int(clst.myers_tpp_cells), "band_cells": int(clst.myers_band_cells), "band_cells_frac": round(clst.myers_band_cells / max(1, clst.myers_cells), 4), "signatures": int(cst.n_signatures), "clusters": int(len(clusters)), "digest": digest}), flush=True)
        del ctx


if __name__ == "__main__":
    main()
