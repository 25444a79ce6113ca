#!/bin/bash
# final round-1 GPU pass: whole -m gpu suite, bench with the BAM leg (decoder trace), memcheck over the new kernels' small tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/d3_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -4 gpurun_out/d3_tests.log
SVIM_BAMIO_TRACE=1 timeout 600 python bench.py --with-bam > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14.err; echo "bench rc=$?"
grep bamio gpurun_out/bench_v14.err | tail -12
tail -3 gpurun_out/bench_v14.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_v14.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "clocks", "e2e_from_bam")}, indent=1))
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("clocks"), "roofline", d["roofline"]["frac"])
print("stages", d["stages_ms"])
g = d.get("genotype") or {}
print("genotype", g.get("value"), g.get("prepare_ms"), g.get("cpu_baseline"))
PY
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_genotype.py -m gpu -q -k "golden or edge or cutpaste" > gpurun_out/sanitizer_memcheck_geno.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitizer_memcheck_geno.log
