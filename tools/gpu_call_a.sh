#!/bin/bash
# one GPU-box pass: new-row tests first, then the whole -m gpu suite, bench with the BAM leg (phase trace), ncu of the genotype kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_genotype.py -m gpu -x -q > gpurun_out/d1_geno_tests.log 2>&1; echo "geno tests rc=$?"
tail -15 gpurun_out/d1_geno_tests.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/d1_tests.log 2>&1; echo "all gpu tests rc=$?"
tail -5 gpurun_out/d1_tests.log
SVIM_BAMIO_TRACE=1 timeout 900 python bench.py --with-bam > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; echo "bench rc=$?"
grep bamio gpurun_out/bench_v13.err | tail -24
timeout 600 bash tools/gpu_profile.sh v13 geno > /dev/null 2>&1
ncu -i gpurun_out/prof_geno_v13.ncu-rep --page raw --csv > gpurun_out/prof_geno_v13_raw.csv 2>/dev/null
tail -3 gpurun_out/prof_geno_v13.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
cat gpurun_out/bench_v13.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({k:d[k] for k in ('value','ms_per_step','genotype','e2e_from_bam') if k in d}, indent=1)); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
