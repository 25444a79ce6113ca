#!/bin/bash
# Final single-GPU validation of a round (run under gpurun): the driver's own sequence (tests, smoke, default bench line) plus the
# N=1 points of the record-range sharded workloads.  Usage: gpu_session_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (default line)"
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
for w in config5 config4; do
  echo "== bench $w x0.25, one GPU"
  timeout 300 python bench.py --workload $w --scale 0.25 --shard records --no-bam > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "rc=$?"; tail -2 gpurun_out/${tag}_bench_$w.err
done
ls -la gpurun_out | tail -8
