#!/bin/bash
# Multi-GPU evidence session (run under `gpurun --gpus N`): golden parity at world N, H2D topology, weak-scaling bench line,
# record-range sharded config4 and config5.  Usage: gpu_session_multi.sh <tag> <N> [what...]   what: check topo weak c4 c5
tag=$1; N=$2; shift 2
what=${@:-check topo weak c4 c5}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
nproc; free -g | head -2
for w in $what; do
  case $w in
    check) echo "== golden parity at world $N"; timeout 300 bash -c "$(declare -f run); N=$N; run 29501 tests/multi_gpu_check.py" > gpurun_out/${tag}_check.log 2>&1; echo "rc=$?"; grep "multi-gpu ok" gpurun_out/${tag}_check.log | tail -4 ;;
    topo)  echo "== H2D topology"; timeout 300 bash -c "$(declare -f run); N=$N; run 29502 tools/h2d_topo_check.py" > gpurun_out/${tag}_topo.log 2>&1; echo "rc=$?"; grep -A12 "^rank" gpurun_out/${tag}_topo.log | head -14 ;;
    weak)  echo "== bench, weak scaling (config2 per GPU)"; timeout 500 bash -c "$(declare -f run); N=$N; run 29503 bench.py --gpus $N --steps 5 --warmup 3" > gpurun_out/${tag}_bench_weak.json 2> gpurun_out/${tag}_bench_weak.err; echo "rc=$?"; tail -2 gpurun_out/${tag}_bench_weak.err ;;
    c4)    echo "== bench, config4 x0.25 sharded by record ranges"; timeout 600 bash -c "$(declare -f run); N=$N; run 29504 bench.py --gpus $N --steps 5 --warmup 3 --workload config4 --scale 0.25 --shard records" > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err; echo "rc=$?"; tail -2 gpurun_out/${tag}_bench_config4.err ;;
    c5)    echo "== bench, config5 x0.25 sharded by record ranges"; timeout 600 bash -c "$(declare -f run); N=$N; run 29505 bench.py --gpus $N --steps 5 --warmup 3 --workload config5 --scale 0.25 --shard records" > gpurun_out/${tag}_bench_config5.json 2> gpurun_out/${tag}_bench_config5.err; echo "rc=$?"; tail -2 gpurun_out/${tag}_bench_config5.err ;;
  esac
done
ls -la gpurun_out | tail -12
