#!/bin/bash
# One single-GPU evidence session (run under gpurun): tests, bench line, full-size checks, ncu launch list and top-kernel profile.
# Everything lands in gpurun_out/<tag>_*; the .ncu-rep stays on the box (only CSV pages come back).
tag=${1:-r2q}
mkdir -p gpurun_out
echo "== BAM decoder token statistics, inline on / off"
SVIM_BAM_DEBUG=1 SVIM_BAM_INLINE=1 timeout 200 python tools/prof_bam.py 2>&1 | tail -2
SVIM_BAM_DEBUG=1 SVIM_BAM_INLINE=0 timeout 200 python tools/prof_bam.py 2>&1 | tail -2
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
echo "== bench"
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
echo "== full-size checks"
for w in config2 config3 config5; do
  timeout 900 python tools/fullsize_check.py --workload $w > gpurun_out/${tag}_fullsize_$w.json 2> gpurun_out/${tag}_fullsize_$w.err; echo "$w rc=$?"; tail -c 600 gpurun_out/${tag}_fullsize_$w.json; tail -2 gpurun_out/${tag}_fullsize_$w.err
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-steps --steps 1 --warmup 1 > gpurun_out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
echo "== ncu --set full, k_myers_tpp (second step's 18 launches) and k_cigar_scan"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_myers_tpp|k_cigar_scan_s" -s 19 -c 19 -o /tmp/prof_top python bench.py --profile-steps --steps 1 --warmup 1 > gpurun_out/${tag}_prof.log 2>&1; echo "profile rc=$?"
ncu -i /tmp/prof_top.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_top_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
