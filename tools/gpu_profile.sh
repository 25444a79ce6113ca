#!/bin/bash
# One GPU-box pass of ncu over config2 (bench.py --profile-steps runs only resident steps, 1 warm-up + 1 step here).
# Outputs under gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
#   tools/gpu_profile.sh <tag> [launches] [scan] [myers]
set -u
tag=${1:-vX}; shift
parts=${*:-launches scan myers}
mkdir -p gpurun_out
for p in $parts; do
  case $p in
    launches) ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
                python bench.py --profile-steps --steps 1 --warmup 1 > gpurun_out/launches_${tag}.log 2>&1 ;;
    scan)     ncu --set full --clock-control none --import-source on -k regex:k_cigar_scan -s 1 -c 1 -f -o gpurun_out/prof_scan_${tag} \
                python bench.py --profile-steps --steps 1 --warmup 1 > gpurun_out/prof_scan_${tag}.log 2>&1 ;;
    myers)    # the banded shapes of the second step (5 shapes carry pairs on config2), then the unbanded bins
              ncu --set full --clock-control none --import-source on -k regex:k_myers_band -s 5 -c 5 -f -o gpurun_out/prof_myers_band_${tag} \
                python bench.py --profile-steps --steps 1 --warmup 1 > gpurun_out/prof_myers_band_${tag}.log 2>&1 ;;
    geno)     ncu --set full --clock-control none --import-source on -k regex:"k_ref_end|k_genotype" -c 3 -f -o gpurun_out/prof_geno_${tag} \
                python bench.py --profile-steps --profile-genotype --steps 1 --warmup 1 > gpurun_out/prof_geno_${tag}.log 2>&1 ;;
  esac
done
ls -la gpurun_out/
