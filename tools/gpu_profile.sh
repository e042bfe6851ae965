#!/bin/bash
# usage (GPU box, under gpurun): tools/gpu_profile.sh <tag>   ->   gpurun_out/<tag>_*.csv(.gz)
# 1. launch list of the default bench command (per-launch times, cold caches, serialised)
# 2. `ncu --set full` of one launch of each hot kernel at 2^20 items (tools/prof_driver.py), exported as raw + source CSV
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches_ncu.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_miller|k_final_exp|k_hash_to_g1|k_glued|k_g1_mul|k_g2_mul|k_check_products' -s 8 -c 8 \
    -f -o gpurun_out/${tag}_prof python tools/prof_driver.py 20 > gpurun_out/${tag}_prof.log 2>&1
tail -2 gpurun_out/${tag}_prof.log
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_raw.csv 2>/dev/null
for kn in k_miller k_final_exp k_hash_to_g1 k_glued k_g1_mul k_g2_mul; do
  ncu -i gpurun_out/${tag}_prof.ncu-rep --page source --csv -k regex:"${kn}" > gpurun_out/${tag}_src_${kn}.csv 2>/dev/null
  gzip -f gpurun_out/${tag}_src_${kn}.csv
done
ls -la gpurun_out/ | grep ${tag}
rm -f gpurun_out/${tag}_prof.ncu-rep
