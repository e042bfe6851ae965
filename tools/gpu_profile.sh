#!/bin/bash
# usage (GPU box, under gpurun): tools/gpu_profile.sh <tag>   ->   gpurun_out/<tag>_*.csv(.gz)
# 1. launch list of the default bench command (per-launch times, cold caches, serialised)
# 2. `ncu --set full` of one launch of each hot kernel at 2^20 items (tools/prof_driver.py runs every kernel twice; the
#    first round is skipped), ONE ncu session per kernel: in a session with many kernels ncu collected only the
#    source-level passes for some of them (6 instead of 39 passes, hardware counters "-nan").
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches_ncu.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
one() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/${tag}_$1 \
      python tools/prof_driver.py 20 > gpurun_out/${tag}_prof_$1.log 2>&1
  echo "$1: $(grep -c ' 39 passes\| 4[0-9] passes' gpurun_out/${tag}_prof_$1.log) full captures, $(grep -c ' 6 passes' gpurun_out/${tag}_prof_$1.log) source-only"
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_$1.csv 2>/dev/null
  ncu -i /tmp/${tag}_$1.ncu-rep --page source --csv > gpurun_out/${tag}_src_$1.csv 2>/dev/null
  gzip -f gpurun_out/${tag}_src_$1.csv
  rm -f /tmp/${tag}_$1.ncu-rep
}
one k_miller 'k_miller$' 1 1
one k_final_exp 'k_final_exp$' 2 2
one k_hash_to_g1 'k_hash_to_g1' 2 1
one k_glued13 'k_glued' 3 1
one k_glued40 'k_glued' 4 1
one k_g1_mul 'k_g1_mul$' 1 1
one k_g2_mul 'k_g2_mul$' 1 1
one k_check_products 'k_check_products' 1 1
ls -la gpurun_out/ | grep ${tag}_ | awk '{print $5, $9}'
