#!/bin/bash
tag=r02
mkdir -p gpurun_out
one() {
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/${tag}_$1 \
      python tools/prof_driver.py 20 > gpurun_out/${tag}_prof_$1.log 2>&1
  echo "$1: $(grep -c ' 39 passes\| 4[0-9] passes' gpurun_out/${tag}_prof_$1.log) full captures"
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_$1.csv 2>/dev/null
  ncu -i /tmp/${tag}_$1.ncu-rep --page source --csv > gpurun_out/${tag}_src_$1.csv 2>/dev/null
  gzip -f gpurun_out/${tag}_src_$1.csv
}
one k_final_exp 'k_final_exp$' 2 2
one k_check_products 'k_check_products' 1 1
