#!/bin/bash
mkdir -p gpurun_out
for k in k_miller k_final_exp; do
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o /tmp/prof_$k -f \
     python bench.py --steps 1 --warmup 3 --verify-log2n 0 --extras 0 --cpu-seconds 1 > gpurun_out/bench_ncu_$k.log 2>&1; echo "ncu $k exit $?"
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/r01l_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv > /tmp/src_$k.csv 2>/dev/null
  python tools/opcode_stalls.py /tmp/src_$k.csv > gpurun_out/r01l_${k}_opcodes.md 2>&1
done
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01l.csv \
   python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list exit $?"
ls -la gpurun_out/
