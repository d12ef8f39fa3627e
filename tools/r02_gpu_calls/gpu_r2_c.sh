#!/bin/bash
# round 2, call C: why is the 16-lane mapping slower? ncu of one dense k_iterate launch per mapping
mkdir -p gpurun_out
for L in 16 32; do
  CB200_LANES=$L CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 > gpurun_out/r2c_plain_$L.log 2>&1
  CB200_LANES=$L CB200_GROUPS=1 CB200_NO_PERSISTENT=1 CB200_NO_LIVE_LIST=1 python tools/prof_selfplay.py 4096 800 300 bf16 > gpurun_out/r2c_plain_nolist_$L.log 2>&1
  CB200_LANES=$L CB200_GROUPS=1 CB200_NO_PERSISTENT=1 ncu --set full --clock-control none --import-source on -k regex:"k_iterate" -s 250 -c 1 -o gpurun_out/r2c_prof_$L -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/r2c_ncu_$L.log 2>&1
  echo "ncu $L rc=$?"
  ncu -i gpurun_out/r2c_prof_$L.ncu-rep --page raw --csv > gpurun_out/r2c_raw_$L.csv 2>/dev/null
  ncu -i gpurun_out/r2c_prof_$L.ncu-rep --page source --print-source cuda,sass --csv > /tmp/dense_$L.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/dense_$L.csv > gpurun_out/r2c_lines_$L.txt
done
cat gpurun_out/r2c_plain_16.log gpurun_out/r2c_plain_32.log gpurun_out/r2c_plain_nolist_16.log
rm -f gpurun_out/r2c_prof_32.ncu-rep
