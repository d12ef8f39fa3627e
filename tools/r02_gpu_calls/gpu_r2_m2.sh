#!/bin/bash
# profiles of the closing build (lock-step game step at 6 CTAs per SM): launch list of the bench command and the
# full-set ncu capture of a dense k_iterate / k_mlp_tc launch
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m2_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5100 -c 600 --csv --log-file gpurun_out/r2m2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m2_ncu_launches.log 2>&1
echo "launch list rc=$?"
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/r2m2_plain_prof.log 2>&1 && \
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_iterate|k_mlp_tc" -s 500 -c 2 -o gpurun_out/r2m2_dense -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/r2m2_ncu_dense.log 2>&1
echo "dense rc=$?"
ncu -i gpurun_out/r2m2_dense.ncu-rep --page raw --csv > gpurun_out/r2m2_dense_raw.csv 2>/dev/null
rm -f gpurun_out/r2m2_dense.ncu-rep
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2m2_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]; v = float(r[-1].replace(",", ""))
    unit = r[-2]
    if unit == "ns": v /= 1000.0
    elif unit == "ms": v *= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-46s %5d launches %10.1f us total %8.1f us avg %6.1f%% of kernel time" % (k[:46], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
PY
tail -2 gpurun_out/r2m2_plain_prof.log
