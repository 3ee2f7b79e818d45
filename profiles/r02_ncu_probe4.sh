#!/bin/bash
# per-kernel durations (ncu launch list) of the cfg4-shape probe at one subset size; usage: r02_ncu_probe4.sh K [POP]
k=${1:-200}; pop=${2:-148}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_probe4_k${k}_launches.csv python profiles/k1_probe4.py $k $pop > gpurun_out/r02_probe4_k${k}_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_probe4_k${k}_launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-12:]: print(r[4][:60], r[-1], r[-2])
PY
