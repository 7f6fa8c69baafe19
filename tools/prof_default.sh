#!/bin/bash
# default-mode launch list + one full capture of the resize kernel (run on the GPU box)
O=gpurun_out
CMD="python bench.py --mode default --steps 2 --warmup 3 --min-seconds 0 --no-cpu-baseline --no-extras --no-e2e"
$CMD > $O/r2y_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2y_launches_default.csv $CMD > $O/r2y_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_resize_rows -s 3 -c 1 -o $O/r2y_rows $CMD > $O/r2y_ncu_full.log 2>&1
ls -la $O/r2y_rows.ncu-rep
