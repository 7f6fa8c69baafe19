#!/bin/bash
# default mode, row-per-lane against warp-per-row INTER_AREA kernels, over frame sizes (one B200)
O=gpurun_out; out=$O/ab_resize_sizes.log; : > $out
timeout 600 python -m pytest tests/test_gpu_resize.py -x -q --timeout 300 2>&1 | tail -2 | tee -a $out
line() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1', d['value'], d['roofline']['groups_ms_per_step'], d['roofline']['frac'])
"; }
for cfg in "1920x1080 8 32" "3840x2160 2 16" "1280x720 8 32" "640x480 8 32"; do
  set -- $cfg
  B="python bench.py --mode default --size $1 --streams $2 --ring $3 --steps 20 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras"
  timeout 200 $B --front-end warp-resize 2>>$O/ab.err | line "$1 warp" >> $out
  timeout 200 $B 2>>$O/ab.err | line "$1 rows" >> $out
done
cat $out
