#!/bin/bash
# A/B of the front-end kernels on the bench workload (run on the GPU box): one JSON line per variant
out=${1:-gpurun_out/ab.jsonl}; : > $out
for bs in 384 20; do
  for fe in auto umma umma-apron mma-sync; do
    if [ "$bs" = "20" ] && [ "$fe" = "auto" ]; then continue; fi
    timeout 200 python bench.py --steps 20 --warmup 3 --min-seconds 1 --no-e2e --no-cpu-baseline --no-extras --blur-scale $bs --front-end $fe 2>>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(json.dumps({'blur_scale': $bs, 'front_end': '$fe', 'gaussian': d['config']['gaussian'], 'fps': d['value'], 'ms_per_step': d['ms_per_step'], 'groups': d['roofline']['groups_ms_per_step'], 'frac': d['roofline']['frac']}))
" >> $out
  done
done
cat $out
